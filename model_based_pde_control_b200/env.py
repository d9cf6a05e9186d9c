"""``KSVecEnv`` -- GPU-resident vectorised drop-in for ``KuramotoSivashinskyEnv-v0``.

Host-side mirror of the reference's env interface over the C ABI of ``libks_b200.so``:

* constructor kwargs, attributes and helper objects of the single env
  (``pdegym/kuramoto/kuramoto.py:29-76,131-150``): ``L, N, cfg_steps, Tmax, dt, sigma, dx, x,
  max_episode_steps, forcing, noop, reward_func, rhs, scenario, time``;
* the gym 0.25.2 vector-env protocol as the reference consumes it
  (``pdecontrol/mbrl/mbrl.py:81-86``, ``pdegym/common/vec_wrappers.py``,
  ``pdecontrol/mbrl/worker.py:48-88``): ``reset``, ``step_async`` / ``step_wait`` / ``step``
  returning ``obs (B,1,N) float32``, ``rewards (B,) float64``, ``terminated`` (always False),
  ``truncated``, ``infos`` with ``"step"`` and -- on the truncation step -- ``"final_observation"``
  / ``"_final_observation"``, followed by the auto-reset (burn-in of 800 no-op periods);
* a device-tensor API (``step_device`` / ``rollout_device``) for policies that live on the GPU.

All numerics run in the CUDA kernels; nothing here computes the PDE on the CPU.
"""
from __future__ import annotations

import ctypes
import math
import os
import sys
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .forcing import GaussianForcing
from .spaces import Box, VectorEnvBase

BURNIN_TIME = 200.0    # kuramoto.py:103
IC_AMPLITUDE = 0.4     # kuramoto.py:106
DEFAULT_XI = (0.0, 0.25, 0.5, 0.75)   # kuramoto.py:18


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class KSVecEnv(VectorEnvBase):
    """``num_envs`` independent Kuramoto-Sivashinsky control environments on one B200.

    Parameters mirror ``KuramotoSivashinskyEnv.__init__`` (``config`` dict or keywords) plus:

    ``Xi``            relative jet positions (a class attribute in the reference, kuramoto.py:18)
    ``device``        CUDA device ordinal (default: ``LOCAL_RANK`` or the current device)
    ``precision``     ``"f64"`` (parity mode, default) or ``"f32"``
    ``reward_mode``   ``"l2"`` / ``"dissipation"``; default follows the reference's selector
                      ``l2control if self.objective else dissipation`` (kuramoto.py:72), i.e. any
                      non-empty ``objective`` string -- including the default ``"dissipation"`` --
                      selects the L2 reward
    ``ic``            ``"numpy"``: ``reset(seed=s)`` draws env ``i``'s initial condition on the host
                      exactly like ``np.random.seed(s+i); np.random.uniform(-0.4,0.4,N)``;
                      ``"device"``: counter-based Philox on the GPU.  ``reset(seed=None)`` and
                      auto-resets always use the device generator with a fresh OS seed (the
                      reference reseeds from OS entropy there, kuramoto.py:101).
    ``burnin_periods`` override of ``int(200/dt/cfg_steps)`` (= 800) no-op periods in ``reset``
    ``reset_mode``    ``"burnin"`` (default, the reference's semantics: an auto-reset runs the burn-in
                      launch before ``step`` returns) or ``"pool"``: auto-resets inject states that a
                      side stream burned in ahead of time (``reset_pool.ResetPool``, ``pool_slots``
                      batches in flight); explicit ``reset()`` calls always burn in synchronously.
    ``solver``        ``"fd_rk4"`` (default): the reference's own scheme -- periodic finite differences +
                      classic RK4 (kuramoto.py:83-90,118-129), the parity path.  ``"etdrk4"``: the
                      pseudo-spectral exponential integrator the north star names (Cox & Matthews
                      2002, Kassam & Trefethen 2005; hand-written FFT kernel, N = 64, L2 reward).  It
                      is NOT the reference's discretisation (states differ by ~2e-3 per control
                      period at the default grid); ``dt`` / ``cfg_steps`` are then the ETDRK4 step
                      and the steps per period, e.g. ``dt=0.025, cfg_steps=10`` for the reference's
                      0.25 time units per control period.  ``dealias`` = 2/3 rule on ``(u^2)_x``.
    ``env_index_base`` global index of this env's member 0 (a shard of a larger batch): enters the counter of
                      the device generator, so that ``reset_device(seed)`` of a sharded run draws the initial
                      conditions the single-GPU run draws, whatever the GPU count.
    ``sensor_stride`` observation sampling fused into the kernel's output stage: observations are
                      ``u[..., stride//2::stride]`` as ``SensorTransform(stride)`` would return
                      (``pdegym/common/transforms.py:231-247``); 1 = full state, which is what the
                      MBRL loop uses (``mbrl.py:171,174``).  The state itself is always full.
    """

    metadata = {"render.modes": ["rgb_array"]}
    reward_range = (-float("inf"), float("inf"))
    eps = np.finfo(np.float32).eps

    def __init__(self, num_envs: int, config: Optional[dict] = None, *, Xi: Optional[Sequence[float]] = None,
                 device: Optional[int] = None, precision: str = "f64", reward_mode: Optional[str] = None,
                 ic: str = "numpy", burnin_periods: Optional[int] = None, points_per_lane: int = 0,
                 sensor_stride: int = 1, copy: bool = True, reset_mode: str = "burnin", pool_slots: int = 2,
                 solver: str = "fd_rk4", dealias: bool = True, env_index_base: int = 0, **kwargs):
        cfg = dict(config or {})
        cfg.update(kwargs)
        self.L = float(cfg.pop("L", 22.0))
        self.N = int(cfg.pop("N", 64))
        self.cfg_steps = int(cfg.pop("cfg_steps", 250))
        self.Ttrans = cfg.pop("Ttrans", 40)
        self.Tmax = float(cfg.pop("Tmax", 100.0))
        self.dt = float(cfg.pop("dt", 0.001))
        self.noise = cfg.pop("noise", 0.1)
        self.sigma = float(cfg.pop("sigma", 0.4))
        self.lmbda = cfg.pop("lmbda", 0.0)
        self.objective = cfg.pop("objective", "dissipation")
        if cfg:
            raise TypeError(f"unexpected config keys: {sorted(cfg)}")
        self.Xi = list(DEFAULT_XI if Xi is None else Xi)
        if reward_mode is None:
            reward_mode = "l2" if self.objective else "dissipation"     # kuramoto.py:72
        if reward_mode not in _lib.REWARD_MODES:
            raise ValueError(f"reward_mode must be one of {sorted(_lib.REWARD_MODES)}")
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        if solver not in _lib.SOLVERS:
            raise ValueError(f"solver must be one of {sorted(_lib.SOLVERS)}")
        self.solver, self.dealias = solver, bool(dealias)
        if ic not in ("numpy", "device"):
            raise ValueError("ic must be 'numpy' or 'device'")
        self.reward_mode, self.precision, self.ic, self.copy = reward_mode, precision, ic, copy
        if reset_mode not in ("burnin", "pool"):
            raise ValueError("reset_mode must be 'burnin' or 'pool'")
        self.reset_mode, self._pool_slots, self._pool = reset_mode, int(pool_slots), None
        self.points_per_lane_request, self.env_index_base = int(points_per_lane), int(env_index_base)
        self.sensor_stride = int(sensor_stride)
        if not (1 <= self.sensor_stride <= self.N):
            raise ValueError("sensor_stride must be in [1, N]")
        self.obs_len = len(range(self.sensor_stride // 2, self.N, self.sensor_stride))

        self.dx = self.L / self.N                                                         # :55
        self.x = np.linspace(0.0, self.L - self.L / self.N, self.N, dtype=np.float32)     # :56
        self.max_episode_steps = math.ceil(self.Tmax / (self.dt * self.cfg_steps))        # :57
        self.burnin_periods = int(BURNIN_TIME / self.dt / self.cfg_steps) if burnin_periods is None \
            else int(burnin_periods)                                                      # :103
        self.forcing = GaussianForcing(self.x, self.Xi, self.sigma, self.L, self.N)       # :60
        self.J = self.forcing.J
        self.noop = np.zeros((1, self.J), dtype=np.float32)                               # :62

        single_action = Box(-1.0, 1.0, shape=(1, self.J), dtype=np.float32)               # :75
        single_obs = Box(-np.inf, np.inf, shape=(1, self.obs_len), dtype=np.float32)      # :76
        super().__init__(num_envs, single_obs, single_action)

        # ---- the CUDA side: fails loudly without the built library / a B200 ----
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("KSVecEnv needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", torch.cuda.current_device()))
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self._F_host = self.forcing.matrix()
        c = _lib.KsConfig(
            abi_version=_lib.KS_ABI_VERSION, num_envs=num_envs, N=self.N, J=self.J, cfg_steps=self.cfg_steps,
            max_episode_steps=self.max_episode_steps, burnin_periods=self.burnin_periods,
            precision=_lib.PRECISIONS[precision], reward_mode=_lib.REWARD_MODES[reward_mode],
            device=self.device_index, points_per_lane=points_per_lane, obs_stride=self.sensor_stride,
            solver=_lib.SOLVERS[solver], dealias=int(self.dealias), env_index_base=self.env_index_base, reserved0=0,
            L=self.L, dt=self.dt,
            forcing=self._F_host.ctypes.data)
        handle = ctypes.c_void_p()
        torch.cuda.init()
        _lib.check(None, lib.ks_create(ctypes.byref(c), ctypes.byref(handle)))
        self._lib, self._h = lib, handle

        offs = _lib._SIZE5()
        total = ctypes.c_size_t()
        _lib.check(self._h, lib.ks_out_layout(self._h, ctypes.byref(offs), ctypes.byref(total)))
        B, N = num_envs, self.N
        self._out_offsets = [int(x) for x in offs]
        self._out_total = int(total.value)
        # pinned result blocks (see _free_block); three up front, because a loop that keeps the last
        # result alive needs two and allocating pinned memory later costs milliseconds
        self._blocks = [self._new_block() for _ in range(3 if copy else 1)]
        self._spill = None
        self._out_pinned = self._blocks[0]["pinned"]
        self._act_pinned = torch.empty((B, self.J), dtype=torch.float32, pin_memory=True)
        self._h_act = self._act_pinned.numpy()
        # device-side outputs of the tensor API (allocated on first use)
        self._d_out = {}
        self._gather = None
        self._pending_actions = None
        self.h2d_bytes_per_step = B * self.J * 4
        self.d2h_bytes_per_step = int(total.value)

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_open(self):
        if self._h is None:
            raise RuntimeError("KSVecEnv is closed")

    @property
    def launch_count(self) -> int:
        """Number of CUDA kernels launched through this env's handle so far."""
        return int(self._lib.ks_launch_count(self._h))

    def launch_info(self) -> dict:
        v = [ctypes.c_int32() for _ in range(5)]
        _lib.check(self._h, self._lib.ks_launch_info(self._h, *[ctypes.byref(x) for x in v]))
        keys = ("points_per_lane", "lanes_per_env", "block_threads", "grid_blocks", "regs_per_thread")
        return {k: x.value for k, x in zip(keys, v)}

    @property
    def time(self) -> np.ndarray:
        """Physical time per env, ``timestep * cfg_steps * dt`` (kuramoto.py:131-133)."""
        return self.get_state()[1] * self.cfg_steps * self.dt

    @property
    def scenario(self) -> dict:
        """kuramoto.py:135-150, including its hard-coded ``noise`` / ``lmbda`` entries."""
        return {"cfg_steps": self.cfg_steps, "Ttrans": self.Ttrans, "L": self.L, "N": self.N, "dx": self.dx,
                "Tmax": self.Tmax, "dt": self.dt, "Xi": self.Xi, "noise": 0.1, "lmbda": 1.0,
                "objective": self.objective}

    # ------------------------------------------------------------------ state access
    def set_state(self, u, timestep=None) -> None:
        """``env.u = u; env.timestep = t`` for every env.  ``u``: ``[B,N]`` float64 array / tensor."""
        self._check_open()
        B, N = self.num_envs, self.N
        ts_arr = None
        if timestep is not None:
            ts_arr = np.ascontiguousarray(np.broadcast_to(np.asarray(timestep, dtype=np.int32), (B,)))
        if isinstance(u, torch.Tensor) and u.is_cuda:
            ud = u.to(device=self.device, dtype=torch.float64).reshape(B, N).contiguous()
            _lib.check(self._h, self._lib.ks_set_state(self._h, _ptr(ud), None, _lib.KS_DEVICE, self._stream()))
            if ts_arr is not None:
                _lib.check(self._h, self._lib.ks_set_state(self._h, None, ts_arr.ctypes.data, _lib.KS_HOST,
                                                           self._stream()))
            return
        ua = None
        if u is not None:
            ua = np.ascontiguousarray(np.asarray(u, dtype=np.float64).reshape(B, N))
        _lib.check(self._h, self._lib.ks_set_state(
            self._h, None if ua is None else ua.ctypes.data, None if ts_arr is None else ts_arr.ctypes.data,
            _lib.KS_HOST, self._stream()))

    def get_state(self):
        """``(u [B,N] float64, timestep [B] int32)`` as NumPy arrays (synchronises)."""
        self._check_open()
        u = np.empty((self.num_envs, self.N), dtype=np.float64)
        ts = np.empty(self.num_envs, dtype=np.int32)
        _lib.check(self._h, self._lib.ks_get_state(self._h, u.ctypes.data, ts.ctypes.data, _lib.KS_HOST,
                                                   self._stream()))
        return u, ts

    def get_state_device(self):
        """``(u [B,N] float64, timestep [B] int32)`` as CUDA tensors (stream-ordered, no sync)."""
        self._check_open()
        u = torch.empty((self.num_envs, self.N), dtype=torch.float64, device=self.device)
        ts = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        _lib.check(self._h, self._lib.ks_get_state(self._h, _ptr(u), _ptr(ts), _lib.KS_DEVICE, self._stream()))
        return u, ts

    # ------------------------------------------------------------------ reset
    def initial_conditions(self, seed: Optional[int], mask: Optional[np.ndarray] = None) -> np.ndarray:
        """Host draws of ``reset``'s initial condition: env ``i`` uses the legacy MT19937 stream of
        ``np.random.seed(seed + i)`` (gym 0.25.2 seeds sub-env ``i`` with ``seed + i``), then
        ``uniform(-0.4, 0.4, N)`` (kuramoto.py:101,106)."""
        # One MT19937 key schedule per env is what the seeded-parity contract costs (init_genrand + 624-word
        # twist: ~6 us per env, 0.4 s at 65 536 envs); ``reset(seed=None)`` / ``ic="device"`` / auto-resets
        # never come here.  The generator object is re-seeded in place instead of rebuilt per env.
        u0 = np.zeros((self.num_envs, self.N), dtype=np.float64)
        rs = np.random.RandomState(0)
        rows = range(self.num_envs) if mask is None else np.nonzero(np.asarray(mask))[0]
        for i in rows:
            rs.seed(None if seed is None else seed + int(i))
            u0[i] = rs.uniform(-IC_AMPLITUDE, IC_AMPLITUDE, size=self.N)
        return u0

    def _reset_impl(self, seed, u0, mask, burnin_periods):
        mask_arr = None
        if mask is not None:
            mask_arr = np.ascontiguousarray(np.asarray(mask, dtype=bool).astype(np.uint8))
        K = -1 if burnin_periods is None else int(burnin_periods)
        if u0 is None and self.ic == "numpy" and seed is not None:
            u0 = self.initial_conditions(seed, mask_arr)
        if u0 is not None:
            u0 = np.ascontiguousarray(np.asarray(u0, dtype=np.float64).reshape(self.num_envs, self.N))
            _lib.check(self._h, self._lib.ks_reset(
                self._h, u0.ctypes.data, None if mask_arr is None else mask_arr.ctypes.data, _lib.KS_HOST, 0, K,
                self._stream()))
        else:
            dev_seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & (2 ** 64 - 1)
            _lib.check(self._h, self._lib.ks_reset(
                self._h, None, None if mask_arr is None else mask_arr.ctypes.data, _lib.KS_HOST, dev_seed, K,
                self._stream()))

    def _auto_reset(self, mask) -> None:
        """gym's per-sub-env auto-reset: burn-in launch (reference semantics) or pooled states."""
        if self.reset_mode == "burnin":
            self._reset_impl(None, None, mask, None)
            return
        if self._pool is None:
            from .reset_pool import ResetPool

            self._pool = ResetPool(self, slots=self._pool_slots)
        batch, slot = self._pool.take()
        m = None
        if mask is not None:
            m = torch.as_tensor(np.asarray(mask, dtype=np.uint8)).to(self.device)
        _lib.check(self._h, self._lib.ks_reset(self._h, _ptr(batch), _ptr(m), _lib.KS_DEVICE, 0, 0, self._stream()))
        self._pool.release(slot)

    def reset(self, seed: Optional[int] = None, return_info: bool = False, options: Optional[dict] = None,
              *, u0=None, burnin_periods: Optional[int] = None, **kwargs):
        """``KuramotoSivashinskyEnv.reset`` for every env (kuramoto.py:100-116): initial condition,
        ``burnin_periods`` no-op control periods in one launch, ``timestep = 0``.

        Returns ``obs (B,1,N) float32`` (and ``{"step": zeros}`` when ``return_info``).  ``u0``
        injects initial conditions (``[B,N]`` float64) instead of drawing them.
        """
        self._check_open()
        self._pending_actions = None
        self._reset_impl(seed, u0, None, burnin_periods)
        u, ts = self.get_state()
        self._raise_if_nonfinite(u)
        obs = self._observe(u)
        if return_info:
            return obs, {"step": ts.astype(np.int64), "_step": np.ones(self.num_envs, dtype=bool)}
        return obs

    # ------------------------------------------------------------------ step (host / NumPy API)
    def step_async(self, actions) -> None:
        """Stage actions ``(B,1,J)`` / ``(B,J)`` as float32 (``np.array(action, float32)``, :79)."""
        self._check_open()
        a = np.asarray(actions, dtype=np.float32)
        if a.size != self.num_envs * self.J:
            raise ValueError(f"actions have shape {a.shape}, expected ({self.num_envs}, 1, {self.J})")
        np.copyto(self._h_act, a.reshape(self.num_envs, self.J))
        self._pending_actions = True

    def step_wait(self, **kwargs):
        """One control period for every env; H2D of the actions, one kernel, one D2H of the packed
        outputs.  Envs whose episode ends are reset (800-period burn-in) before returning, with the
        pre-reset observation in ``infos["final_observation"]`` as gym 0.25.2 does."""
        self._check_open()
        if not self._pending_actions:
            raise RuntimeError("step_wait() called without step_async()")
        self._pending_actions = None
        blk = self._free_block()
        _lib.check(self._h, self._lib.ks_step_host(self._h, self._h_act.ctypes.data, blk["pinned"].data_ptr(),
                                                   self._stream()))
        if blk["bad"].any():
            bad = np.nonzero(blk["bad"])[0]
            raise FloatingPointError(f"overflow encountered in KS state of env(s) {bad[:8].tolist()}"
                                     f"{'...' if bad.size > 8 else ''} (np.seterr(over='raise') in the reference)")
        if self.copy and blk["owned"]:
            # zero-copy: fresh view objects into a pinned block that no earlier result still uses
            obs, rewards = blk["obs"][...], blk["reward"][...]
        elif self.copy:
            obs, rewards = np.array(blk["obs"]), np.array(blk["reward"])
        else:
            obs, rewards = blk["obs"], blk["reward"]
        steps = blk["step"].astype(np.int64)
        truncated = blk["trunc"].astype(bool)
        terminated = np.zeros(self.num_envs, dtype=bool)
        infos = {"step": steps, "_step": np.ones(self.num_envs, dtype=bool)}
        if truncated.any():
            # gym 0.25.2 vector-env auto-reset: final obs is the single env's float64 (1,N) array
            # (the float64 state has to be fetched for that: the packed block carries float32 observations)
            u, _ = self.get_state()
            finals_block = np.ascontiguousarray(u[:, self.sensor_stride // 2::self.sensor_stride]).reshape(
                self.num_envs, 1, self.obs_len)
            finals = np.fromiter(iter(finals_block), dtype=object, count=self.num_envs)   # (1,No) views of ONE block
            idx = np.nonzero(truncated)[0]
            if idx.size != self.num_envs:
                finals[~truncated] = None
            infos["final_observation"] = finals
            infos["_final_observation"] = truncated.copy()
            self._auto_reset(None if truncated.all() else truncated)
            # post-reset observations: float32 on the device, only the reset rows cross PCIe
            u_dev, _ = self.get_state_device()
            o_dev = u_dev[:, self.sensor_stride // 2::self.sensor_stride].to(torch.float32)
            obs = np.array(obs)
            if truncated.all():
                new = o_dev.cpu().numpy()
                self._raise_if_nonfinite(new)
                obs[:, 0] = new
            else:
                new = o_dev[torch.as_tensor(idx, device=self.device)].cpu().numpy()
                self._raise_if_nonfinite(new)
                obs[idx, 0] = new
        return obs, rewards, terminated, truncated, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    MAX_RESULT_BLOCKS = 8

    def _new_block(self) -> dict:
        """One pinned host block with the layout of ``ks_out_layout`` and typed views into it."""
        B, o = self.num_envs, self._out_offsets
        pinned = torch.empty(self._out_total, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        host = pinned.numpy()
        blk = dict(pinned=pinned, host=host, owned=True,
                   reward=host[o[0]:o[0] + 8 * B].view(np.float64),
                   obs=host[o[1]:o[1] + 4 * B * self.obs_len].view(np.float32).reshape(B, 1, self.obs_len),
                   step=host[o[2]:o[2] + 4 * B].view(np.int32), trunc=host[o[3]:o[3] + B].view(np.uint8),
                   bad=host[o[4]:o[4] + B].view(np.uint8))
        del host, pinned
        blk["refs"] = self._block_refs(blk)     # our own references; anything above = a result still in use
        return blk

    @staticmethod
    def _block_refs(blk: dict) -> int:
        return sys.getrefcount(blk["host"])

    def _free_block(self) -> dict:
        """The pinned block the next ``ks_step_host`` may DMA into.  With ``copy=True`` the arrays
        ``step`` returns are views of such a block (reading 1 MiB back out of freshly DMA-written
        pinned memory costs a single core ~70 us, more than the whole D2H copy), so a block is reused
        only when no earlier result -- or any view derived from one -- still references it (NumPy's
        base-object reference count says so).  A caller that keeps every observation alive (a replay
        buffer of views) ends up on the last block with ``owned = False``: results are then copied
        out, exactly like ``np.array``.  Either way a returned array is never overwritten."""
        if not self.copy:
            return self._blocks[0]
        for blk in self._blocks:
            if self._block_refs(blk) == blk["refs"]:
                blk["owned"] = True
                return blk
        if len(self._blocks) < self.MAX_RESULT_BLOCKS:
            blk = self._new_block()
            self._blocks.append(blk)
            return blk
        if self._spill is None:
            self._spill = self._new_block()
        self._spill["owned"] = False
        return self._spill

    def _observe(self, u: np.ndarray) -> np.ndarray:
        """float32 observation ``(B,1,No)`` of states ``u [B,N]`` (cast + sensor sampling)."""
        return u[:, self.sensor_stride // 2::self.sensor_stride].astype(np.float32).reshape(-1, 1, self.obs_len)

    def _raise_if_nonfinite(self, u):
        if not np.all(np.isfinite(u)):
            raise FloatingPointError("overflow encountered in KS state")

    # ------------------------------------------------------------------ device-tensor API
    def _device_outputs(self, K: int):
        """Output tensors of the device API.  For a single period (K = 0) all five are typed views
        into ONE contiguous byte block with the layout of ``ks_out_layout`` -- a sharded run
        all-gathers that block with a single collective (``sharding.gather_packed``)."""
        out = self._d_out.get(K)
        if out is None:
            # The single-period block (K = 0) lives as long as the env: CUDA graphs captured by
            # ``DeviceEnvPipeline.rollout_graphed`` and the fused gather have its raw pointers baked in.
            # Rollout buffers are kept for the most recent K only.
            B, dev = self.num_envs, self.device
            if K == 0:
                offs, total = self._out_offsets, self.d2h_bytes_per_step
                block = torch.zeros(total, dtype=torch.uint8, device=dev)

                def view(i, nbytes, dtype, shape):
                    return block[offs[i]:offs[i] + nbytes].view(dtype).reshape(shape)

                out = dict(reward=view(0, 8 * B, torch.float64, (B,)),
                           obs=view(1, 4 * B * self.obs_len, torch.float32, (B, self.obs_len)),
                           step=view(2, 4 * B, torch.int32, (B,)),
                           truncated=view(3, B, torch.uint8, (B,)),
                           nonfinite=view(4, B, torch.uint8, (B,)), packed=block)
            else:
                lead = (K, B)
                out = dict(obs=torch.empty(lead + (self.obs_len,), dtype=torch.float32, device=dev),
                           reward=torch.empty(lead, dtype=torch.float64, device=dev),
                           truncated=torch.empty(lead, dtype=torch.uint8, device=dev),
                           step=torch.empty(lead, dtype=torch.int32, device=dev),
                           nonfinite=torch.empty(lead, dtype=torch.uint8, device=dev))
            for old in [k for k in self._d_out if k != 0]:
                del self._d_out[old]
            self._d_out[K] = out
        return out

    def packed_fields(self) -> dict:
        """``{name: (byte offset, torch dtype, per-env shape)}`` of the packed single-period block."""
        o = self._out_offsets
        return {"reward": (o[0], torch.float64, ()), "obs": (o[1], torch.float32, (self.obs_len,)),
                "step": (o[2], torch.int32, ()), "truncated": (o[3], torch.uint8, ()),
                "nonfinite": (o[4], torch.uint8, ())}

    def step_device(self, actions: torch.Tensor, phi: Optional[torch.Tensor] = None) -> dict:
        """One control period with device-resident inputs and outputs; asynchronous on the current
        stream, no host synchronisation, no auto-reset.  ``actions``: CUDA float32 ``[B,J]`` (or
        ``[B,1,J]``).  ``phi`` (CUDA float32 ``[B,N]``) overrides the in-kernel ``a @ F``.
        Returns a dict of CUDA tensors that are REUSED by the next call:
        ``obs [B,No] f32, reward [B] f64, truncated [B] u8, step [B] i32, nonfinite [B] u8`` (views
        of one contiguous block, also returned as ``packed``)."""
        self._check_open()
        out = self._device_outputs(0)
        a = None
        if actions is not None:
            a = actions.to(device=self.device, dtype=torch.float32).reshape(self.num_envs, self.J).contiguous()
        p = None
        if phi is not None:
            p = phi.to(device=self.device, dtype=torch.float32).reshape(self.num_envs, self.N).contiguous()
        _lib.check(self._h, self._lib.ks_step(self._h, _ptr(a), _ptr(p), _ptr(out["obs"]), _ptr(out["reward"]),
                                              _ptr(out["truncated"]), _ptr(out["step"]), _ptr(out["nonfinite"]),
                                              self._stream()))
        return out

    # ------------------------------------------------------------------ fused wrapper plumbing
    def collect_step(self, *, actions, out, obs_store, act_store, vminmax, agent_obs, rec, lower: float,
                     upper: float, frozen: bool, agent_stride: int, slot_index=None) -> None:
        """``ks_collect``: the per-step bookkeeping of the reference's wrapper stack + ``Worker.rollout``
        (stores of length 1) in two small kernels instead of ~20 tensor ops.  ``out`` = the dict
        ``step_device`` returned, ``rec`` = ``(obs, actions, nxtobs, reward, truncated, step)`` slot views
        of the transition buffers -- or, with ``slot_index`` (a CUDA int64 counter), the whole ``[T,B,...]``
        buffers, of which slot ``t`` is written and ``t`` incremented on the device (graph-capturable);
        everything contiguous CUDA tensors."""
        r_obs, r_act, r_nxt, r_rew, r_trunc, r_step = rec
        a = _lib.KsCollectArgs(
            actions=actions.data_ptr(), obs=out["obs"].data_ptr(), reward=out["reward"].data_ptr(),
            truncated=out["truncated"].data_ptr(), step=out["step"].data_ptr(), obs_store=obs_store.data_ptr(),
            act_store=act_store.data_ptr(), vminmax=vminmax.data_ptr(), agent_obs=agent_obs.data_ptr(),
            rec_obs=r_obs.data_ptr(), rec_actions=r_act.data_ptr(), rec_nxtobs=r_nxt.data_ptr(),
            rec_reward=r_rew.data_ptr(), rec_truncated=r_trunc.data_ptr(), rec_step=r_step.data_ptr(),
            lower=float(lower), scale_width=float(upper - lower), frozen=int(bool(frozen)), agent_stride=int(agent_stride),
            slot_index=None if slot_index is None else slot_index.data_ptr())
        _lib.check(self._h, self._lib.ks_collect(self._h, ctypes.byref(a), self._stream()))

    # ------------------------------------------------------------------ fused all-gather (multi-GPU)
    def gather_init(self, world: int, rank: int) -> bytes:
        """Allocate this rank's gather buffer (``ks_gather_init``); returns the 64-byte CUDA-IPC
        handle that the other ranks need (exchange it with any host-side collective)."""
        self._check_open()
        handle = ctypes.create_string_buffer(_lib.KS_IPC_HANDLE_BYTES)
        slot = ctypes.c_size_t()
        _lib.check(self._h, self._lib.ks_gather_init(self._h, world, rank, handle, ctypes.byref(slot)))
        self._gather = dict(world=world, rank=rank, slot=int(slot.value), views={})
        return handle.raw

    def gather_connect(self, handles: Sequence[bytes]) -> None:
        """Map every peer's gather buffer into this process (``ks_gather_connect``); ``handles`` in rank order."""
        self._check_open()
        blob = b"".join(bytes(h) for h in handles)
        if len(blob) != self._gather["world"] * _lib.KS_IPC_HANDLE_BYTES:
            raise ValueError("need one 64-byte IPC handle per rank, in rank order")
        _lib.check(self._h, self._lib.ks_gather_connect(self._h, blob))

    def gather_layout(self, world: int) -> tuple[int, int]:
        """``(slot_bytes, total_bytes)`` every rank's gather buffer needs for a world of ``world`` ranks."""
        slot, total = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(self._h, self._lib.ks_gather_layout(self._h, world, ctypes.byref(slot), ctypes.byref(total)))
        return int(slot.value), int(total.value)

    def gather_attach(self, world: int, rank: int, peer_ptrs: Sequence[int], multicast_ptr: int = 0, nbytes: int = 0) -> None:
        """Run the fused exchange on buffers the caller allocated and mapped (``ks_gather_attach``): symmetric
        memory instead of CUDA IPC.  ``peer_ptrs[r]`` = rank r's buffer as mapped into this process;
        ``multicast_ptr`` (0 = none) = an NVLS multicast address bound to all of them, which lets the period
        kernel send its observation rows once (``multimem.st``) instead of once per peer."""
        self._check_open()
        arr = (ctypes.c_void_p * world)(*[int(p) for p in peer_ptrs])
        _lib.check(self._h, self._lib.ks_gather_attach(self._h, world, rank, arr, ctypes.c_void_p(int(multicast_ptr) or None),
                                                       int(nbytes)))
        self._gather = dict(world=world, rank=rank, slot=self.gather_layout(world)[0], views={})

    def step_gather(self, actions: torch.Tensor) -> dict:
        """One control period of the local shard whose kernel epilogue also stores the outputs into
        every peer's gather buffer over NVLink (``ks_step_gather``).  Returns full-batch views
        ``{name: [world, local_envs, ...]}`` (rank order = env order) into the double-buffered gather
        block: valid until the second next call.  Stream-ordered, no host synchronisation."""
        self._check_open()
        if self._gather is None:
            raise RuntimeError("step_gather() before gather_init() / gather_connect()")
        a = actions.to(device=self.device, dtype=torch.float32).reshape(self.num_envs, self.J).contiguous()
        ptr = ctypes.c_void_p()
        # fails (KsError, KS_ERR_STATE) once an earlier exchange timed out: never hands out incomplete blocks
        _lib.check(self._h, self._lib.ks_step_gather(self._h, _ptr(a), ctypes.byref(ptr), self._stream()))
        g = self._gather
        views = g["views"].get(ptr.value)
        if views is None:
            world, slot, B = g["world"], g["slot"], self.num_envs

            class _Raw:     # zero-copy torch view of library-owned device memory
                __cuda_array_interface__ = {"shape": (world * slot,), "typestr": "|u1", "data": (ptr.value, False),
                                            "version": 3, "strides": None}

            rows = torch.as_tensor(_Raw(), device=self.device).view(world, slot)
            views = {"packed": rows}
            for name, (off, dtype, shape) in self.packed_fields().items():
                n = B
                for d in shape:
                    n *= d
                nbytes = n * torch.empty((), dtype=dtype).element_size()
                views[name] = rows[:, off:off + nbytes].view(dtype).reshape((world, B) + tuple(shape))
            g["views"][ptr.value] = views
        return views

    def gather_timed_out(self) -> bool:
        """True once a peer failed to signal within the handshake's bound (synchronises the stream)."""
        flag = ctypes.c_int32()
        _lib.check(self._h, self._lib.ks_gather_status(self._h, ctypes.byref(flag), self._stream()))
        return bool(flag.value)

    def gather_clear(self) -> None:
        """Re-arm the fused exchange after a time-out (the application has re-synchronised its ranks)."""
        _lib.check(self._h, self._lib.ks_gather_clear(self._h, self._stream()))

    def gather_barrier(self) -> None:
        """Stream-ordered device-side rendezvous of all ranks (``ks_gather_barrier``; no host synchronisation)."""
        _lib.check(self._h, self._lib.ks_gather_barrier(self._h, self._stream()))

    def rollout_device(self, actions: Optional[torch.Tensor], K: Optional[int] = None, outputs: bool = True) -> dict:
        """``K`` control periods in ONE persistent launch (open loop).  ``actions``: CUDA float32
        ``[K,B,J]`` or ``None`` for no-op periods (then ``K`` is required)."""
        self._check_open()
        a = None
        if actions is not None:
            a = actions.to(device=self.device, dtype=torch.float32).reshape(-1, self.num_envs, self.J).contiguous()
            K = int(a.shape[0])
        if not K or K < 1:
            raise ValueError("K >= 1 required")
        out = self._device_outputs(K) if outputs else dict(obs=None, reward=None, truncated=None, step=None,
                                                           nonfinite=None)
        _lib.check(self._h, self._lib.ks_rollout(self._h, K, _ptr(a), _ptr(out["obs"]), _ptr(out["reward"]),
                                                 _ptr(out["truncated"]), _ptr(out["step"]), _ptr(out["nonfinite"]),
                                                 self._stream()))
        return out

    def reset_device(self, seed: Optional[int] = None, mask: Optional[torch.Tensor] = None,
                     burnin_periods: Optional[int] = None) -> None:
        """Stream-ordered reset with device-generated initial conditions (no host sync)."""
        self._check_open()
        dev_seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed) & (2 ** 64 - 1)
        m = None if mask is None else mask.to(device=self.device, dtype=torch.uint8).contiguous()
        K = -1 if burnin_periods is None else int(burnin_periods)
        _lib.check(self._h, self._lib.ks_reset(self._h, None, _ptr(m), _lib.KS_DEVICE, dev_seed, K, self._stream()))

    def nonfinite(self) -> np.ndarray:
        """Sticky per-env blow-up flags (the reference raises ``FloatingPointError`` instead)."""
        self._check_open()
        flags = np.empty(self.num_envs, dtype=np.uint8)
        _lib.check(self._h, self._lib.ks_status(self._h, flags.ctypes.data, None, self._stream()))
        return flags.astype(bool)

    # ------------------------------------------------------------------ env.rhs / env.reward_func
    def evaluate(self, u, phi=None, want=("rhs", "ux", "uxx", "uxxxx", "reward")) -> dict:
        """Batched ``rhs`` / derivative triple / reward of arbitrary states on the GPU
        (kuramoto.py:118-129, 64-70).  ``u [M,N]`` float64, ``phi [M,N]`` float32 or None."""
        self._check_open()
        is_np = not isinstance(u, torch.Tensor)
        ud = torch.as_tensor(np.asarray(u, dtype=np.float64) if is_np else u).to(self.device, torch.float64)
        ud = ud.reshape(-1, self.N).contiguous()
        M = ud.shape[0]
        pd = None
        if phi is not None:
            pd = torch.as_tensor(np.asarray(phi, dtype=np.float32) if not isinstance(phi, torch.Tensor) else phi)
            pd = pd.to(self.device, torch.float32).broadcast_to(ud.shape).contiguous()
        bufs = {k: torch.empty((M, self.N) if k != "reward" else (M,), dtype=torch.float64, device=self.device)
                for k in want}
        _lib.check(self._h, self._lib.ks_eval(self._h, M, _ptr(ud), _ptr(pd), _ptr(bufs.get("rhs")),
                                              _ptr(bufs.get("ux")), _ptr(bufs.get("uxx")), _ptr(bufs.get("uxxxx")),
                                              _ptr(bufs.get("reward")), self._stream()))
        if is_np:
            return {k: v.cpu().numpy() for k, v in bufs.items()}
        return bufs

    def rhs(self, u, phi):
        """``env.rhs(u, phi)`` -> ``(rhs, (ux, uxx, uxxxx))`` for one state or a batch (GPU)."""
        shape = np.shape(u) if not isinstance(u, torch.Tensor) else tuple(u.shape)
        r = self.evaluate(u, phi, want=("rhs", "ux", "uxx", "uxxxx"))
        rs = lambda a: a.reshape(shape)  # noqa: E731
        return rs(r["rhs"]), (rs(r["ux"]), rs(r["uxx"]), rs(r["uxxxx"]))

    def reward_func(self, obs, phi=None, *args, **kwargs):
        """``env.reward_func(obs, phi)`` (FuncTransform over ``l2control`` / ``dissipation``,
        kuramoto.py:64-73) for one observation ``(1,N)`` / ``(N,)`` -> scalar, or a batch
        ``[M,(1,)N]`` -> ``[M]``; evaluated by the CUDA ``ks_eval`` kernel in float64.  Like the reference's
        ``FuncTransform`` the result comes back in the caller's container: NumPy in -> NumPy out, tensor in ->
        tensor on the SAME device (``surrogates/training.py:214`` stacks CPU tensors and calls ``.numpy()``),
        floating inputs keep their dtype."""
        is_np = not isinstance(obs, torch.Tensor)
        o = np.asarray(obs) if is_np else obs
        in_dtype = o.dtype
        single = o.size == self.N if is_np else o.numel() == self.N
        if self.reward_mode == "dissipation" and phi is None:
            raise TypeError("dissipation reward needs phi")
        if self.reward_mode == "l2":
            phi = None          # the L2 objective ignores its second argument (callers pass phi or even the action there)
        if is_np:
            u = np.asarray(o, dtype=np.float64).reshape(-1, self.N)
            p = None if phi is None else np.asarray(phi.detach().cpu() if isinstance(phi, torch.Tensor) else phi,
                                                    dtype=np.float32).reshape(-1, self.N)
            r = self.evaluate(u, p, want=("reward",))["reward"]
            if np.issubdtype(in_dtype, np.floating):
                r = r.astype(in_dtype, copy=False)
            return r[0] if single else r
        u = o.to(device=self.device, dtype=torch.float64).reshape(-1, self.N)
        p = None if phi is None else torch.as_tensor(phi).to(device=self.device, dtype=torch.float32).reshape(-1, self.N)
        r = self.evaluate(u, p, want=("reward",))["reward"].to(device=o.device)
        if in_dtype.is_floating_point:
            r = r.to(in_dtype)
        return r[0] if single else r

    # ------------------------------------------------------------------ teardown
    def close_extras(self, **kwargs):
        if getattr(self, "_pool", None) is not None:
            self._pool.close()
            self._pool = None
        if getattr(self, "_h", None) is not None:
            self._lib.ks_destroy(self._h)
            self._h = None

    def close(self, **kwargs):
        self.close_extras(**kwargs)
        self.closed = True

    def __del__(self):
        try:
            self.close_extras()
        except Exception:
            pass
