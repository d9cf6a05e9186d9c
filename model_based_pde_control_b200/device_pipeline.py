"""Device-side vector-env plumbing: the reference's wrapper stack re-implemented on torch tensors so
that observations never leave the GPU between the env kernel and a GPU-resident policy.

What is mirrored (``pdegym/common/vec_wrappers.py``, ``pdegym/common/transforms.py``,
``pdecontrol/mbrl/mbrl.py:257-275``, ``pdecontrol/mbrl/worker.py:39-93``)::

    env -> StoreNObsVecWrapper(num_steps) -> TransformObsWrapper(ScaleTransform, running min/max)
        -> TransformObsWrapper(SensorTransform) -> StoreNActionsVecWrapper(num_steps)
        -> TransformActionWrapper(ScaleTransform(bounds).Inverse, frozen)
    Worker.rollout: (obs, actions, nxtobs, rewards, terminated, truncated, steps) per step,
                    nxtobs replaced by the stored final observation where an episode ended.

The classes below are backend-agnostic torch code (they run on CPU tensors too, which is how they
are tested against the reference's own wrappers, ``tests/test_device_pipeline.py``); the env
object only has to provide ``step_device`` / ``reset_device`` / ``get_state_device`` like
``KSVecEnv``.  Episode boundaries are tracked on the host from the step count (all envs of a batch
run synchronous fixed-length episodes, ``vec_wrappers.py:26-30`` relies on that too), so a rollout
issues no device->host synchronisation at all.
"""
from __future__ import annotations

from typing import Callable, NamedTuple, Optional

import torch


class ScaleTransformDevice:
    """``ScaleTransform(scale, bounds, aggregate=True, batched=True)`` (transforms.py:141-210) with
    scalar running min / max kept as 0-dim tensors on the data's device.  ``update`` follows
    ``ScaleTransform.update`` (min/max over everything seen so far), ``__call__`` maps
    ``[vmin, vmax] -> [lower, upper]`` in float32, ``inverse`` maps back."""

    def __init__(self, scale=(-1.0, 1.0), bounds=(None, None), frozen: bool = False):
        self.lower, self.upper = float(scale[0]), float(scale[1])
        self.vmin = None if bounds[0] is None else torch.as_tensor(bounds[0], dtype=torch.float32).amin()
        self.vmax = None if bounds[1] is None else torch.as_tensor(bounds[1], dtype=torch.float32).amax()
        self.frozen = frozen
        self.minmax = None

    def _on(self, device):
        """Keep the running bounds on the data's device (a per-call ``.to`` of a CPU scalar is a
        synchronous pageable copy)."""
        if self.vmin is not None and self.vmin.device != device:
            self.vmin = self.vmin.to(device)
        if self.vmax is not None and self.vmax.device != device:
            self.vmax = self.vmax.to(device)

    def update(self, values: torch.Tensor) -> None:
        if self.frozen:
            return
        self._on(values.device)
        lo, hi = torch.aminmax(values)
        lo, hi = lo.to(torch.float32), hi.to(torch.float32)
        # in place once the bounds exist: a captured CUDA graph keeps pointing at these tensors
        if self.vmin is None:
            self.minmax = torch.stack([lo, hi])            # one 2-element tensor: ks_collect updates it in place
            self.vmin, self.vmax = self.minmax[0], self.minmax[1]
        else:
            torch.minimum(lo, self.vmin, out=self.vmin)
            torch.maximum(hi, self.vmax, out=self.vmax)

    def __call__(self, values: torch.Tensor) -> torch.Tensor:
        self._on(values.device)
        return (values - self.vmin) / (self.vmax - self.vmin) * (self.upper - self.lower) + self.lower

    def inverse(self, values: torch.Tensor) -> torch.Tensor:
        self._on(values.device)
        return (values - self.lower) / (self.upper - self.lower) * (self.vmax - self.vmin) + self.vmin


class SensorTransformDevice:
    """``SensorTransform(stride)`` (transforms.py:231-247): ``values[..., stride//2::stride]``."""

    def __init__(self, stride: int = 1):
        self.stride = int(stride)

    def __call__(self, values: torch.Tensor) -> torch.Tensor:
        return values[..., int(self.stride / 2)::self.stride]


class RolloutBatch(NamedTuple):
    """Field order of ``pdecontrol/mbrl/types.py:9-17`` ``Sample``; leading dims ``[T, B]``."""

    obs: torch.Tensor          # [T,B,S,1,N]  stored (raw) observations before the step
    actions: torch.Tensor      # [T,B,S,1,J]  env-scale actions
    nxtobs: torch.Tensor       # [T,B,S,1,N]  stored observations after the step (final obs where done)
    rewards: torch.Tensor      # [T,B]
    terminated: torch.Tensor   # [T,B] bool
    truncated: torch.Tensor    # [T,B] bool
    steps: torch.Tensor        # [T,B] int


class DeviceEnvPipeline:
    """The data-collection stack of ``setup_wrapped_envs`` (mbrl.py:257-275) on device tensors.

    ``env``            object with ``num_envs, N (or obs_len), J, max_episode_steps, step_device(actions),
                       reset_device(seed=...), get_state_device()`` (``KSVecEnv``)
    ``num_steps``      history length of the obs / action stores (1 in the reference's MBRL loop)
    ``obs_scale``      target range of the running min/max observation scaling
    ``fused``          let ``rollout`` use ``ks_collect`` (two CUDA kernels per step for the whole store /
                       scaling / record bookkeeping) when the env provides it; results are bitwise
                       those of the tensor-op path
    ``action_bounds``  ``(low, high)`` of the env's action space; agent actions in ``[-1,1]`` are mapped
                       onto it (identity for the KS env), frozen like the reference's ``ascaling``
    """

    def __init__(self, env, num_steps: int = 1, obs_scale=(-1.0, 1.0), frozen_obs_scaling: bool = False,
                 action_bounds=(-1.0, 1.0), agent_sensor_stride: int = 1, fused: bool = True):
        self.env = env
        self.B = env.num_envs
        self.num_steps = num_steps
        self.oscaling = ScaleTransformDevice(scale=obs_scale, frozen=frozen_obs_scaling)
        self.ascaling = ScaleTransformDevice(scale=(-1.0, 1.0), bounds=action_bounds, frozen=True)
        self.agent_sensor = SensorTransformDevice(agent_sensor_stride)
        self.obs_store = self.finals = self.act_store = None
        self._obs_row = self._act_row = None
        self.fused = fused                 # use the env library's ks_collect kernels where possible
        self._rollout_buf = {}
        self._graphs = {}
        self._episode_step = None          # host-side step counter (all envs synchronous)
        self._pending_final = None

    # -- StoreNObsVecWrapper -------------------------------------------------------------------
    # All envs of the batch run synchronous fixed-length episodes (the reference relies on that too,
    # vec_wrappers.py:26-30), so the validity masks of the stores have identical rows: they are kept
    # on the HOST as one row each, and every masked read below is a plain slice -- no boolean
    # indexing, hence no device->host synchronisation anywhere in a rollout.
    def _ostore_reset(self, obs):
        self.obs_store = obs.unsqueeze(1).repeat(1, self.num_steps, *([1] * (obs.dim() - 1))).clone()
        self.finals = torch.zeros_like(self.obs_store)
        self._obs_row = [False] * (self.num_steps - 1) + [True]

    def _ostore_step(self, obs, final_obs):
        if final_obs is not None:
            self.finals[:] = final_obs.unsqueeze(1)
            self._obs_row = [False] * self.num_steps
        self.obs_store[:, 0] = obs
        self._obs_row[0] = True
        if self.num_steps > 1:
            self.obs_store = torch.roll(self.obs_store, -1, dims=1)
            self._obs_row = self._obs_row[1:] + self._obs_row[:1]

    @property
    def obs_mask(self) -> torch.Tensor:
        """``StoreNObsVecWrapper.mask`` (B, num_steps) -- materialised on request only."""
        return torch.tensor(self._obs_row, dtype=torch.bool, device=self.obs_store.device).repeat(self.B, 1)

    # -- StoreNActionsVecWrapper ---------------------------------------------------------------
    def _astore_reset(self, like):
        self.act_store = torch.zeros((self.B, self.num_steps, 1, self.env.J), dtype=torch.float32, device=like.device)
        self._act_row = [False] * self.num_steps

    def _astore_step(self, actions):
        self.act_store[:, 0] = actions
        self._act_row[0] = True
        if self.num_steps > 1:
            self.act_store = torch.roll(self.act_store, -1, dims=1)
            self._act_row = self._act_row[1:] + self._act_row[:1]

    @property
    def act_mask(self) -> torch.Tensor:
        return torch.tensor(self._act_row, dtype=torch.bool, device=self.act_store.device).repeat(self.B, 1)

    @staticmethod
    def _valid(store: torch.Tensor, row) -> torch.Tensor:
        """``store[mask].reshape(B, n_valid, ...)`` for a mask whose rows all equal ``row`` (a copy)."""
        idx = [i for i, v in enumerate(row) if v]
        if idx == list(range(idx[0], idx[-1] + 1)):
            return store[:, idx[0]:idx[-1] + 1].clone()
        return store[:, idx]

    # -- the stack -----------------------------------------------------------------------------
    def _env_obs(self):
        u, _ = self.env.get_state_device()
        stride = getattr(self.env, "sensor_stride", 1)
        return u[:, stride // 2::stride].to(torch.float32).unsqueeze(1)      # (B,1,No), gym's float32 buffer

    def reset(self, seed: Optional[int] = None) -> torch.Tensor:
        """``stack.envs.reset()``: env reset (device ICs + burn-in launch), stores re-initialised,
        running scaling updated; returns the scaled agent observation ``(B,1,No')``."""
        self.env.reset_device(seed=seed)
        obs = self._env_obs()
        self._ostore_reset(obs)
        self._astore_reset(obs)
        self._episode_step = 0
        self.oscaling.update(obs)
        return self.agent_sensor(self.oscaling(obs))

    def step(self, agent_actions: torch.Tensor):
        """``stack.envs.step(actions)``: returns ``(scaled obs, rewards, terminated, truncated, infos)``
        as device tensors; ``infos`` has ``step`` and, when the episode ended, ``final_observation`` (scaled,
        like ``TransformObsWrapper`` does when not frozen) and ``_final_observation``."""
        a = self.ascaling.inverse(agent_actions.to(torch.float32).reshape(self.B, 1, self.env.J))   # TransformActionWrapper
        self._astore_step(a)
        out = self.env.step_device(a.reshape(self.B, self.env.J))
        obs = out["obs"].unsqueeze(1).clone()
        rewards = out["reward"].clone()
        steps = out["step"].to(torch.int64)
        truncated = out["truncated"].to(torch.bool).clone()
        terminated = torch.zeros_like(truncated)
        infos = {"step": steps}
        self._episode_step += 1
        final_obs = final_mask = None
        if self._episode_step >= self.env.max_episode_steps:            # known on the host: no sync
            final_obs, final_mask = obs, truncated
            self.env.reset_device(seed=None)                            # gym auto-reset (800-period burn-in)
            obs = self._env_obs()
            self._episode_step = 0
        self._ostore_step(obs, final_obs)
        if final_obs is not None:
            self._act_row = [False] * (self.num_steps - 1) + self._act_row[-1:]    # StoreNActionsVecWrapper.step_wait
        self.oscaling.update(obs)
        scaled = self.oscaling(obs)
        if final_obs is not None:
            if not self.oscaling.frozen:
                self.oscaling.update(final_obs)
                infos["final_observation"] = self.oscaling(final_obs)
            else:
                infos["final_observation"] = final_obs
            infos["_final_observation"] = final_mask
        return self.agent_sensor(scaled), rewards, terminated, truncated, infos

    def rollout(self, select_action: Callable[[torch.Tensor], torch.Tensor], num_steps: int,
                last_obs: Optional[torch.Tensor] = None, seed: Optional[int] = None, reuse_buffers: bool = False):
        """Batched ``Worker.rollout`` (worker.py:39-93): ``num_steps`` env steps with a GPU-resident
        ``select_action(scaled_obs) -> actions in [-1,1]``; returns ``(RolloutBatch, last_obs)`` with all
        tensors on the device.  No host synchronisation inside the loop: the transitions are written
        into ``[T, B, ...]`` buffers allocated up front.  ``reuse_buffers=True`` keeps those buffers
        for the next rollout of the same length (the returned batch is then overwritten by it --
        copy what must survive, e.g. into the replay), which removes every allocation from the loop."""
        if last_obs is None:
            last_obs = self.reset(seed=seed)
        last_stored = self._valid(self.obs_store, self._obs_row)
        S, dev = last_stored.shape[1], last_stored.device
        key = (num_steps, S)
        buf = self._rollout_buf.get(key) if reuse_buffers else None
        if buf is None:
            lead = (num_steps, self.B)
            buf = RolloutBatch(
                obs=torch.empty(lead + tuple(last_stored.shape[1:]), dtype=last_stored.dtype, device=dev),
                actions=torch.empty(lead + (S, 1, self.env.J), dtype=torch.float32, device=dev),
                nxtobs=torch.empty(lead + tuple(last_stored.shape[1:]), dtype=last_stored.dtype, device=dev),
                rewards=torch.empty(lead, dtype=torch.float64, device=dev),
                terminated=torch.zeros(lead, dtype=torch.bool, device=dev),
                truncated=torch.empty(lead, dtype=torch.bool, device=dev),
                steps=torch.empty(lead, dtype=torch.int64, device=dev))
            if reuse_buffers:
                self._rollout_buf = {key: buf}
        # ks_collect (two small CUDA kernels of the env library) does the per-step bookkeeping when the
        # env offers it, the stores have length 1 and the step does not end the episode; everything
        # else goes through step() and tensor copies.  Both paths keep the same in-place state.
        fused = (self.fused and hasattr(self.env, "collect_step") and S == 1 and last_obs.is_cuda
                 and self.oscaling.minmax is not None and self.act_store.is_contiguous() and self.obs_store.is_contiguous())
        agent_buf = torch.empty_like(last_obs) if fused else None
        stale = False                                # last_stored out of date (fused steps do not maintain it)
        for t in range(num_steps):
            with torch.no_grad():
                actions = select_action(last_obs)
            if fused and self._episode_step + 1 < self.env.max_episode_steps:
                a = self.ascaling.inverse(actions.to(torch.float32).reshape(self.B, 1, self.env.J)).contiguous()
                out = self.env.step_device(a.reshape(self.B, self.env.J))
                self.env.collect_step(actions=a, out=out, obs_store=self.obs_store, act_store=self.act_store,
                                      vminmax=self.oscaling.minmax, agent_obs=agent_buf,
                                      rec=(buf.obs[t], buf.actions[t], buf.nxtobs[t], buf.rewards[t], buf.truncated[t],
                                           buf.steps[t]),
                                      lower=self.oscaling.lower, upper=self.oscaling.upper, frozen=self.oscaling.frozen,
                                      agent_stride=self.agent_sensor.stride)
                self._act_row[0] = self._obs_row[0] = True
                self._episode_step += 1
                last_obs, stale = agent_buf, True
                continue
            if stale:
                last_stored, stale = self._valid(self.obs_store, self._obs_row), False
            last_obs, rewards, terminated, truncated, infos = self.step(actions)
            buf.obs[t].copy_(last_stored)
            last_stored = self._valid(self.obs_store, self._obs_row)
            buf.actions[t].copy_(self._valid(self.act_store, self._act_row))
            if "final_observation" in infos:                        # every env of the batch at once
                buf.nxtobs[t].copy_(self.finals[:, -S:])
            else:
                buf.nxtobs[t].copy_(last_stored)
            buf.rewards[t].copy_(rewards)
            buf.truncated[t].copy_(truncated)
            buf.steps[t].copy_(infos["step"])
        if fused:
            last_obs = last_obs.clone()              # agent_buf is reused by the next fused step
        return buf, last_obs

    # -- CUDA-graph rollout ----------------------------------------------------------------------
    def rollout_graphed(self, select_action: Callable[[torch.Tensor], torch.Tensor], num_steps: int,
                        last_obs: Optional[torch.Tensor] = None, seed: Optional[int] = None):
        """``rollout`` with the whole step -- policy forward, action scaling, the env kernel, store /
        scaling bookkeeping and the writes into the transition buffers -- captured ONCE in a CUDA
        graph and replayed per step (one graph launch instead of ~40 small kernel launches; the
        loop is launch-bound, above all with the spectral solver whose period kernel takes 30 us).
        Requirements: ``num_steps`` of the stores = 1 (the reference's MBRL loop), a ``select_action``
        made of capturable torch ops with fixed shapes, CUDA tensors.  The step on which the episode
        ends (auto-reset, burn-in launch with a fresh seed) runs eagerly on the same in-place state.
        Returns ``(RolloutBatch, last_obs)``; the batch lives in buffers that the next graphed
        rollout of the same length overwrites."""
        if self.num_steps != 1:
            raise ValueError("rollout_graphed needs num_steps == 1")
        if last_obs is None:
            last_obs = self.reset(seed=seed)
        dev = last_obs.device
        key = (id(select_action), num_steps)
        g = self._graphs.get(key)
        if g is None:
            lead = (num_steps, self.B)
            st = self.obs_store
            buf = RolloutBatch(
                obs=torch.empty(lead + tuple(st.shape[1:]), dtype=st.dtype, device=dev),
                actions=torch.empty(lead + (1, 1, self.env.J), dtype=torch.float32, device=dev),
                nxtobs=torch.empty(lead + tuple(st.shape[1:]), dtype=st.dtype, device=dev),
                rewards=torch.empty(lead, dtype=torch.float64, device=dev),
                terminated=torch.zeros(lead, dtype=torch.bool, device=dev),
                truncated=torch.empty(lead, dtype=torch.bool, device=dev),
                steps=torch.empty(lead, dtype=torch.int64, device=dev))
            g = dict(buf=buf, obs_in=last_obs.clone(), t=torch.zeros(1, dtype=torch.int64, device=dev), graph=None,
                     store=self.obs_store)
            self._graphs = {key: g}
        if g["store"] is not self.obs_store:         # an explicit reset() re-created the stores
            g["graph"], g["store"] = None, self.obs_store
        buf, obs_in, t_idx = g["buf"], g["obs_in"], g["t"]
        obs_in.copy_(last_obs)
        t_idx.zero_()

        def one_step():
            with torch.no_grad():
                actions = select_action(obs_in)
            buf.obs.index_copy_(0, t_idx, self.obs_store.unsqueeze(0))          # stored obs before the step
            new_obs, rewards, terminated, truncated, infos = self.step(actions)
            buf.actions.index_copy_(0, t_idx, self.act_store.unsqueeze(0))
            nxt = self.finals if "final_observation" in infos else self.obs_store
            buf.nxtobs.index_copy_(0, t_idx, nxt.unsqueeze(0))
            buf.rewards.index_copy_(0, t_idx, rewards.unsqueeze(0))
            buf.truncated.index_copy_(0, t_idx, truncated.unsqueeze(0))
            buf.steps.index_copy_(0, t_idx, infos["step"].unsqueeze(0))
            obs_in.copy_(new_obs)
            t_idx.add_(1)

        use_collect = (self.fused and hasattr(self.env, "collect_step") and self.oscaling.minmax is not None
                       and self.act_store.is_contiguous() and self.obs_store.is_contiguous())

        def fused_step():                            # the regular step with ks_collect: ~22 kernels, no tensor-op bookkeeping
            with torch.no_grad():
                actions = select_action(obs_in)
            a = self.ascaling.inverse(actions.to(torch.float32).reshape(self.B, 1, self.env.J)).contiguous()
            out = self.env.step_device(a.reshape(self.B, self.env.J))
            self.env.collect_step(actions=a, out=out, obs_store=self.obs_store, act_store=self.act_store,
                                  vminmax=self.oscaling.minmax, agent_obs=obs_in,
                                  rec=(buf.obs, buf.actions, buf.nxtobs, buf.rewards, buf.truncated, buf.steps),
                                  lower=self.oscaling.lower, upper=self.oscaling.upper,
                                  frozen=self.oscaling.frozen, agent_stride=self.agent_sensor.stride, slot_index=t_idx)
            self._episode_step += 1

        regular_step = fused_step if use_collect else one_step
        for _ in range(num_steps):
            if self._episode_step + 1 >= self.env.max_episode_steps:
                one_step()                           # episode end: eager (auto-reset inside step())
                continue
            if g["graph"] is None:
                # one real step on a side stream as warm-up (library handles, allocator), then -- unless
                # the next step ends the episode -- the capture; a capture executes nothing, so exactly
                # one transition is produced in this iteration
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    regular_step()
                torch.cuda.current_stream(dev).wait_stream(side)
                if self._episode_step + 1 < self.env.max_episode_steps:
                    graph = torch.cuda.CUDAGraph()
                    host_count = self._episode_step
                    with torch.cuda.graph(graph):
                        regular_step()
                    self._episode_step = host_count      # undo the capture's host-side bookkeeping
                    g["graph"] = graph
                continue
            g["graph"].replay()
            self._episode_step += 1
        return buf, obs_in.clone()
