"""``KSEnv`` -- the single-env ``gym.Env`` surface of ``KuramotoSivashinskyEnv-v0`` on the GPU.

What the reference's id resolves to is ``TimeLimit(KuramotoSivashinskyEnv(**config))``
(``pdegym/kuramoto/__init__.py:8-12,26-31``).  Its consumers are

* ``pdecontrol/surrogates/evaluation/generate.py:23-38``: ``obs = env.reset()``; ``env.step(
  env.action_space.sample())`` with ``(1,J)`` actions -> ``(1,N)`` observation, scalar reward, two
  Python bools, ``{"step": int}``;
* ``pdecontrol/mbrl/mbrl.py:78,157,196,215-240,298-300,310-312``: the controller keeps ONE such env
  for its metadata -- the *single-env* ``observation_space`` ``(1,N)`` / ``action_space`` ``(1,J)``
  handed to ``WorldVecEnv``, ``reward_func``, ``forcing``, ``scenario``, ``cfg_steps``, ``dt``,
  ``unwrapped.max_episode_steps``.

``KSEnv`` is that object over a one-env ``KSVecEnv`` of its own, or -- ``KSEnv(vec=envs)`` -- a view of
an existing vector env that shares its handle and costs nothing: spaces, helpers and state access
work, ``step`` / ``reset`` are refused there (the members of a vector env advance together).  Semantics follow
``kuramoto.py:78-116`` -- no auto-reset (that is the vector env's job), observations are the
float64 state ``(1,N)``, ``reset(seed)`` seeds the legacy MT19937 stream exactly like
``np.random.seed(seed)`` does.  All numerics run in the CUDA kernels of ``libks_b200.so``.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .env import KSVecEnv
from .spaces import EnvBase


class KSEnv(EnvBase):
    """Single Kuramoto-Sivashinsky control env (``gym.Env`` surface, 5-tuple step API)."""

    metadata = {"render.modes": ["rgb_array"]}
    reward_range = (-float("inf"), float("inf"))
    eps = np.finfo(np.float32).eps

    # attributes of the reference env that are simply those of the vector env
    _FORWARDED = ("L", "N", "J", "cfg_steps", "Ttrans", "Tmax", "dt", "noise", "sigma", "lmbda", "objective", "dx", "x",
                  "Xi", "max_episode_steps", "forcing", "noop", "scenario", "rhs", "evaluate", "reward_mode",
                  "precision", "solver", "burnin_periods", "sensor_stride", "obs_len", "device")

    def __init__(self, config: Optional[dict] = None, *, vec: Optional[KSVecEnv] = None, index: int = 0, **kwargs):
        if vec is None:
            vec, index, self._owns = KSVecEnv(1, dict(config or {}), **kwargs), 0, True
        else:
            if config or kwargs:
                raise TypeError("pass either a config (own one-env handle) or vec= (shared handle), not both")
            self._owns = False
        if not (0 <= index < vec.num_envs):
            raise IndexError(f"env index {index} outside [0, {vec.num_envs})")
        self.vec, self.index = vec, int(index)
        self.observation_space = vec.single_observation_space      # (1, N)  kuramoto.py:76
        self.action_space = vec.single_action_space                # (1, J)  kuramoto.py:75
        self._actions = torch.zeros((1, vec.J), dtype=torch.float32, device=vec.device)

    def __getattr__(self, name):
        # only reached when normal lookup fails
        if name in KSEnv._FORWARDED:
            return getattr(self.__dict__["vec"], name)
        raise AttributeError(f"{type(self).__name__!r} object has no attribute {name!r}")

    @property
    def unwrapped(self):
        return self

    # ------------------------------------------------------------------ state (plain attributes in the reference)
    @property
    def u(self) -> np.ndarray:
        """The float64 state ``[N]`` (``env.u``, kuramoto.py:106)."""
        return self.vec.get_state()[0][self.index]

    @u.setter
    def u(self, value) -> None:
        if self.vec.num_envs == 1:
            self.vec.set_state(np.asarray(value, dtype=np.float64).reshape(1, self.vec.N))
        else:
            u, _ = self.vec.get_state()
            u[self.index] = np.asarray(value, dtype=np.float64).reshape(self.vec.N)
            self.vec.set_state(u)

    @property
    def timestep(self) -> int:
        return int(self.vec.get_state()[1][self.index])

    @timestep.setter
    def timestep(self, value: int) -> None:
        ts = self.vec.get_state()[1]
        ts[self.index] = int(value)
        self.vec.set_state(None, ts)

    @property
    def time(self) -> float:
        """``timestep * cfg_steps * dt`` (kuramoto.py:131-133)."""
        return self.timestep * self.vec.cfg_steps * self.vec.dt

    # ------------------------------------------------------------------ gym.Env
    def _stepping_vec(self, what: str) -> KSVecEnv:
        if self.vec is None:
            raise RuntimeError("KSEnv is closed")
        if self.vec.num_envs != 1:
            raise RuntimeError(f"KSEnv.{what}() on a view of a {self.vec.num_envs}-env vector env: its members advance "
                               "together -- step the vector env, or build a stand-alone KSEnv(config)")
        return self.vec

    def _obs(self, u: np.ndarray) -> np.ndarray:
        s = self.vec.sensor_stride
        return u[self.index, s // 2::s].reshape(1, -1)

    def reset(self, seed: Optional[int] = None, return_info: bool = False, options: Optional[dict] = None,
              *, u0=None, burnin_periods: Optional[int] = None, **kwargs):
        """``KuramotoSivashinskyEnv.reset`` (kuramoto.py:100-116): ``np.random.seed(seed)``-compatible
        initial condition, 800 no-op control periods in one launch, ``timestep = 0``.  Returns the
        float64 observation ``(1,N)`` (and ``{"step": 0}`` with ``return_info``)."""
        vec = self._stepping_vec("reset")
        vec.reset(seed=seed, u0=None if u0 is None else np.asarray(u0, dtype=np.float64).reshape(1, vec.N),
                  burnin_periods=burnin_periods)
        u, ts = vec.get_state()
        vec._raise_if_nonfinite(u[self.index])
        obs = self._obs(u)
        if return_info:
            return obs, {"step": int(ts[self.index])}
        return obs

    def step(self, action):
        """One control period (kuramoto.py:78-98): ``action`` ``(1,J)`` / ``(J,)`` ->
        ``(obs (1,N) float64, reward float, False, truncated bool, {"step": int})``.  No auto-reset."""
        vec = self._stepping_vec("step")
        a = np.array(action, dtype=np.float32)                                            # :79
        if a.size != vec.J:
            raise ValueError(f"action has shape {a.shape}, expected (1, {vec.J})")
        self._actions.copy_(torch.from_numpy(a.reshape(1, vec.J)))
        out = vec.step_device(self._actions)
        packed = out["packed"].cpu()                                                       # one D2H, synchronises
        o = vec._out_offsets
        i = self.index
        reward = float(packed[o[0] + 8 * i:o[0] + 8 * i + 8].view(torch.float64)[0])
        step = int(packed[o[2] + 4 * i:o[2] + 4 * i + 4].view(torch.int32)[0])
        truncated = bool(packed[o[3] + i])
        if bool(packed[o[4] + i]):
            raise FloatingPointError("overflow encountered in KS state (np.seterr(over='raise') in the reference)")
        u, _ = vec.get_state()
        return self._obs(u), reward, False, truncated, {"step": step}

    def reward_func(self, obs, phi=None, *args, **kwargs):
        """``env.reward_func(obs, phi)`` (kuramoto.py:64-73).  Also accepts a whole batch
        ``[M,(1,)N]`` -> ``[M]`` in ONE launch: ``mbrl/world/world.py:170`` calls it once per sample from
        a Python loop; replacing that list comprehension by ``self.reward_func(orescaled, arescaled)``
        costs one ``ks_eval`` launch per model step instead of one per sample."""
        return self.vec.reward_func(obs, phi)      # (the L2 objective ignores phi: the world model passes the ACTION there)

    def render(self, mode="rgb_array"):
        raise NotImplementedError("rendering is not part of the control path (the reference has none either)")

    def close(self):
        if self._owns and self.vec is not None:
            self.vec.close()
        self.vec = None

    def __del__(self):
        try:
            if self.__dict__.get("_owns") and self.__dict__.get("vec") is not None:
                self.vec.close()
        except Exception:
            pass


class TimeLimit:
    """``gym.wrappers.TimeLimit(env, max_episode_steps, new_step_api=True)`` as the reference's
    ``make`` applies it (``pdegym/kuramoto/__init__.py:10``); used when gym is not importable.
    ``truncated`` is set once ``max_episode_steps`` steps have elapsed since ``reset``."""

    def __init__(self, env, max_episode_steps: int, new_step_api: bool = True):
        if not new_step_api:
            raise ValueError("only the 5-tuple step API (new_step_api=True) is implemented, as the reference registers it")
        self.env = env
        self._max_episode_steps = int(max_episode_steps)
        self._elapsed_steps = None
        self.observation_space, self.action_space = env.observation_space, env.action_space
        self.metadata, self.reward_range = env.metadata, env.reward_range

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, reward, terminated, truncated, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)

    def close(self):
        return self.env.close()
