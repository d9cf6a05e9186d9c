"""Minimal stand-ins for the gym 0.25 objects the KS env exposes, used only when ``gym`` is not
installed (it is pinned to 0.25.2 by the reference, ``pyproject.toml:13``, and absent here).
When gym is importable the real classes are used so that ``gym.vector.VectorEnvWrapper``'s
isinstance checks pass."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    import gym as _gym
    from gym.spaces import Box  # type: ignore
    from gym.vector import VectorEnv as _GymVectorEnv  # type: ignore

    _GymEnv = _gym.Env
    HAVE_GYM = True
except Exception:  # ImportError or a broken install
    _gym = None
    HAVE_GYM = False

    class Box:  # type: ignore[no-redef]
        """``gym.spaces.Box(low, high, shape, dtype)`` with the attributes the wrappers read."""

        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low)
            self.shape = tuple(shape)
            self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else \
                np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else \
                np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi, size=self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class _GymEnv:  # type: ignore[no-redef]
        """Protocol base: the subset of ``gym.Env`` the reference relies on."""

        metadata: dict = {}

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass

    class _GymVectorEnv:  # type: ignore[no-redef]
        """Protocol base: the subset of ``gym.vector.VectorEnv`` the reference relies on."""

        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs = num_envs
            self.is_vector_env = True
            self.single_observation_space = observation_space
            self.single_action_space = action_space
            self.observation_space = batch_space(observation_space, num_envs)
            self.action_space = batch_space(action_space, num_envs)
            self.closed = False

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()

        def close(self, **kwargs):
            if not self.closed:
                self.close_extras(**kwargs)
                self.closed = True

        def close_extras(self, **kwargs):
            pass

        @property
        def unwrapped(self):
            return self


VectorEnvBase = _GymVectorEnv
EnvBase = _GymEnv


def batch_space(space, n: int):
    """Leading batch axis on a Box (what ``gym.vector.utils.batch_space`` does for Box)."""
    low = np.repeat(space.low[None], n, axis=0)
    high = np.repeat(space.high[None], n, axis=0)
    return Box(low, high, shape=low.shape, dtype=space.dtype)
