"""TEST DOUBLE of the gym 0.25.2 surface the KS env touches -- NOT gym.

gym (pinned to 0.25.2 by the reference, ``pyproject.toml:13``) is neither installed in this image nor in
the offline wheelhouse, so the branch of ``model_based_pde_control_b200/spaces.py`` /
``registration.py`` that runs when ``import gym`` succeeds could never execute.  This package, put on
``sys.path`` by ``tests/test_gym_shaped_package.py`` in a SUBPROCESS only, restates -- from the documented
behaviour of gym 0.25.2, recalled, not copied -- exactly the pieces that branch uses:

* ``gym.Env``, ``gym.spaces.Box``, ``gym.vector.VectorEnv(num_envs, observation_space, action_space)`` with
  ``step = step_async + step_wait``, ``close`` -> ``close_extras``; ``gym.vector.VectorEnvWrapper`` with its
  ``assert isinstance(env, VectorEnv)``;
* ``gym.wrappers.TimeLimit(env, max_episode_steps, new_step_api)``;
* ``gym.envs.register(id, entry_point, **kwargs)``, ``gym.envs.registry``, ``gym.make(id, **kwargs)`` resolving
  ``"module:function"`` entry points.
"""
import importlib
import types

import numpy as np


class Env:
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    observation_space = None
    action_space = None

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env, new_step_api=False):
        self.env = env
        self.observation_space, self.action_space = env.observation_space, env.action_space
        self.metadata = env.metadata

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def close(self):
        return self.env.close()


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else \
            np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else \
            np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
        self._rng = np.random.default_rng(seed)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


spaces = types.ModuleType("gym.spaces")
spaces.Box = _Box


def _batch_space(space, n):
    return _Box(np.repeat(space.low[None], n, axis=0), np.repeat(space.high[None], n, axis=0), dtype=space.dtype)


class _VectorEnv(Env):
    def __init__(self, num_envs, observation_space, action_space, new_step_api=False):
        self.num_envs = num_envs
        self.is_vector_env = True
        self.observation_space = _batch_space(observation_space, num_envs)
        self.action_space = _batch_space(action_space, num_envs)
        self.closed = False
        self.viewer = None
        self.single_observation_space = observation_space
        self.single_action_space = action_space

    def reset_async(self, seed=None, return_info=False, options=None):
        pass

    def reset_wait(self, seed=None, return_info=False, options=None):
        raise NotImplementedError

    def reset(self, *, seed=None, return_info=False, options=None):
        self.reset_async(seed=seed, return_info=return_info, options=options)
        return self.reset_wait(seed=seed, return_info=return_info, options=options)

    def step_async(self, actions):
        pass

    def step_wait(self, **kwargs):
        raise NotImplementedError

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close_extras(self, **kwargs):
        pass

    def close(self, **kwargs):
        if self.closed:
            return
        self.close_extras(**kwargs)
        self.closed = True


class _VectorEnvWrapper(_VectorEnv):
    def __init__(self, env):
        assert isinstance(env, _VectorEnv)
        self.env = env

    def reset_async(self, **kwargs):
        return self.env.reset_async(**kwargs)

    def reset_wait(self, **kwargs):
        return self.env.reset_wait(**kwargs)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step_async(self, actions):
        return self.env.step_async(actions)

    def step_wait(self):
        return self.env.step_wait()

    def close(self, **kwargs):
        return self.env.close(**kwargs)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(f"attempted to get missing private attribute '{name}'")
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped


vector = types.ModuleType("gym.vector")
vector.VectorEnv = _VectorEnv
vector.VectorEnvWrapper = _VectorEnvWrapper


class _TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps=None, new_step_api=False):
        super().__init__(env, new_step_api)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None
        self.new_step_api = new_step_api

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, reward, terminated, truncated, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)


wrappers = types.ModuleType("gym.wrappers")
wrappers.TimeLimit = _TimeLimit

envs = types.ModuleType("gym.envs")
envs.registry = {}


def _register(id, entry_point=None, **kwargs):
    envs.registry[id] = dict(entry_point=entry_point, kwargs=kwargs)


envs.register = _register


def make(id, **kwargs):
    spec = envs.registry[id]
    mod, fn = spec["entry_point"].split(":")
    call_kwargs = {k: v for k, v in spec["kwargs"].items() if k == "new_step_api"}
    call_kwargs.update(kwargs)
    return getattr(importlib.import_module(mod), fn)(**call_kwargs)


import sys as _sys  # noqa: E402

for _m in (spaces, vector, wrappers, envs):
    _sys.modules[_m.__name__] = _m
