"""Factory functions with the reference's names and argument meaning.

``pdegym/kuramoto/__init__.py:8-12``: ``make(config)`` builds the env and wraps it in
``TimeLimit(env, max_episode_steps)``; ``:26-31`` registers it as ``KuramotoSivashinskyEnv-v0``.
``pdecontrol/mbrl/mbrl.py:81-86``: ``gym.vector.make(env_id, num_envs=cpus, new_step_api=True)``.
Here the episode limit is enforced inside the kernel (``truncated = timestep >= 400``), so no
TimeLimit wrapper is needed, and the vector env *is* the env.
"""
from __future__ import annotations

from .env import KSVecEnv

ENV_ID = "KuramotoSivashinskyEnv-v0"


def make(config: dict | None = None, new_step_api: bool = True, num_envs: int = 1, **kwargs) -> KSVecEnv:
    """``pdegym.kuramoto.make(config)`` -> a (vectorised) env with ``num_envs`` members."""
    if not new_step_api:
        raise ValueError("only the 5-tuple step API (new_step_api=True) is implemented, as the reference "
                         "registers it")
    return KSVecEnv(num_envs, dict(config or {}), **kwargs)


def vector_make(env_id: str = ENV_ID, num_envs: int = 1, new_step_api: bool = True, **kwargs) -> KSVecEnv:
    """Stand-in for ``gym.vector.make(env_id, num_envs=..., new_step_api=True)``."""
    if env_id != ENV_ID:
        raise ValueError(f"unknown env id {env_id!r}; this package provides {ENV_ID!r} only")
    return make(kwargs.pop("config", None), new_step_api=new_step_api, num_envs=num_envs, **kwargs)
