"""Factory functions and gym registration with the reference's names and argument meaning.

``pdegym/kuramoto/__init__.py:8-12``: ``make(config)`` builds ONE env and wraps it in
``TimeLimit(env, max_episode_steps, new_step_api=True)``; ``:26-31`` registers that factory as
``KuramotoSivashinskyEnv-v0`` (``order_enforce=False, new_step_api=True``).
``pdecontrol/mbrl/mbrl.py:78``: ``gym.make(env_id, new_step_api=True)`` -- the controller's metadata
env; ``mbrl.py:81-86``: ``gym.vector.make(env_id, num_envs=cpus, new_step_api=True)`` -- gym then
spawns one process per env.  Here:

* ``make(config)``          -> ``TimeLimit(KSEnv(**config))``, the single-env ``gym.Env`` surface
                               (``single_env.py``); gym's own ``TimeLimit`` when gym is importable;
* ``vector_make(id, n)``    -> ONE ``KSVecEnv`` with ``n`` members on the GPU (the episode limit is
                               enforced in the kernel: ``truncated = timestep >= 400``);
* ``register()``            -> ``gym.envs.register(id=ENV_ID, entry_point=<this module>:make, ...)`` with
                               the reference's keyword arguments, when gym is importable.  Called on
                               package import, so ``gym.make("KuramotoSivashinskyEnv-v0", config=...)``
                               (``surrogates/evaluation/generate.py:23``) resolves to the GPU env.
"""
from __future__ import annotations

from .env import KSVecEnv
from .single_env import KSEnv, TimeLimit as _LocalTimeLimit
from .spaces import HAVE_GYM

ENV_ID = "KuramotoSivashinskyEnv-v0"
ENTRY_POINT = "model_based_pde_control_b200.registration:make"


def make(config: dict | None = None, new_step_api: bool = True, **kwargs):
    """``pdegym.kuramoto.make(config)``: one env, wrapped in ``TimeLimit`` (``__init__.py:8-12``)."""
    if not new_step_api:
        raise ValueError("only the 5-tuple step API (new_step_api=True) is implemented, as the reference "
                         "registers it")
    env = KSEnv(dict(config or {}), **kwargs)
    if HAVE_GYM:    # pragma: no cover - gym is not installed in the build image
        from gym.wrappers import TimeLimit

        return TimeLimit(env, env.unwrapped.max_episode_steps, new_step_api=True)
    return _LocalTimeLimit(env, env.unwrapped.max_episode_steps, new_step_api=True)


def vector_make(env_id: str = ENV_ID, num_envs: int = 1, new_step_api: bool = True, **kwargs) -> KSVecEnv:
    """What ``gym.vector.make(env_id, num_envs=..., new_step_api=True)`` is used for at
    ``mbrl.py:81-86``: ``num_envs`` members, here inside one GPU-resident vector env."""
    if env_id != ENV_ID:
        raise ValueError(f"unknown env id {env_id!r}; this package provides {ENV_ID!r} only")
    if not new_step_api:
        raise ValueError("only the 5-tuple step API (new_step_api=True) is implemented")
    return KSVecEnv(num_envs, dict(kwargs.pop("config", None) or {}), **kwargs)


def register(force: bool = False) -> bool:
    """Register ``ENV_ID`` with gym exactly as ``pdegym/kuramoto/__init__.py:26-31`` does, pointing at
    this package's ``make``.  Returns False when gym is not importable (nothing to register with) or
    the id is already taken and ``force`` is False."""
    if not HAVE_GYM:
        return False
    import gym  # pragma: no cover - gym is not installed in the build image

    registry = getattr(gym.envs, "registry", None)          # pragma: no cover
    taken = False                                           # pragma: no cover
    try:                                                    # pragma: no cover
        taken = ENV_ID in registry or ENV_ID in getattr(registry, "env_specs", {})
    except TypeError:                                       # pragma: no cover
        pass
    if taken and not force:                                 # pragma: no cover
        return False
    gym.envs.register(id=ENV_ID, entry_point=ENTRY_POINT, order_enforce=False, new_step_api=True)   # pragma: no cover
    return True                                             # pragma: no cover
