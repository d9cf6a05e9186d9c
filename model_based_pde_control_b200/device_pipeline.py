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

    def update(self, values: torch.Tensor) -> None:
        if self.frozen:
            return
        lo, hi = values.amin().to(torch.float32), values.amax().to(torch.float32)
        self.vmin = lo if self.vmin is None else torch.minimum(lo, self.vmin.to(lo.device))
        self.vmax = hi if self.vmax is None else torch.maximum(hi, self.vmax.to(hi.device))

    def __call__(self, values: torch.Tensor) -> torch.Tensor:
        vmin, vmax = self.vmin.to(values.device), self.vmax.to(values.device)
        return (values - vmin) / (vmax - vmin) * (self.upper - self.lower) + self.lower

    def inverse(self, values: torch.Tensor) -> torch.Tensor:
        vmin, vmax = self.vmin.to(values.device), self.vmax.to(values.device)
        return (values - self.lower) / (self.upper - self.lower) * (vmax - vmin) + vmin


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
    ``action_bounds``  ``(low, high)`` of the env's action space; agent actions in ``[-1,1]`` are mapped
                       onto it (identity for the KS env), frozen like the reference's ``ascaling``
    """

    def __init__(self, env, num_steps: int = 1, obs_scale=(-1.0, 1.0), frozen_obs_scaling: bool = False,
                 action_bounds=(-1.0, 1.0), agent_sensor_stride: int = 1):
        self.env = env
        self.B = env.num_envs
        self.num_steps = num_steps
        self.oscaling = ScaleTransformDevice(scale=obs_scale, frozen=frozen_obs_scaling)
        self.ascaling = ScaleTransformDevice(scale=(-1.0, 1.0), bounds=action_bounds, frozen=True)
        self.agent_sensor = SensorTransformDevice(agent_sensor_stride)
        self.obs_store = self.finals = self.obs_mask = None
        self.act_store = self.act_mask = None
        self._episode_step = None          # host-side step counter (all envs synchronous)
        self._pending_final = None

    # -- StoreNObsVecWrapper -------------------------------------------------------------------
    def _ostore_reset(self, obs):
        self.obs_store = obs.unsqueeze(1).repeat(1, self.num_steps, *([1] * (obs.dim() - 1))).clone()
        self.finals = torch.zeros_like(self.obs_store)
        self.obs_mask = torch.zeros((self.B, self.num_steps), dtype=torch.bool, device=obs.device)
        self.obs_mask[:, -1] = True

    def _ostore_step(self, obs, final_obs, final_mask):
        if final_obs is not None:
            self.finals[final_mask] = final_obs[final_mask].unsqueeze(1)
            self.obs_mask[final_mask] = False
        self.obs_store[:, 0] = obs
        self.obs_store = torch.roll(self.obs_store, -1, dims=1)
        self.obs_mask[:, 0] = True
        self.obs_mask = torch.roll(self.obs_mask, -1, dims=1)

    # -- StoreNActionsVecWrapper ---------------------------------------------------------------
    def _astore_reset(self, like):
        self.act_store = torch.zeros((self.B, self.num_steps, 1, self.env.J), dtype=torch.float32, device=like.device)
        self.act_mask = torch.zeros((self.B, self.num_steps), dtype=torch.bool, device=like.device)

    def _astore_step(self, actions):
        self.act_store[:, 0] = actions
        self.act_mask[:, 0] = True
        self.act_store = torch.roll(self.act_store, -1, dims=1)
        self.act_mask = torch.roll(self.act_mask, -1, dims=1)

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
        self._ostore_step(obs, final_obs, final_mask)
        if final_obs is not None:
            self.act_mask[final_mask, :-1] = False                      # StoreNActionsVecWrapper.step_wait
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
                last_obs: Optional[torch.Tensor] = None, seed: Optional[int] = None):
        """Batched ``Worker.rollout`` (worker.py:39-93): ``num_steps`` env steps with a GPU-resident
        ``select_action(scaled_obs) -> actions in [-1,1]``; returns ``(RolloutBatch, last_obs)`` with all
        tensors on the device.  No host synchronisation inside the loop."""
        if last_obs is None:
            last_obs = self.reset(seed=seed)
        last_stored = self.obs_store[self.obs_mask].reshape(self.B, -1, *self.obs_store.shape[2:]).clone()
        rec = {k: [] for k in RolloutBatch._fields}
        for _ in range(num_steps):
            with torch.no_grad():
                actions = select_action(last_obs)
            last_obs, rewards, terminated, truncated, infos = self.step(actions)
            obs = last_stored
            last_stored = self.obs_store[self.obs_mask].reshape(self.B, -1, *self.obs_store.shape[2:]).clone()
            nxtobs = last_stored.clone()
            stored_actions = self.act_store[self.act_mask].reshape(self.B, -1, 1, self.env.J).clone()
            if "final_observation" in infos:
                idx = infos["_final_observation"]
                nxtobs[idx] = self.finals[idx][:, -obs.shape[1]:]
            for k, v in zip(RolloutBatch._fields, (obs, stored_actions, nxtobs, rewards, terminated, truncated,
                                                   infos["step"])):
                rec[k].append(v)
        return RolloutBatch(*(torch.stack(rec[k]) for k in RolloutBatch._fields)), last_obs
