"""Device-side wrapper plumbing (``device_pipeline.py``) against the REFERENCE'S OWN wrappers
(``pdegym/common/vec_wrappers.py`` + ``transforms.py``) on a deterministic fake env.  CPU only.

Where ``/root/reference`` exists the reference stack of ``mbrl.py:257-275`` is executed live (and
the trace can be re-recorded with ``python tests/test_device_pipeline.py``); elsewhere the
committed trace ``tests/golden/wrapper_trace.npz`` is used.  The worker loop below restates
``pdecontrol/mbrl/worker.py:53-88``."""
import os

import numpy as np
import pytest
import torch

from ks_testutil import GOLDEN
from model_based_pde_control_b200.device_pipeline import DeviceEnvPipeline, ScaleTransformDevice, SensorTransformDevice
from oracle.ref_loader import reference_available

B, N, J, EP, T = 6, 16, 4, 5, 13          # 13 steps cross two episode boundaries (5-step episodes)
TRACE = os.path.join(GOLDEN, "wrapper_trace.npz")


def base_state(epoch):
    b = np.arange(B, dtype=np.float64)[:, None]
    i = np.arange(N, dtype=np.float64)[None, :]
    return np.sin(0.7 * i + 0.3 * b + 1.1 * epoch) * (1.0 + 0.1 * b) + 0.05 * epoch


def advance(s, a, t):
    return 0.9 * s + 0.1 * np.tile(a.reshape(B, J).astype(np.float64), (1, N // J)) + 0.01 * (t + 1)


def agent_actions(k):
    rng = np.random.default_rng(100 + k)
    return rng.uniform(-1, 1, (B, 1, J)).astype(np.float32)


class FakeVecEnvNP:
    """gym-0.25-style vector env with auto-reset, NumPy in/out."""

    def __init__(self, Box):
        self.num_envs = B
        self.single_observation_space = Box(-np.inf, np.inf, shape=(1, N), dtype=np.float32)
        self.single_action_space = Box(-1.0, 1.0, shape=(1, J), dtype=np.float32)
        self.observation_space = Box(-np.inf, np.inf, shape=(B, 1, N), dtype=np.float32)
        self.action_space = Box(-1.0, 1.0, shape=(B, 1, J), dtype=np.float32)
        self.epoch, self.t = -1, 0

    def reset(self, **kwargs):
        self.epoch += 1
        self.t = 0
        self.s = base_state(self.epoch)
        return self.s.astype(np.float32).reshape(B, 1, N)

    def step_async(self, actions):
        self._a = np.asarray(actions, dtype=np.float32)

    def step_wait(self):
        self.s = advance(self.s, self._a, self.t)
        self.t += 1
        obs = self.s.astype(np.float32).reshape(B, 1, N)
        rew = -(self.s ** 2).mean(axis=1)
        trunc = np.full(B, self.t >= EP)
        infos = {"step": np.full(B, self.t)}
        if trunc.all():
            finals = np.empty(B, dtype=object)
            for i in range(B):
                finals[i] = self.s[i].reshape(1, N).copy()
            infos["final_observation"] = finals
            infos["_final_observation"] = trunc.copy()
            obs = self.reset()
        return obs, rew, np.zeros(B, bool), trunc, infos


class FakeVecEnvTorch:
    """The same dynamics behind the device API of ``KSVecEnv``."""

    def __init__(self):
        self.num_envs, self.N, self.J, self.max_episode_steps, self.sensor_stride = B, N, J, EP, 1
        self.epoch, self.t = -1, 0

    def reset_device(self, seed=None, **kw):
        self.epoch += 1
        self.t = 0
        self.s = base_state(self.epoch)

    def get_state_device(self):
        return torch.from_numpy(self.s.copy()), torch.full((B,), self.t, dtype=torch.int32)

    def step_device(self, actions):
        self.s = advance(self.s, actions.numpy(), self.t)
        self.t += 1
        return {"obs": torch.from_numpy(self.s.astype(np.float32)), "reward": torch.from_numpy(-(self.s ** 2).mean(axis=1)),
                "step": torch.full((B,), self.t, dtype=torch.int32),
                "truncated": torch.full((B,), int(self.t >= EP), dtype=torch.uint8)}


def reference_trace():
    """Run the reference's wrapper stack + worker loop on the NumPy fake env."""
    from oracle.ref_loader import load_reference_wrappers

    vw, tr = load_reference_wrappers()
    import gym  # the stub installed by the loader

    env = FakeVecEnvNP(gym.spaces.Box)
    oscaling = tr.ScaleTransform(batched=True, aggregate=True, frozen=False)               # mbrl.py:148
    low = env.single_action_space.low[np.newaxis, ...]
    high = env.single_action_space.high[np.newaxis, ...]
    ascaling = tr.ScaleTransform(bounds=(low, high), aggregate=True, frozen=True, batched=True).Inverse   # :151-155
    sensor = tr.BatchTransform(tr.SensorTransform(stride=1))
    ostore = vw.StoreNObsVecWrapper(env, num_steps=1)                                      # mbrl.py:259-274
    envs = vw.TransformObsWrapper(ostore, oscaling, frozen=False)
    envs = vw.TransformObsWrapper(envs, sensor)
    astore = vw.StoreNActionsVecWrapper(envs, num_steps=1)
    envs = vw.TransformActionWrapper(astore, ascaling, frozen=True)

    rec = {k: [] for k in ("agent_obs", "obs", "actions", "nxtobs", "rewards", "truncated", "steps", "vmin", "vmax")}
    last_obs = envs.reset()
    last_stored = ostore.obs.copy()[ostore.mask]
    rec["agent_obs"].append(np.asarray(last_obs))
    for k in range(T):                                                                     # worker.py:53-88
        last_obs, rewards, terminated, truncated, infos = envs.step(agent_actions(k))
        obs = last_stored.copy()
        last_stored = ostore.obs.copy()[ostore.mask]
        nxtobs = last_stored.copy()
        actions = astore.actions.copy()[astore.mask]
        if "final_observation" in infos:
            index = infos["_final_observation"]
            finals = ostore.finals[index].copy()
            nxtobs[index] = finals[ostore.mask[index]]
        for name, v in (("agent_obs", last_obs), ("obs", obs), ("actions", actions), ("nxtobs", nxtobs),
                        ("rewards", rewards), ("truncated", truncated), ("steps", infos["step"]),
                        ("vmin", oscaling.vmin.numpy().ravel()), ("vmax", oscaling.vmax.numpy().ravel())):
            rec[name].append(np.asarray(v).copy())
    return {k: np.stack(v) for k, v in rec.items()}


def load_trace():
    if reference_available():
        return reference_trace()
    if os.path.exists(TRACE):
        return dict(np.load(TRACE))
    pytest.skip("neither the reference nor tests/golden/wrapper_trace.npz is available")


def test_pipeline_matches_reference_wrapper_stack():
    ref = load_trace()
    pipe = DeviceEnvPipeline(FakeVecEnvTorch(), num_steps=1)
    agent_obs = []

    def policy(o):
        agent_obs.append(o)
        return torch.from_numpy(agent_actions(len(agent_obs) - 1))

    batch, last = pipe.rollout(policy, T, last_obs=pipe.reset())
    # the agent saw exactly the reference's scaled observations (reset + every step)
    seen = torch.stack(agent_obs + [last]).numpy()              # o_0 (reset) .. o_T
    assert np.allclose(seen, ref["agent_obs"], rtol=0, atol=2e-6)
    assert np.allclose(batch.obs[:, :, 0].numpy(), ref["obs"], atol=0)
    assert np.allclose(batch.nxtobs[:, :, 0].numpy(), ref["nxtobs"], atol=0)
    assert np.allclose(batch.actions[:, :, 0].numpy(), ref["actions"], atol=1e-7)
    assert np.allclose(batch.rewards.numpy(), ref["rewards"], rtol=1e-15)
    assert np.array_equal(batch.truncated.numpy(), ref["truncated"])
    assert np.array_equal(batch.steps.numpy(), ref["steps"])
    assert not batch.terminated.any()
    assert ref["truncated"][EP - 1].all() and ref["truncated"][2 * EP - 1].all() and ref["truncated"].sum() == 2 * B
    # at the episode end the replay's next observation is the FINAL one, not the post-reset one
    assert not np.allclose(ref["nxtobs"][EP - 1], ref["obs"][EP])
    assert np.isclose(float(pipe.oscaling.vmin), ref["vmin"][-1, 0]) and np.isclose(float(pipe.oscaling.vmax), ref["vmax"][-1, 0])


def test_scale_and_sensor_transforms_standalone():
    s = ScaleTransformDevice()
    x = torch.tensor([[[-2.0, 0.0, 6.0]]])
    s.update(x)
    assert torch.allclose(s(x), torch.tensor([[[-1.0, -0.5, 1.0]]]))
    s.update(torch.tensor([[[10.0]]]))
    assert float(s.vmax) == 10.0 and float(s.vmin) == -2.0
    assert torch.allclose(s.inverse(s(x)), x, atol=1e-6)
    frozen = ScaleTransformDevice(bounds=(-1.0, 1.0), frozen=True)
    frozen.update(torch.tensor([5.0]))
    assert float(frozen.vmax) == 1.0
    assert torch.allclose(frozen.inverse(torch.tensor([0.25])), torch.tensor([0.25]))       # identity for [-1,1]
    v = torch.arange(10.0)
    assert SensorTransformDevice(1)(v).tolist() == v.tolist()
    assert SensorTransformDevice(4)(v).tolist() == [2.0, 6.0]


if __name__ == "__main__":   # re-record the committed trace from the live reference
    np.savez_compressed(TRACE, **reference_trace())
    print("wrote", TRACE)
