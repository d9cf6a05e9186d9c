"""``DeviceEnvPipeline`` over the real CUDA env: a GPU-resident rollout (policy + env + wrapper
plumbing, no host synchronisation inside the loop) against the same env driven through the host
API.  The wrapper logic itself is pinned to the reference's wrappers in test_device_pipeline.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_gpu_resident_rollout_matches_host_api_path():
    import torch
    from model_based_pde_control_b200 import DeviceEnvPipeline, KSVecEnv

    B, T = 33, 12
    cfg = dict(cfg_steps=10, Tmax=0.05)                       # 5-step episodes
    env = KSVecEnv(B, cfg, ic="device", burnin_periods=3)
    pipe = DeviceEnvPipeline(env)
    gen = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.rand((T, B, 1, 4), generator=gen, device="cuda") * 2 - 1
    k = [0]

    def policy(obs):
        assert obs.is_cuda and obs.shape == (B, 1, 64) and obs.dtype == torch.float32
        k[0] += 1
        return acts[k[0] - 1]

    first = pipe.reset(seed=11)
    batch, last = pipe.rollout(policy, T, last_obs=first)
    assert all(t.is_cuda for t in batch)
    assert batch.obs.shape == (T, B, 1, 1, 64) and batch.actions.shape == (T, B, 1, 1, 4)
    assert batch.rewards.shape == (T, B) and batch.rewards.dtype == torch.float64
    steps = batch.steps.cpu().numpy()
    assert (steps[:, 0] == np.array([1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 1, 2])).all() and (steps == steps[:, :1]).all()
    trunc = batch.truncated.cpu().numpy()
    assert trunc[4].all() and trunc[9].all() and trunc.sum() == 2 * B and not batch.terminated.any()

    # first episode through the host API of a second env with the same device seed
    ref = KSVecEnv(B, cfg, ic="device", burnin_periods=3)
    obs0 = ref.reset(seed=11)
    assert np.array_equal(batch.obs[0, :, 0].cpu().numpy(), obs0)
    for t in range(4):
        o, r, term, tr, info = ref.step(acts[t].cpu().numpy())
        assert np.array_equal(batch.nxtobs[t, :, 0].cpu().numpy(), o)
        assert np.array_equal(batch.obs[t + 1, :, 0].cpu().numpy(), o)
        assert np.array_equal(batch.rewards[t].cpu().numpy(), r)
    # the truncating step: nxtobs is the FINAL observation, the next obs is post-reset
    pre = ref.get_state()[0]
    chk = KSVecEnv(B, cfg)
    chk.set_state(pre, 4)
    final = chk.step_device(acts[4].reshape(B, 4))["obs"].cpu().numpy()
    assert np.array_equal(batch.nxtobs[4, :, 0, 0].cpu().numpy(), final)
    assert not np.array_equal(batch.obs[5, :, 0, 0].cpu().numpy(), final)
    # running min/max scaling saw every raw observation: agent obs within [-1, 1]
    assert float(last.min()) >= -1.0 - 1e-6 and float(last.max()) <= 1.0 + 1e-6
    lo = min(float(batch.obs.min()), float(batch.nxtobs.min()))
    hi = max(float(batch.obs.max()), float(batch.nxtobs.max()))
    assert abs(float(pipe.oscaling.vmin) - lo) < 1e-6 and abs(float(pipe.oscaling.vmax) - hi) < 1e-6
    for e in (env, ref, chk):
        e.close()
