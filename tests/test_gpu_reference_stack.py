"""BASELINE configs[4] -- the drop-in claim itself: the REFERENCE'S OWN data-collection code
(``StoreNObsVecWrapper -> TransformObsWrapper x2 -> StoreNActionsVecWrapper -> TransformActionWrapper``,
``pdecontrol/mbrl/mbrl.py:257-291``, and ``Worker.rollout``, ``pdecontrol/mbrl/worker.py:39-93``) executed
unmodified over a real ``KSVecEnv`` on the GPU, across a truncation with auto-reset.

The reference files are run from ``baseline/_ref`` (``oracle/install_ref.py``: byte-identical install
that travels to the GPU box) or ``/root/reference``, under the stub ``gym`` of ``oracle/ref_loader.py``.
What the replay must contain is derived independently: a twin ``KSVecEnv`` stepped through the device
API from the same initial states with the same actions.
"""
import numpy as np
import pytest
import torch

from oracle.ref_loader import reference_available

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_available(), reason="reference tree not available (baseline/_ref)")]

B, EP = 8, 5
CFG = dict(cfg_steps=10, Tmax=0.05)          # 5-step episodes, 10 RK4 sub-steps per control period


class SeededAgent:
    """Stands in for the SAC policy (``sac.select_action``): seeded uniform actions in agent scale."""

    def __init__(self, J):
        self.J, self.calls, self.seen = J, 0, []

    def select_action(self, obs, deterministic=False):
        self.seen.append(np.asarray(obs).copy())
        rng = np.random.default_rng(500 + self.calls)
        self.calls += 1
        return rng.uniform(-1, 1, (B, 1, self.J)).astype(np.float32)


def build_reference_stack(envs):
    from oracle.ref_loader import load_reference_worker, load_reference_wrappers

    vw, tr = load_reference_wrappers()
    worker_mod, _ = load_reference_worker()
    oscaling = tr.ScaleTransform(batched=True, aggregate=True, frozen=False)                       # mbrl.py:148
    low = envs.single_action_space.low[np.newaxis, ...]
    high = envs.single_action_space.high[np.newaxis, ...]
    ascaling = tr.ScaleTransform(bounds=(low, high), aggregate=True, frozen=True, batched=True).Inverse   # :151-155
    sensor = tr.BatchTransform(tr.SensorTransform(stride=1))                                       # :171,174
    ostore = vw.StoreNObsVecWrapper(envs, num_steps=1)                                             # :259
    stack = vw.TransformObsWrapper(ostore, oscaling, frozen=False)
    stack = vw.TransformObsWrapper(stack, sensor)
    stack = vw.TransformObsWrapper(stack, sensor)             # agent sensor (the world wrapper between them is a pass-through here)
    astore = vw.StoreNActionsVecWrapper(stack, num_steps=1)
    stack = vw.TransformActionWrapper(astore, ascaling, frozen=True)
    return worker_mod.Worker(worker_mod.PDEEnvStack(envs=stack, ostore=ostore, astore=astore)), oscaling


def test_reference_worker_rollout_over_real_ksvecenv():
    from model_based_pde_control_b200 import KSVecEnv

    envs = KSVecEnv(B, CFG, burnin_periods=3, ic="numpy")
    twin = KSVecEnv(B, CFG, burnin_periods=3, ic="numpy")
    worker, oscaling = build_reference_stack(envs)
    agent = SeededAgent(envs.J)

    T = EP + 3                                                  # crosses one truncation + auto-reset
    # Worker.rollout resets the stack itself on its first call (worker.py:48-51) -- without a seed.  Do
    # exactly that here with gym's ``reset(seed=...)`` keyword, which travels down the wrapper stack, so
    # that the twin can start from the same state.
    worker._last_obs = worker.stack.envs.reset(seed=11)
    worker._last_stored_obs = worker.stack.ostore.obs.copy()[worker.stack.ostore.mask]
    n0 = envs.launch_count
    replay = worker.rollout(agent, stop=lambda ts, eps: ts >= B * T)
    assert envs.launch_count >= n0 + T + 2                      # T period kernels + the auto-reset (IC + burn-in launch)
    assert replay.ntimesteps == B * T and replay.nstopped == B

    twin.reset(seed=11)                                         # same MT19937 ICs + same burn-in launch
    obs_prev = twin.get_state()[0].astype(np.float32)
    dev = twin.device
    for k in range(EP):                                         # first episode, open loop on the twin
        a_agent = np.random.default_rng(500 + k).uniform(-1, 1, (B, 1, twin.J)).astype(np.float32)
        # the env-scale actions are what TransformActionWrapper made of the agent's (identity up to rounding)
        a = np.stack([np.asarray(replay.actions[i][k], dtype=np.float32) for i in range(B)])
        assert np.allclose(a.reshape(B, -1), a_agent.reshape(B, -1), atol=1e-6)
        out = twin.step_device(torch.from_numpy(a.reshape(B, twin.J)).to(dev))
        obs_new, rew = out["obs"].cpu().numpy(), out["reward"].cpu().numpy()
        step, trunc = out["step"].cpu().numpy(), out["truncated"].cpu().numpy().astype(bool)
        for i in range(B):
            assert np.array_equal(np.asarray(replay.obs[i][k]).reshape(-1), obs_prev[i]), (k, i)
            assert np.array_equal(np.asarray(replay.nxtobs[i][k]).reshape(-1), obs_new[i]), (k, i)     # incl. the FINAL obs at k = EP-1
            assert replay.rewards[i][k] == rew[i] and int(replay.steps[i][k]) == int(step[i]) == k + 1
            assert bool(replay.truncated[i][k]) == bool(trunc[i]) == (k == EP - 1)
            assert not bool(replay.terminated[i][k])
        obs_prev = obs_new
    # after the truncation: new episodes (fresh replay slots), steps restart at 1, obs = post-reset observation
    for i in range(B):
        ep2 = B + i
        assert len(replay.obs[ep2]) == T - EP and int(replay.steps[ep2][0]) == 1
        assert not np.array_equal(np.asarray(replay.obs[ep2][0]), np.asarray(replay.nxtobs[i][EP - 1]))
        # within the second episode consecutive samples chain: nxtobs[k] == obs[k+1]
        for k in range(T - EP - 1):
            assert np.array_equal(np.asarray(replay.nxtobs[ep2][k]), np.asarray(replay.obs[ep2][k + 1]))
    # the agent saw the running-min/max scaled observations in [-1, 1]
    seen = np.stack(agent.seen)
    assert seen.shape == (T, B, 1, envs.N) and np.all(seen >= -1 - 1e-6) and np.all(seen <= 1 + 1e-6)
    assert float(oscaling.vmax.max()) > float(oscaling.vmin.min())
    for e in (envs, twin):
        e.close()


def test_reference_stack_spaces_are_single_env_shaped():
    """``mbrl.py:298-299`` hands ``self.env.observation_space`` / ``action_space`` (SINGLE-env shapes) to
    ``WorldVecEnv``: the facade gives those, whether stand-alone or as a view of the vector env."""
    from model_based_pde_control_b200 import KSEnv, KSVecEnv

    envs = KSVecEnv(B, CFG, burnin_periods=1)
    env = KSEnv(vec=envs)
    assert env.observation_space.shape == (1, envs.N) and env.action_space.shape == (1, envs.J)
    assert env.unwrapped.max_episode_steps == EP and env.cfg_steps * env.dt == pytest.approx(0.01)
    u = np.random.default_rng(0).uniform(-1, 1, (3, 1, envs.N))
    assert np.allclose(env.reward_func(u, np.zeros((3, 1, envs.J), np.float32)), -(u ** 2).mean(axis=(1, 2)), rtol=1e-14)
    with pytest.raises(RuntimeError):
        env.step(np.zeros((1, envs.J), np.float32))
    envs.close()
