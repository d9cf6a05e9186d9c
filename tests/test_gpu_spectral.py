"""GPU parity of the spectral ETDRK4 solver (``solver="etdrk4"``, csrc/ks_etd.cuh) through the C ABI.

The reference has no spectral solver, so the checker is the NumPy restatement
of the published algorithm (``oracle/ks_etdrk4.py``: Cox & Matthews 2002 / Kassam & Trefethen 2005)
on identical inputs: state <= 1e-10 relative L2 per control period in fp64 (<= 1e-4 in fp32), reward
to the same level, counters / flags exact.  The jets, reward definition, truncation and observation
cast are the reference's and are compared with the same fixtures as the FD-RK4 path.  The distance
to the reference's own (finite-difference) trajectories is asserted only as an order of magnitude.
"""
import numpy as np
import pytest

from ks_testutil import load_golden, rel_l2

pytestmark = pytest.mark.gpu

TOL64 = 1e-10
TOL32 = 1e-4
DT, S = 0.025, 10          # 10 ETDRK4 steps of 0.025 = the reference's 0.25 time units per control period


# Two lane layouts of the same solver at N = 64: 8 lanes x 8 registers per env pair (csrc/ks_etd.cuh, large
# batches) and 16 lanes x 4 registers (csrc/ks_etd16.cuh, small batches; what ks_create picks by itself
# for every batch size used here).  Every test of this file runs with both.
_PPL = {"value": 0}


@pytest.fixture(autouse=True, params=[4, 8], ids=["16lanes", "8lanes"])
def etd_layout(request):
    _PPL["value"] = request.param
    yield request.param
    _PPL["value"] = 0


def ppl(N=64):
    return _PPL["value"] if N == 64 else 0


def make_env(B, **kw):
    from model_based_pde_control_b200 import KSVecEnv

    kw.setdefault("dt", DT)
    kw.setdefault("cfg_steps", S)
    kw.setdefault("points_per_lane", ppl(kw.get("N", 64)))
    env = KSVecEnv(B, solver="etdrk4", **kw)
    if env.N == 64:
        assert env.launch_info()["lanes_per_env"] == (16 if _PPL["value"] == 4 else 8)
    return env


def oracle_step(env, u0, actions, **kw):
    from oracle import ks_etdrk4 as ke, ks_numpy as ko

    phi = ko.forcing(actions.reshape(len(u0), env.J), env.forcing.matrix())
    return ke.step(u0, phi, env.N, env.L, env.dt, env.cfg_steps, **kw)


def smooth_states(rng, B, N=64, amp=1.5):
    """Attractor-like fields: a few low Fourier modes with random phases."""
    x = np.arange(N) / N
    u = np.zeros((B, N))
    for m in range(1, 6):
        u += rng.normal(0, amp / m, (B, 1)) * np.cos(2 * np.pi * m * (x[None] + rng.uniform(0, 1, (B, 1))))
    return u


@pytest.mark.parametrize("B", [1, 2, 13, 64, 1000])
def test_one_period_vs_oracle(B):
    rng = np.random.default_rng(B)
    env = make_env(B)
    assert env.max_episode_steps == 400 and env.burnin_periods == 800
    u0 = smooth_states(rng, B)
    a = rng.uniform(-1, 1, (B, 1, env.J)).astype(np.float32)
    env.set_state(u0, 3)
    obs, rew, term, trunc, info = env.step(a)
    u1, ts = env.get_state()
    u_ref, r_ref = oracle_step(env, u0, a)
    assert rel_l2(u1, u_ref).max() <= TOL64, rel_l2(u1, u_ref).max()
    assert np.abs((rew - r_ref) / r_ref).max() <= TOL64
    assert np.array_equal(obs[:, 0], u1.astype(np.float32))
    assert (ts == 4).all() and (info["step"] == 4).all() and not trunc.any() and not term.any()
    env.close()


@pytest.mark.parametrize("N,L,J,B,precision", [(128, 44.0, 4, 9, "f64"), (256, 88.0, 8, 7, "f64"), (256, 88.0, 8, 64, "f64"),
                                                 (128, 44.0, 4, 33, "f32"), (256, 88.0, 8, 10, "f32"),
                                                 (256, 22.0, 4, 5, "f64")])
def test_larger_grids_vs_oracle(N, L, J, B, precision):
    """N = 128 / 256: the four-step core plus radix-2 / radix-4 shuffle butterflies across lanes
    (large-domain configuration L = 88, 8 jets; and a finer grid on the default domain)."""
    from model_based_pde_control_b200 import KSVecEnv

    rng = np.random.default_rng(N + B)
    dt, steps = (0.025, 10) if L / N > 0.3 else (0.0125, 20)
    env = KSVecEnv(B, dict(N=N, L=L, dt=dt, cfg_steps=steps), Xi=[k / J for k in range(J)], solver="etdrk4",
                   precision=precision)
    u0 = np.concatenate([smooth_states(rng, B, 64)] * (N // 64), axis=1) * 0.7 + rng.uniform(-0.2, 0.2, (B, N))
    if precision == "f32":
        u0 = u0.astype(np.float32).astype(np.float64)
    a = rng.uniform(-1, 1, (B, 1, J)).astype(np.float32)
    env.set_state(u0, 0)
    obs, rew, term, trunc, info = env.step(a)
    u1, ts = env.get_state()
    u_ref, r_ref = oracle_step(env, u0, a)
    tol = TOL64 if precision == "f64" else TOL32
    assert rel_l2(u1, u_ref).max() <= tol, rel_l2(u1, u_ref).max()
    assert np.abs((rew - r_ref) / r_ref).max() <= tol
    assert np.array_equal(obs[:, 0], u1.astype(np.float32)) and (ts == 1).all()
    # K periods in one launch == K steps, also on the larger grids
    import torch
    acts = torch.as_tensor(rng.uniform(-1, 1, (3, B, J)).astype(np.float32)).cuda()
    env.set_state(u0, 0)
    env.rollout_device(acts)
    ur, _ = env.get_state()
    env.set_state(u0, 0)
    for k in range(3):
        env.step_device(acts[k])
    us, _ = env.get_state()
    assert np.array_equal(ur, us)
    env.close()


@pytest.mark.parametrize("N,L,J,B", [(64, 22.0, 4, 21), (256, 88.0, 8, 6)])
def test_dissipation_reward_mode(N, L, J, B):
    """reward_mode="dissipation": -(mean uxx^2 + mean ux^2 + mean u*phi) per sub-step with spectral
    derivatives (two extra inverse transforms per ETDRK4 step), "ux" = d(u^2)/dx as in kuramoto.py:67-70."""
    from model_based_pde_control_b200 import KSVecEnv
    from oracle import ks_etdrk4 as ke, ks_numpy as ko

    rng = np.random.default_rng(N)
    env = KSVecEnv(B, dict(N=N, L=L, dt=DT, cfg_steps=S), Xi=[k / J for k in range(J)], solver="etdrk4",
                   reward_mode="dissipation", points_per_lane=ppl(N))
    u0 = np.concatenate([smooth_states(rng, B, 64)] * (N // 64), axis=1) * 0.7 + rng.uniform(-0.1, 0.1, (B, N))
    a = rng.uniform(-1, 1, (B, 1, J)).astype(np.float32)
    env.set_state(u0, 0)
    _, rew, *_ = env.step(a)
    u1, _ = env.get_state()
    phi = ko.forcing(a.reshape(B, J), env.forcing.matrix())
    u_ref, r_ref = ke.step(u0, phi, N, L, DT, S, reward_mode="dissipation")
    assert rel_l2(u1, u_ref).max() <= TOL64
    assert np.abs((rew - r_ref) / r_ref).max() <= TOL64, np.abs((rew - r_ref) / r_ref).max()
    # the state does not depend on the reward mode (two kernel instantiations: equal up to rounding)
    env2 = KSVecEnv(B, dict(N=N, L=L, dt=DT, cfg_steps=S), Xi=[k / J for k in range(J)], solver="etdrk4",
                    points_per_lane=ppl(N))
    env2.set_state(u0, 0)
    _, rew_l2, *_ = env2.step(a)
    assert rel_l2(env2.get_state()[0], u1).max() < 1e-13 and not np.allclose(rew_l2, rew)
    env.close(); env2.close()


def test_rough_initial_condition_and_no_dealias():
    """White-noise states (every mode excited, as the reset's U(-0.4,0.4) draw) with and without the 2/3 rule."""
    rng = np.random.default_rng(5)
    B = 32
    u0 = rng.uniform(-0.4, 0.4, (B, 64))
    a = rng.uniform(-1, 1, (B, 1, 4)).astype(np.float32)
    for dealias in (True, False):
        env = make_env(B, dealias=dealias)
        env.set_state(u0, 0)
        env.step(a)
        u1, _ = env.get_state()
        u_ref, _ = oracle_step(env, u0, a, dealias=dealias)
        assert rel_l2(u1, u_ref).max() <= TOL64
        env.close()


@pytest.mark.parametrize("dt,steps", [(0.25, 1), (0.05, 5), (0.001, 250)])
def test_other_step_sizes(dt, steps):
    rng = np.random.default_rng(11)
    B = 16
    env = make_env(B, dt=dt, cfg_steps=steps)
    u0 = smooth_states(rng, B)
    a = rng.uniform(-1, 1, (B, 1, 4)).astype(np.float32)
    env.set_state(u0, 0)
    _, rew, *_ = env.step(a)
    u1, _ = env.get_state()
    u_ref, r_ref = oracle_step(env, u0, a)
    assert rel_l2(u1, u_ref).max() <= TOL64
    assert np.abs((rew - r_ref) / r_ref).max() <= TOL64
    env.close()


def test_distance_to_the_reference_scheme_is_the_spatial_truncation_error():
    """Against the reference's FD-RK4 trajectory (golden fixture) the spectral state differs by the
    finite-difference truncation error at dx = 0.34 (~2e-3 per control period), not by 1e-10: this
    is why the parity path is the FD-RK4 kernel and this solver is validated against its own oracle."""
    g = load_golden("attractor_default_random")
    env = make_env(2)
    env.set_state(np.tile(g["u0"], (2, 1)), int(g["t0"]))
    import torch
    out = env.step_device(torch.as_tensor(np.tile(g["actions"][0], (2, 1))).cuda())
    u1, _ = env.get_state()
    d = rel_l2(u1[0], g["u"][0])
    assert 1e-4 < d < 5e-3, d
    # reward: left Riemann sum with 10 instead of 250 samples of -mean(u^2)
    assert abs(out["reward"][0].item() - g["reward"][0]) < 5e-3 * abs(g["reward"][0])
    env.close()


def test_rollout_equals_repeated_steps_bitwise_and_batch_position_independence():
    import torch

    rng = np.random.default_rng(3)
    B, K = 24, 5
    u0 = smooth_states(rng, B)
    acts = torch.as_tensor(rng.uniform(-1, 1, (K, B, 4)).astype(np.float32)).cuda()
    env = make_env(B)
    env.set_state(u0, 0)
    out = env.rollout_device(acts)
    u_roll, ts = env.get_state()
    obs_roll = out["obs"].cpu().numpy()
    rew_roll = out["reward"].cpu().numpy()
    env.set_state(u0, 0)
    for k in range(K):
        o = env.step_device(acts[k])
        assert np.array_equal(o["obs"].cpu().numpy().reshape(B, -1), obs_roll[k].reshape(B, -1))
        assert np.array_equal(o["reward"].cpu().numpy(), rew_roll[k])
    u_steps, _ = env.get_state()
    assert np.array_equal(u_roll, u_steps) and (ts == K).all()
    env.close()
    # Two envs share one complex transform, so an env's bits depend on its partner at rounding level:
    # whole PAIRS may move anywhere in the batch / to another shard bit for bit (even shard
    # boundaries keep multi-GPU results identical); a different partner changes the result by ~1e-15.
    pairs = rng.permutation(B // 2)[:8]
    perm = np.stack([2 * pairs, 2 * pairs + 1], 1).reshape(-1)
    env2 = make_env(len(perm))
    env2.set_state(u0[perm], 0)
    env2.rollout_device(acts[:, torch.as_tensor(perm).cuda()].contiguous())
    u2, _ = env2.get_state()
    assert np.array_equal(u2, u_roll[perm])
    env2.close()
    odd = rng.permutation(B)[:17]
    env3 = make_env(17)
    env3.set_state(u0[odd], 0)
    env3.rollout_device(acts[:, torch.as_tensor(odd).cuda()].contiguous())
    u3, _ = env3.get_state()
    assert rel_l2(u3, u_roll[odd]).max() < 1e-12
    env3.close()


def test_fp32_mode():
    rng = np.random.default_rng(8)
    B = 40
    env = make_env(B, precision="f32")
    u0 = smooth_states(rng, B)
    a = rng.uniform(-1, 1, (B, 1, 4)).astype(np.float32)
    env.set_state(u0, 0)
    _, rew, *_ = env.step(a)
    u1, _ = env.get_state()
    u_ref, r_ref = oracle_step(env, u0.astype(np.float32).astype(np.float64), a)
    assert rel_l2(u1, u_ref).max() <= TOL32
    assert np.abs((rew - r_ref) / r_ref).max() <= TOL32
    env.close()


def test_env_semantics_truncation_autoreset_sensors_nonfinite():
    rng = np.random.default_rng(2)
    B = 6
    env = make_env(B, burnin_periods=3, sensor_stride=4, ic="device")
    obs = env.reset(seed=1)
    assert obs.shape == (B, 1, 16)
    u, ts = env.get_state()
    assert (ts == 0).all() and np.array_equal(obs[:, 0], u[:, 2::4].astype(np.float32))
    env.set_state(u, env.max_episode_steps - 1)
    obs, rew, term, trunc, info = env.step(rng.uniform(-1, 1, (B, 1, 4)).astype(np.float32))
    assert trunc.all() and info["_final_observation"].all() and (info["step"] == 400).all()
    _, ts = env.get_state()
    assert (ts == 0).all()                      # auto-reset ran the (shortened) burn-in
    # masked reset leaves the other envs untouched
    before, _ = env.get_state()
    import torch
    mask = np.array([1, 0, 0, 1, 0, 1], bool)
    env.set_state(before, 7)
    env.reset_device(seed=9, mask=torch.as_tensor(mask))
    after, ts = env.get_state()
    assert np.array_equal(after[~mask], before[~mask]) and (ts[~mask] == 7).all()
    assert (ts[mask] == 0).all() and not np.array_equal(after[mask], before[mask])
    # non-finite states raise, as np.seterr(over="raise") does in the reference
    bad = before.copy()
    bad[1, 5] = np.inf
    env.set_state(bad, 0)
    with pytest.raises(FloatingPointError):
        env.step(np.zeros((B, 1, 4), np.float32))
    env.close()


def test_burnin_reaches_the_attractor():
    """reset() = IC + 800 no-op periods in one launch; mean u^2 afterwards is the attractor's."""
    env = make_env(512, ic="device")
    env.reset(seed=3)
    u, ts = env.get_state()
    assert (ts == 0).all() and np.isfinite(u).all()
    m = (u * u).mean()
    assert 1.0 < m < 2.5, m          # FD reference fixture: 1.43 under random actions; unforced is bimodal
    env.close()


def test_reset_pool_uses_the_parents_solver():
    """reset_mode="pool" with the spectral solver: the burner must be the parent's twin (same solver, dt, dealiasing) --
    an FD-RK4 burner at dt = 0.025 would blow up and inject non-finite states."""
    B = 12
    env = make_env(B, Tmax=0.75, burnin_periods=6, reset_mode="pool", pool_slots=2, ic="device")
    assert env.max_episode_steps == 3
    env.reset(seed=1)
    a = np.zeros((B, 4), np.float32)
    for ep in range(3):
        for k in range(3):
            obs, rew, term, trunc, info = env.step(a)          # would raise FloatingPointError on a blown-up pool state
        assert trunc.all() and np.isfinite(obs).all() and (info["step"] == 3).all()
    pool = env._pool
    assert pool is not None and pool.burner.solver == "etdrk4" and pool.burner.dt == env.dt
    assert pool.burner.launch_info()["lanes_per_env"] == env.launch_info()["lanes_per_env"]
    u, ts = env.get_state()
    assert (ts == 0).all() and np.abs(u).max() < 10.0
    env.close()
