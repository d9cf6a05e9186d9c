"""GPU parity tests proper: the CUDA path (through the C ABI via ``KSVecEnv``) against

* the committed golden fixtures produced by executing the reference (``tests/golden``), and
* the oracle (``oracle/``) on the same seeded inputs,

at the tolerance BASELINE.json's north star states: state after one control period <= 1e-10
relative L2 in fp64 (<= 1e-4 in fp32 mode); rewards to the same level; flags / counters exact.
"""
import numpy as np
import pytest

from ks_testutil import STEP_CASES, load_golden, rel_l2

pytestmark = pytest.mark.gpu

TOL64 = 1e-10   # north star: "<= 1e-10 relative L2 in fp64"
TOL32 = 1e-4    # north star: "<= 1e-4 in an optional fp32 mode"


def make_env(g, num_envs, **kw):
    from model_based_pde_control_b200 import KSVecEnv

    cfg = dict(L=float(g["L"]), N=int(g["N"]), cfg_steps=int(g["cfg_steps"]), dt=float(g["dt"]),
               sigma=float(g["sigma"]), Tmax=float(g["Tmax"]))
    return KSVecEnv(num_envs, cfg, Xi=list(g["Xi"]), **kw)


@pytest.mark.parametrize("name", STEP_CASES)
def test_golden_step_sequences(name):
    """Every reference trajectory: inject u0/t0, apply the recorded actions period by period.
    The batch holds the same env 5 times (different warp slots must agree bit for bit)."""
    g = load_golden(name)
    B = 5
    env = make_env(g, B)
    assert np.array_equal(env.forcing.matrix(), g["F"]), "forcing matrix not bit-identical to the reference"
    assert env.max_episode_steps == int(g["max_episode_steps"])
    env.set_state(np.tile(g["u0"], (B, 1)), int(g["t0"]))
    for k, a in enumerate(g["actions"]):
        # compare period by period from the reference's own pre-state (no error accumulation)
        if k > 0:
            env.set_state(np.tile(g["u"][k - 1], (B, 1)), int(g["step"][k - 1]))
        out = env.step_device(__import__("torch").as_tensor(np.tile(a, (B, 1))).cuda())
        u, ts = env.get_state()
        obs = out["obs"].cpu().numpy()
        rew = out["reward"].cpu().numpy()
        assert rel_l2(u[0], g["u"][k]) <= TOL64, (name, k, rel_l2(u[0], g["u"][k]))
        assert abs(rew[0] - g["reward"][k]) <= TOL64 * abs(g["reward"][k])
        assert (u == u[0]).all() and (rew == rew[0]).all(), "warp slots disagree"
        assert np.array_equal(obs, u.astype(np.float32)), "obs must be the float32 cast of the state"
        assert (ts == g["step"][k]).all()
        assert (out["step"].cpu().numpy() == g["step"][k]).all()
        assert (out["truncated"].cpu().numpy().astype(bool) == bool(g["truncated"][k])).all()
        assert not out["nonfinite"].cpu().numpy().any()
    env.close()


def test_golden_trajectory_free_running():
    """KAT 2 without re-injection: 10 consecutive periods through the host (NumPy) API."""
    g = load_golden("kat2_default_10periods")
    env = make_env(g, 3)
    env.set_state(np.tile(g["u0"], (3, 1)), int(g["t0"]))
    for k, a in enumerate(g["actions"]):
        obs, rew, term, trunc, info = env.step(np.tile(a.reshape(1, 1, -1), (3, 1, 1)))
        assert obs.shape == (3, 1, 64) and obs.dtype == np.float32
        assert rew.shape == (3,) and rew.dtype == np.float64
        assert term.dtype == bool and not term.any() and trunc.dtype == bool
        assert abs(rew[0] - g["reward"][k]) <= 1e-9 * abs(g["reward"][k])
        assert (info["step"] == g["step"][k]).all()
    u, _ = env.get_state()
    assert rel_l2(u[0], g["u"][-1]) <= 1e-9   # ten periods of (non-chaotic-scale) error growth
    env.close()


@pytest.mark.parametrize("N,L,J,P", [(64, 22.0, 4, 0), (64, 22.0, 4, 2), (64, 22.0, 4, 4), (64, 22.0, 4, 16), (256, 88.0, 8, 0),
                                     (48, 16.5, 3, 3), (30, 10.3125, 2, 3), (18, 6.1875, 1, 2),
                                     (256, 88.0, 8, 16), (128, 44.0, 4, 0), (96, 33.0, 4, 0), (96, 33.0, 3, 8),
                                     (100, 34.375, 5, 0), (32, 11.0, 2, 0), (16, 5.5, 1, 0), (512, 176.0, 8, 0)])
def test_random_batch_vs_oracle(N, L, J, P):
    """Seeded random batch vs the C oracle for several grid sizes / lane layouts (P = points per
    lane; 0 = automatic), including ragged last warps (B not a multiple of envs-per-warp)."""
    from model_based_pde_control_b200 import KSVecEnv
    from oracle import ks_c, ks_numpy as ko

    B = 37
    Xi = [k / J for k in range(J)]
    cfg = ko.KSConfig(L=L, N=N, Xi=Xi, cfg_steps=50)
    env = KSVecEnv(B, dict(L=L, N=N, cfg_steps=50), Xi=Xi, points_per_lane=P)
    F = ko.forcing_matrix(cfg)
    assert np.array_equal(F, env.forcing.matrix())
    rng = np.random.default_rng(N * 7 + J)
    u0 = rng.uniform(-2.0, 2.0, (B, N))
    a = rng.uniform(-1, 1, (B, J)).astype(np.float32)
    env.set_state(u0, 3)
    obs, rew, _, trunc, info = env.step(a)
    u1, ts = env.get_state()
    u_ref, r_ref = ks_c.step(cfg, u0, ko.forcing(a, F))
    assert rel_l2(u1, u_ref).max() <= TOL64
    assert np.abs((rew - r_ref) / r_ref).max() <= TOL64
    assert (ts == 4).all() and (info["step"] == 4).all() and not trunc.any()
    env.close()


def test_phi_override_equals_in_kernel_forcing():
    import torch
    from model_based_pde_control_b200 import KSVecEnv
    from oracle import ks_numpy as ko

    B = 16
    env = KSVecEnv(B, dict(cfg_steps=20))
    rng = np.random.default_rng(5)
    u0 = rng.uniform(-1, 1, (B, 64))
    a = rng.uniform(-1, 1, (B, 4)).astype(np.float32)
    env.set_state(u0, 0)
    out = env.step_device(torch.as_tensor(a).cuda())
    u_a = env.get_state()[0]
    r_a = out["reward"].cpu().numpy()
    phi = ko.forcing(a, env.forcing.matrix())
    env.set_state(u0, 0)
    out = env.step_device(None, phi=torch.as_tensor(phi).cuda())
    u_p = env.get_state()[0]
    assert np.array_equal(u_a, u_p), "in-kernel a@F FMA chain differs from the oracle's phi"
    assert np.array_equal(r_a, out["reward"].cpu().numpy())
    env.close()


def test_dissipation_reward_mode_vs_reference_closure():
    """The intended reward against vectors produced by the REFERENCE'S OWN ``dissipation`` closure and ``rhs``
    (``tests/golden/make_golden_dissipation.py``; ``env.step`` itself raises in that mode): the batched
    ``ks_eval`` reward and one whole control period on the default and on the large domain."""
    from model_based_pde_control_b200 import KSVecEnv

    g = load_golden("dissipation_kat")
    env = KSVecEnv(1, reward_mode="dissipation")
    r = env.reward_func(g["U"], g["PHI"])
    assert np.abs(r / g["reward"] - 1).max() <= 1e-12
    env.set_state(g["u0"][None], 0)
    obs, rew, *_ = env.step(g["a1"])
    assert rel_l2(env.get_state()[0][0], g["u1"]) <= TOL64 and abs(rew[0] / float(g["r1"]) - 1) <= TOL64
    env.close()
    big = KSVecEnv(1, dict(L=88.0, N=256, cfg_steps=int(g["cfg_steps_large"])), Xi=[k / 8 for k in range(8)],
                   reward_mode="dissipation")
    big.set_state(g["u0_large"][None], 0)
    obs, rew, *_ = big.step(g["a8"])
    assert rel_l2(big.get_state()[0][0], g["u1_large"]) <= TOL64 and abs(rew[0] / float(g["r1_large"]) - 1) <= TOL64
    big.close()


def test_dissipation_reward_mode_vs_oracle():
    from model_based_pde_control_b200 import KSVecEnv
    from oracle import ks_c, ks_numpy as ko

    B = 9
    cfg = ko.KSConfig(reward_mode="dissipation", cfg_steps=40)
    env = KSVecEnv(B, dict(cfg_steps=40), reward_mode="dissipation")
    rng = np.random.default_rng(9)
    u0 = rng.uniform(-2, 2, (B, 64))
    a = rng.uniform(-1, 1, (B, 4)).astype(np.float32)
    env.set_state(u0, 0)
    _, rew, _, _, _ = env.step(a)
    u1 = env.get_state()[0]
    u_ref, r_ref = ks_c.step(cfg, u0, ko.forcing(a, env.forcing.matrix()))
    assert rel_l2(u1, u_ref).max() <= TOL64
    assert np.abs((rew - r_ref) / r_ref).max() <= TOL64
    # the reference's selector: any truthy objective string means L2 (kuramoto.py:72)
    assert KSVecEnv(1, objective="dissipation").reward_mode == "l2"
    assert KSVecEnv(1, objective="").reward_mode == "dissipation"
    env.close()


def test_fp32_mode_tolerance():
    from model_based_pde_control_b200 import KSVecEnv
    from oracle import ks_c, ks_numpy as ko

    g = load_golden("attractor_default_random")
    B = 8
    env = KSVecEnv(B, precision="f32")
    env.set_state(np.tile(g["u0"], (B, 1)), 0)
    _, rew, _, _, _ = env.step(np.tile(g["actions"][0], (B, 1)))
    u1 = env.get_state()[0]
    assert rel_l2(u1[0], g["u"][0]) <= TOL32
    assert abs(rew[0] - g["reward"][0]) <= TOL32 * abs(g["reward"][0])
    # and a random batch against the fp64 oracle
    cfg = ko.KSConfig()
    rng = np.random.default_rng(3)
    u0 = rng.uniform(-2, 2, (B, 64))
    a = rng.uniform(-1, 1, (B, 4)).astype(np.float32)
    env.set_state(u0, 0)
    env.step(a)
    u_ref, _ = ks_c.step(cfg, u0.astype(np.float32).astype(np.float64), ko.forcing(a, env.forcing.matrix()))
    assert rel_l2(env.get_state()[0], u_ref).max() <= TOL32
    env.close()


def test_rollout_equals_repeated_steps_bitwise():
    import torch
    from model_based_pde_control_b200 import KSVecEnv

    B, K = 21, 4
    env = KSVecEnv(B, dict(cfg_steps=25, Tmax=10.0))     # 10 / (1e-3 * 25) = 400 steps per episode
    assert env.max_episode_steps == 400
    rng = np.random.default_rng(11)
    u0 = rng.uniform(-1.5, 1.5, (B, 64))
    acts = torch.as_tensor(rng.uniform(-1, 1, (K, B, 4)).astype(np.float32)).cuda()
    env.set_state(u0, 397)
    single = []
    for k in range(K):
        o = env.step_device(acts[k])
        single.append({n: t.clone() for n, t in o.items()})
    u_single = env.get_state()[0]
    env.set_state(u0, 397)
    roll = env.rollout_device(acts)
    u_roll, ts = env.get_state()
    assert np.array_equal(u_single, u_roll) and (ts == 397 + K).all()
    for k in range(K):
        for n in ("obs", "reward", "truncated", "step", "nonfinite"):
            assert torch.equal(single[k][n], roll[n][k]), (k, n)
    assert roll["truncated"].cpu().numpy()[:, 0].tolist() == [0, 0, 1, 1]   # steps 398,399,400,401
    env.close()


def test_batch_position_independence():
    """An env's result must not depend on its index, the batch size or the lane layout of its
    neighbours: this is what makes multi-GPU sharding bit-exact."""
    from model_based_pde_control_b200 import KSVecEnv

    rng = np.random.default_rng(13)
    B = 50
    u0 = rng.uniform(-2, 2, (B, 64))
    a = rng.uniform(-1, 1, (B, 4)).astype(np.float32)
    full = KSVecEnv(B, dict(cfg_steps=30))
    full.set_state(u0, 0)
    _, r_full, _, _, _ = full.step(a)
    u_full = full.get_state()[0]
    for lo, hi in [(0, 25), (25, 50), (7, 8), (13, 50)]:
        part = KSVecEnv(hi - lo, dict(cfg_steps=30))
        part.set_state(u0[lo:hi], 0)
        _, r_part, _, _, _ = part.step(a[lo:hi])
        assert np.array_equal(part.get_state()[0], u_full[lo:hi])
        assert np.array_equal(r_part, r_full[lo:hi])
        part.close()
    full.close()


def test_eval_kernel_vs_golden_rhs():
    from model_based_pde_control_b200 import KSVecEnv

    g = load_golden("rhs_default")
    env = KSVecEnv(1)
    rhs, (ux, uxx, uxxxx) = env.rhs(g["u"], g["phi"])
    for got, want in ((rhs, g["rhs"]), (ux, g["ux"]), (uxx, g["uxx"]), (uxxxx, g["uxxxx"])):
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    r = env.reward_func(g["u"].reshape(1, -1))
    assert abs(r - (-(1 / 64) * np.linalg.norm(g["u"]) ** 2)) <= 1e-13 * abs(r)
    env.close()


def test_nonfinite_raises_like_numpy_overflow():
    from model_based_pde_control_b200 import KSVecEnv

    env = KSVecEnv(6, dict(cfg_steps=5))
    u0 = np.zeros((6, 64))
    u0[4] = 1e200          # u**2 overflows on the first sub-step (np.seterr(over="raise") in the reference)
    env.set_state(u0, 0)
    with pytest.raises(FloatingPointError):
        env.step(np.zeros((6, 4), np.float32))
    flags = env.nonfinite()
    assert flags.tolist() == [False] * 4 + [True, False]
    env.set_state(np.zeros((6, 64)), 0)
    assert not env.nonfinite().any()
    env.close()


def test_state_is_independent_of_the_lane_layout():
    """The layout chooser picks the points per lane from the batch size (P = 4 for small batches, 16
    for large ones), and a sharded run may use another P than the single-GPU run of the same envs.
    Every grid point's arithmetic is the same sequence of operations whatever P is, so the STATE is
    bitwise identical across layouts; the reward's summation order follows the layout (tree over the
    lane's points, then over the lanes), so rewards agree to rounding only."""
    from model_based_pde_control_b200 import KSVecEnv

    rng = np.random.default_rng(21)
    B = 37
    u0 = rng.uniform(-2, 2, (B, 64))
    a = rng.uniform(-1, 1, (B, 4)).astype(np.float32)
    out = {}
    for P in (2, 4, 8, 16):
        env = KSVecEnv(B, dict(cfg_steps=40), points_per_lane=P)
        env.set_state(u0, 0)
        _, r, *_ = env.step(a)
        out[P] = (env.get_state()[0], r)
        env.close()
    for P in (2, 8, 16):
        assert np.array_equal(out[P][0], out[4][0])
        assert np.abs(out[P][1] / out[4][1] - 1).max() < 1e-14


def test_full_size_batch_properties():
    """BASELINE sizes (65 536 envs on one GPU): size-independent properties instead of an oracle run --
    the first / last envs equal the same envs stepped in a small batch (bitwise state), K periods in
    one launch equal K single-period launches (bitwise), nothing goes non-finite."""
    import torch
    from model_based_pde_control_b200 import KSVecEnv

    B, K = 65536, 3
    rng = np.random.default_rng(5)
    u0 = rng.uniform(-1.5, 1.5, (B, 64))
    acts = rng.uniform(-1, 1, (K, B, 4)).astype(np.float32)
    big = KSVecEnv(B, dict(cfg_steps=25))
    big.set_state(u0, 0)
    big.rollout_device(torch.as_tensor(acts).cuda())
    u_roll, ts = big.get_state()
    big.set_state(u0, 0)
    for k in range(K):
        out = big.step_device(torch.as_tensor(acts[k]).cuda())
    u_steps, _ = big.get_state()
    assert np.array_equal(u_roll, u_steps) and (ts == K).all() and not big.nonfinite().any()
    for sl in (slice(0, 40), slice(B - 33, B)):
        small = KSVecEnv(sl.stop - sl.start, dict(cfg_steps=25))
        small.set_state(u0[sl], 0)
        for k in range(K):
            small.step_device(torch.as_tensor(acts[k][sl]).cuda())
        assert np.array_equal(small.get_state()[0], u_roll[sl])
        small.close()
    big.close()
