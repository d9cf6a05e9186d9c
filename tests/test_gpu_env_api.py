"""Env-API conformance of ``KSVecEnv`` on the GPU: the gym 0.25.2 vector-env protocol as the
reference consumes it (``pdegym/common/vec_wrappers.py``, ``pdecontrol/mbrl/worker.py:39-93``),
reset / burn-in semantics (``kuramoto.py:100-116``), observation sampling
(``transforms.py:231-247``) and the helper attributes the controller reads (``mbrl.py``)."""
import numpy as np
import pytest

from ks_testutil import load_golden, rel_l2

pytestmark = pytest.mark.gpu


def test_reset_seeded_numpy_ics_and_short_burnin_vs_oracle():
    from model_based_pde_control_b200 import KSVecEnv
    from oracle import ks_c, ks_numpy as ko

    B = 6
    env = KSVecEnv(B, burnin_periods=5)
    obs, info = env.reset(seed=123, return_info=True)
    assert obs.shape == (B, 1, 64) and obs.dtype == np.float32
    assert (info["step"] == 0).all()
    ic = load_golden("reset_ic")["seed123"]
    assert np.array_equal(env.initial_conditions(123)[0], ic), "env 0 must use np.random.seed(seed) stream"
    cfg = ko.KSConfig()
    u0 = np.stack([ko.initial_condition(cfg, 123 + i) for i in range(B)])       # gym seeds env i with seed+i
    assert np.array_equal(env.initial_conditions(123), u0)
    u_ref, _, _ = ks_c.rollout(cfg, u0, None, ko.forcing_matrix(cfg), K=5, want_obs=False)
    u, ts = env.get_state()
    assert rel_l2(u, u_ref).max() <= 1e-10 and (ts == 0).all()
    assert np.array_equal(obs[:, 0], u.astype(np.float32))
    assert env.launch_count == 2          # IC scatter + ONE burn-in launch for all 5 periods
    env.close()


def test_full_reference_reset_800_periods():
    """The reference's reset(seed=5) (golden): 800 burn-in periods in one launch.  Chaotic growth of
    rounding differences over 200 time units limits agreement to ~1e-5 (see tests/test_oracle.py)."""
    from model_based_pde_control_b200 import KSVecEnv

    g = load_golden("reset_full_seed5")
    env = KSVecEnv(2)
    assert env.burnin_periods == 800
    obs = env.reset(seed=5)
    u, ts = env.get_state()
    assert np.array_equal(env.initial_conditions(5)[0], g["u0"])
    assert rel_l2(u[0], g["u"]) <= 1e-4 and (ts == 0).all()
    assert 3.0 < np.linalg.norm(u[1]) < 20.0             # env 1 (seed 6) is on the attractor too (bimodal norm)
    env.close()


def test_device_philox_initial_conditions():
    from model_based_pde_control_b200 import KSVecEnv

    B = 512
    env = KSVecEnv(B, ic="device", burnin_periods=0)
    env.reset(seed=7)
    a = env.get_state()[0]
    env.reset(seed=7)
    b = env.get_state()[0]
    env.reset(seed=8)
    c = env.get_state()[0]
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.min() >= -0.4 and a.max() < 0.4 and abs(a.mean()) < 0.005 and abs(a.std() - 0.8 / 12 ** 0.5) < 0.005
    assert len(np.unique(a)) == a.size
    assert abs(np.corrcoef(a[:-1].ravel(), a[1:].ravel())[0, 1]) < 0.02     # envs are independent streams
    env.reset()                                                              # seed=None: OS entropy
    assert not np.array_equal(env.get_state()[0], a)
    env.close()


def test_autoreset_on_truncation_like_gym_vector_env():
    """All envs truncate on the same step: infos carry final_observation exactly in the form the
    reference's wrappers consume (np.asarray(list(...), float32), vec_wrappers.py:26-30)."""
    from model_based_pde_control_b200 import KSVecEnv

    B = 5
    env = KSVecEnv(B, dict(cfg_steps=10, Tmax=0.03), burnin_periods=3)   # 3 steps per episode
    assert env.max_episode_steps == 3
    rng = np.random.default_rng(0)
    env.reset(seed=1)
    for k in range(2):
        obs, rew, term, trunc, info = env.step(rng.uniform(-1, 1, (B, 1, 4)).astype(np.float32))
        assert not trunc.any() and "final_observation" not in info and (info["step"] == k + 1).all()
    pre_u = env.get_state()[0]
    a = rng.uniform(-1, 1, (B, 1, 4)).astype(np.float32)
    obs, rew, term, trunc, info = env.step(a)
    assert trunc.all() and not term.any() and (info["step"] == 3).all()
    assert info["_final_observation"].all() and info["final_observation"].dtype == object
    finals = np.asarray(list(info["final_observation"]), dtype=np.float32)      # the reference's conversion
    assert finals.shape == (B, 1, 64)
    # final obs = state after the truncating step, computed from pre_u by one more period
    chk = KSVecEnv(B, dict(cfg_steps=10, Tmax=0.03))
    chk.set_state(pre_u, 2)
    out = chk.step_device(__import__("torch").as_tensor(a.reshape(B, 4)).cuda())
    assert np.array_equal(finals[:, 0], out["obs"].cpu().numpy())
    assert np.array_equal(rew, out["reward"].cpu().numpy())
    # returned obs is the post-reset observation; counters restarted
    u_new, ts = env.get_state()
    assert (ts == 0).all() and np.array_equal(obs[:, 0], u_new.astype(np.float32))
    assert not np.array_equal(obs, finals)
    obs, rew, term, trunc, info = env.step(a)
    assert (info["step"] == 1).all() and not trunc.any()
    env.close(); chk.close()


def test_partial_truncation_resets_only_finished_envs():
    from model_based_pde_control_b200 import KSVecEnv

    B = 9
    env = KSVecEnv(B, dict(cfg_steps=10, Tmax=0.05), burnin_periods=2)     # 5 steps per episode
    rng = np.random.default_rng(2)
    u0 = rng.uniform(-1, 1, (B, 64))
    ts0 = np.array([4, 0, 4, 1, 2, 4, 3, 0, 4], dtype=np.int32)
    env.set_state(u0, ts0)
    a = rng.uniform(-1, 1, (B, 4)).astype(np.float32)
    obs, rew, term, trunc, info = env.step(a)
    done = ts0 == 4
    assert (trunc == done).all() and (info["_final_observation"] == done).all()
    assert all((info["final_observation"][i] is not None) == bool(done[i]) for i in range(B))
    u, ts = env.get_state()
    assert (ts[done] == 0).all() and (ts[~done] == ts0[~done] + 1).all()
    ref = KSVecEnv(B, dict(cfg_steps=10, Tmax=0.05))
    ref.set_state(u0, ts0)
    out = ref.step_device(__import__("torch").as_tensor(a).cuda())
    u_ref = ref.get_state()[0]
    assert np.array_equal(u[~done], u_ref[~done]), "unfinished envs must be untouched by the masked reset"
    assert not np.array_equal(u[done], u_ref[done])
    for i in np.nonzero(done)[0]:
        assert np.array_equal(info["final_observation"][i][0], u_ref[i])      # float64 state, like the single env
    env.close(); ref.close()


def test_masked_device_reset_and_burnin_is_noop_action():
    import torch
    from model_based_pde_control_b200 import KSVecEnv

    B = 12
    env = KSVecEnv(B, dict(cfg_steps=10), burnin_periods=4)
    rng = np.random.default_rng(3)
    u0 = rng.uniform(-1, 1, (B, 64))
    env.set_state(u0, 17)
    mask = torch.zeros(B, dtype=torch.uint8, device="cuda")
    mask[[1, 5, 6, 11]] = 1
    env.reset_device(seed=99, mask=mask)
    u, ts = env.get_state()
    m = mask.cpu().numpy().astype(bool)
    assert np.array_equal(u[~m], u0[~m]) and (ts[~m] == 17).all() and (ts[m] == 0).all()
    # burn-in == the same number of zero-action periods
    env2 = KSVecEnv(B, dict(cfg_steps=10), burnin_periods=0)
    env2.reset_device(seed=99)                       # same Philox stream, no burn-in
    ic = env2.get_state()[0]
    assert np.abs(ic).max() < 0.4
    for _ in range(4):
        env2.step(np.zeros((B, 4), np.float32))
    assert np.array_equal(env2.get_state()[0][m], u[m])
    env.close(); env2.close()


@pytest.mark.parametrize("stride", [1, 2, 4, 8, 5])
def test_sensor_stride_matches_sensor_transform(stride):
    from model_based_pde_control_b200 import KSVecEnv

    B = 7
    env = KSVecEnv(B, dict(cfg_steps=10), sensor_stride=stride)
    rng = np.random.default_rng(stride)
    env.set_state(rng.uniform(-1, 1, (B, 64)), 0)
    obs, *_ = env.step(rng.uniform(-1, 1, (B, 4)).astype(np.float32))
    u = env.get_state()[0]
    want = u[..., int(stride / 2)::stride].astype(np.float32)         # SensorTransform.__call__ (transforms.py:238)
    assert obs.shape == (B, 1, want.shape[-1]) and np.array_equal(obs[:, 0], want)
    assert env.single_observation_space.shape == (1, want.shape[-1])
    assert env.observation_space.shape == (B, 1, want.shape[-1])
    env.close()


def test_spaces_attributes_and_controller_surface():
    from model_based_pde_control_b200 import KSVecEnv, vector_make

    env = vector_make("KuramotoSivashinskyEnv-v0", num_envs=3)
    assert isinstance(env, KSVecEnv) and env.num_envs == 3
    assert env.single_action_space.shape == (1, 4) and env.action_space.shape == (3, 1, 4)
    assert (env.single_action_space.low == -1).all() and (env.single_action_space.high == 1).all()
    assert env.single_observation_space.shape == (1, 64) and env.observation_space.dtype == np.float32
    # attributes the controller reads (mbrl.py:157,196,215-240,298-300)
    assert (env.cfg_steps, env.dt, env.N, env.L, env.max_episode_steps) == (250, 0.001, 64, 22.0, 400)
    assert env.dx == 22.0 / 64 and env.x.dtype == np.float32 and env.x.shape == (64,)
    assert env.unwrapped is env and env.noop.shape == (1, 4)
    sc = env.scenario
    assert sc["noise"] == 0.1 and sc["lmbda"] == 1.0 and sc["Xi"] == [0.0, 0.25, 0.5, 0.75] and sc["objective"] == "dissipation"
    env.set_state(np.zeros((3, 64)), [0, 4, 400])
    assert np.allclose(env.time, np.array([0, 4, 400]) * 250 * 0.001)
    big = vector_make("KuramotoSivashinskyEnv-v0", num_envs=2, config={"L": 88.0, "N": 256}, Xi=[k / 8 for k in range(8)])
    assert big.single_action_space.shape == (1, 8) and big.launch_info()["lanes_per_env"] in (16, 32)
    env.close(); big.close()


def test_step_protocol_misuse_and_close():
    from model_based_pde_control_b200 import KSVecEnv

    env = KSVecEnv(4, dict(cfg_steps=5))
    with pytest.raises(RuntimeError):
        env.step_wait()
    with pytest.raises(ValueError):
        env.step_async(np.zeros((3, 4), np.float32))
    n0 = env.launch_count
    env.step_async(np.zeros((4, 1, 4), np.float32))
    out = env.step_wait()
    assert len(out) == 5 and env.launch_count == n0 + 1          # one kernel per control period
    env.step(np.zeros((4, 4)))                                    # float64 actions are cast like np.array(action, float32)
    env.close()
    with pytest.raises(RuntimeError):
        env.step(np.zeros((4, 4), np.float32))
    env.close()                                                   # idempotent


def test_reward_func_and_forcing_helpers():
    from model_based_pde_control_b200 import KSVecEnv

    env = KSVecEnv(2)
    rng = np.random.default_rng(1)
    u = rng.uniform(-2, 2, (5, 1, 64))
    r = env.reward_func(u)
    assert r.shape == (5,) and np.allclose(r, -(u ** 2).mean(axis=(1, 2)), rtol=1e-14)
    assert np.isscalar(env.reward_func(u[0])) or np.ndim(env.reward_func(u[0])) == 0
    a = rng.uniform(-1, 1, (1, 4)).astype(np.float32)
    phi = env.forcing(a)
    assert phi.shape == (1, 64) and phi.dtype == np.float32
    d = KSVecEnv(2, reward_mode="dissipation")
    from oracle import ks_numpy as ko
    want = ko.reward_dissipation(u[:, 0], np.tile(phi, (5, 1)), d.dx)
    got = d.reward_func(u, np.tile(phi, (5, 1)))
    assert np.allclose(got, want, rtol=1e-12)
    with pytest.raises(TypeError):
        d.reward_func(u)
    # tensor in -> tensor out on the SAME device, floating dtype kept (surrogates/training.py:214 stacks CPU float32
    # tensors and calls .numpy(); mbrl/world/world.py:170 passes float32 arrays and the ACTION as second argument)
    import torch
    t32 = torch.from_numpy(u[:, 0].astype(np.float32))
    rs = torch.stack([env.reward_func(s, p) for s, p in zip(t32, torch.zeros(5, 64))], dim=0)
    assert rs.device.type == "cpu" and rs.dtype == torch.float32 and rs.shape == (5,)
    assert np.allclose(rs.numpy(), -(u[:, 0].astype(np.float32).astype(np.float64) ** 2).mean(axis=1), rtol=1e-6)
    r32 = np.asarray([env.reward_func(o, a_) for o, a_ in zip(u.astype(np.float32), np.zeros((5, 1, 4), np.float32))], dtype=np.float32)
    assert r32.shape == (5,) and np.allclose(r32, rs.numpy(), rtol=1e-6)
    rc = env.reward_func(t32.cuda())
    assert rc.is_cuda and rc.shape == (5,)
    # env.rhs on NumPy rows as training.py:228 calls it: (rhs, (ux, uxx, uxxxx)), each shaped like the input
    rhs, (ux, uxx, uxxxx) = env.rhs(u[0].astype(np.float32), phi)
    assert rhs.shape == ux.shape == uxx.shape == uxxxx.shape == (1, 64)
    env.close(); d.close()


def test_reset_pool_mode_injects_preburned_states():
    """reset_mode="pool": the auto-reset takes a batch that the side stream burned in ahead of time
    (same IC distribution, same 800-period burn-in, each state used once)."""
    import torch
    from model_based_pde_control_b200 import KSVecEnv

    B = 16
    env = KSVecEnv(B, dict(cfg_steps=10, Tmax=0.03), burnin_periods=6, reset_mode="pool", pool_slots=2, ic="device")
    env.reset(seed=3)
    a = np.zeros((B, 4), np.float32)
    seen = []
    for ep in range(4):                                    # 4 episode ends -> pool slots are recycled twice
        for k in range(3):
            obs, rew, term, trunc, info = env.step(a)
        assert trunc.all() and (info["step"] == 3).all() and info["_final_observation"].all()
        u, ts = env.get_state()
        assert (ts == 0).all() and np.array_equal(obs[:, 0], u.astype(np.float32))
        seen.append(u.copy())
    assert env._pool is not None and env._pool.refills >= 2 + 4
    for i in range(4):
        for j in range(i):
            assert not np.array_equal(seen[i], seen[j]), "every pooled batch must be fresh"
    # a pooled state is exactly what reset_device(seed) + burn-in produces on an identical env
    chk = KSVecEnv(B, dict(cfg_steps=10, Tmax=0.03), burnin_periods=6, ic="device")
    pool = env._pool
    first_seed = (pool._next_seed - pool.refills * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
    chk.reset_device(seed=first_seed)
    torch.cuda.synchronize()
    assert np.array_equal(chk.get_state()[0], seen[0])
    # partial truncation goes through the same path with a mask
    env.set_state(seen[0], np.array([2] * 8 + [0] * 8, dtype=np.int32))
    obs, rew, term, trunc, info = env.step(a)
    u, ts = env.get_state()
    assert trunc[:8].all() and not trunc[8:].any() and (ts[:8] == 0).all() and (ts[8:] == 1).all()
    env.close(); chk.close()
