"""Single-env ``gym.Env`` facade (``KSEnv`` / ``make``) -- SURVEY 8 row a-8.

The reference's id resolves to ``TimeLimit(KuramotoSivashinskyEnv(**config))``
(``pdegym/kuramoto/__init__.py:8-12,26-31``); ``pdecontrol/surrogates/evaluation/generate.py:23-38`` steps
that object with ``(1,J)`` actions.  The loop below is ``generate.py:28-38`` verbatim except that the
actions come from the golden fixture (so that the result can be compared) instead of
``env.action_space.sample()``.
"""
import numpy as np
import pytest

from ks_testutil import load_golden, rel_l2

pytestmark = pytest.mark.gpu


def test_generate_loop_reproduces_the_reference_trajectory():
    from model_based_pde_control_b200 import make

    g = load_golden("kat2_default_10periods")
    env = make({})                                            # generate.py:23  gym.make(env, config=config, new_step_api=True)
    assert env.observation_space.shape == (1, 64) and env.action_space.shape == (1, 4)
    obs = env.reset(seed=0, burnin_periods=0)                 # (the fixture injects its own start state)
    assert obs.shape == (1, 64) and obs.dtype == np.float64
    env.unwrapped.u = g["u0"]
    env.unwrapped.timestep = int(g["t0"])
    terminated, truncated, episode = False, False, []
    k = 0
    while not terminated and not truncated and k < len(g["actions"]):
        action = g["actions"][k].reshape(1, 4)                # generate.py:33 samples from action_space: shape (1,4) f32
        nxt, rew, terminated, truncated, info = env.step(action)
        episode.append((obs, action, nxt, rew, terminated, truncated))
        obs = nxt
        # types exactly as the reference returns them (kuramoto.py:92-98)
        assert nxt.shape == (1, 64) and nxt.dtype == np.float64 and isinstance(rew, float)
        assert terminated is False and isinstance(truncated, bool) and info == {"step": int(g["step"][k])}
        assert rel_l2(nxt[0], g["u"][k]) <= 1e-10 and abs(rew - g["reward"][k]) <= 1e-10 * abs(g["reward"][k])
        assert truncated == bool(g["truncated"][k])
        k += 1
    assert k == 10
    # generate.py:41-50: the per-episode tuples stack into float32 arrays of these shapes
    o, a, n, r, tm, tr = (np.array(x) for x in zip(*episode))
    assert o.shape == (10, 1, 64) and a.shape == (10, 1, 4) and n.shape == (10, 1, 64) and r.shape == (10,)
    assert np.isclose(env.unwrapped.time, (int(g["t0"]) + 10) * 0.25)
    env.close()


def test_reset_matches_np_random_seed_stream_and_runs_the_burn_in():
    from model_based_pde_control_b200 import KSEnv

    ic = load_golden("reset_ic")
    env = KSEnv()
    obs, info = env.reset(seed=5, return_info=True, burnin_periods=0)
    assert np.array_equal(obs[0], ic["seed5"]) and info == {"step": 0}      # np.random.seed(5); uniform(-0.4, 0.4, 64)
    full = load_golden("reset_full_seed5")                                   # the reference's 800-period reset(seed=5)
    obs = env.reset(seed=5)
    assert env.timestep == 0 and rel_l2(obs[0], full["u"]) <= 1e-4           # chaos-limited
    # 1-D actions are accepted like np.array(action, float32) (kuramoto.py:79; reset steps with a 4-list, :109)
    nxt, rew, term, trunc, info = env.step([0.0, 0.0, 0.0, 0.0])
    assert nxt.shape == (1, 64) and info["step"] == 1 and not trunc
    with pytest.raises(ValueError):
        env.step(np.zeros((1, 5), np.float32))
    env.close()
    with pytest.raises(RuntimeError):
        env.step([0.0] * 4)


def test_time_limit_truncates_like_the_reference_factory():
    from model_based_pde_control_b200 import make

    env = make(dict(cfg_steps=5, Tmax=0.02), burnin_periods=2)             # ceil(0.02 / (0.001 * 5)) = 4 steps
    assert env.unwrapped.max_episode_steps == 4
    env.reset(seed=1)
    flags = [env.step(env.action_space.sample())[3] for _ in range(4)]
    assert flags == [False, False, False, True]
    # no auto-reset in the single env: a further step just continues (timestep 5), as in the reference
    *_, info = env.step(env.action_space.sample())
    assert info["step"] == 5
    env.reset(seed=2)
    assert env.unwrapped.timestep == 0 and env.step(env.action_space.sample())[3] is False
    env.close()


def test_large_domain_facade_and_overflow_error():
    from model_based_pde_control_b200 import KSEnv

    g = load_golden("kat3_large_1period")
    env = KSEnv({"L": 88.0, "N": 256}, Xi=[k / 8 for k in range(8)])
    assert env.action_space.shape == (1, 8) and env.observation_space.shape == (1, 256)
    env.u, env.timestep = g["u0"], int(g["t0"])
    nxt, rew, *_ = env.step(g["actions"][0])
    assert rel_l2(nxt[0], g["u"][0]) <= 1e-10 and abs(rew - g["reward"][0]) <= 1e-10 * abs(g["reward"][0])
    env.u = np.full(256, 1e200)
    with pytest.raises(FloatingPointError):
        env.step(np.zeros((1, 8), np.float32))
    env.close()
