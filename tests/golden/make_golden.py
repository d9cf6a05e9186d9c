"""Regenerate the golden fixtures in this directory by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container (needs ``/root/reference``)::

    python tests/golden/make_golden.py            # all cases (about one minute)

Every array below is an output of the reference's own code (``pdegym/kuramoto/kuramoto.py``,
``pdegym/common/transforms.py``), loaded where it lies through ``oracle/ref_loader.py``; states
are injected with ``env.u = u0; env.timestep = t0`` because ``reset()`` is not reproducible from
the reference's callers.  Library versions are recorded in
``manifest.json`` -- agreement across versions is expected to ~1e-13, not bitwise.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

import torch  # noqa: E402

from oracle.ref_loader import make_reference_env  # noqa: E402

torch.set_num_threads(1)


def run_case(name, config, Xi, u0, actions, t0=0, action_shape="2d"):
    """Step the reference through ``actions [K,J]`` from the injected state; record everything."""
    env = make_reference_env(Xi=Xi, **config)
    env.u = np.array(u0, dtype=np.float64)
    env.timestep = int(t0)
    J = len(env.Xi)
    us, rewards, truncs, steps, phis = [], [], [], [], []
    for a in actions:
        a = np.asarray(a, dtype=np.float32)
        act = a.reshape(1, J) if action_shape == "2d" else a.reshape(J)
        phis.append(np.squeeze(env.forcing(np.array(act, dtype=np.float32))).copy())
        obs, r, term, trunc, info = env.step(act)
        assert term is False and obs.shape == (1, env.N) and obs.dtype == np.float64
        us.append(obs[0].copy())
        rewards.append(float(r))
        truncs.append(bool(trunc))
        steps.append(int(info["step"]))
    out = dict(
        L=float(env.L), N=int(env.N), cfg_steps=int(env.cfg_steps), dt=float(env.dt),
        sigma=float(env.sigma), Tmax=float(env.Tmax), Xi=np.asarray(env.Xi, dtype=np.float64),
        x=env.x.copy(), F=env.forcing.forcing.numpy().copy(),
        max_episode_steps=int(env.max_episode_steps), t0=int(t0),
        u0=np.asarray(u0, dtype=np.float64), actions=np.asarray(actions, dtype=np.float32),
        phi=np.asarray(phis, dtype=np.float32), u=np.asarray(us), reward=np.asarray(rewards),
        truncated=np.asarray(truncs), step=np.asarray(steps, dtype=np.int64),
    )
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: K={len(actions)} |u_K|={np.linalg.norm(us[-1]):.17g} r_K={rewards[-1]:.17g}")
    return out


def attractor_state(config, Xi, seed, periods):
    """A developed state: ``periods`` no-op periods from the reference's own IC distribution."""
    env = make_reference_env(Xi=Xi, **config)
    env.u = np.random.default_rng(seed).uniform(-0.4, 0.4, env.N)
    env.timestep = 0
    zero = np.zeros((1, len(env.Xi)), dtype=np.float32)
    for _ in range(periods):
        env.step(zero)
    return env.u.copy()


def main():
    default = {}
    Xi4 = [0, 0.25, 0.5, 0.75]
    Xi8 = [k / 8 for k in range(8)]
    large = dict(L=88.0, N=256)

    # KAT 1: one period from a small random state
    u0 = np.random.default_rng(1000).uniform(-0.4, 0.4, 64)
    k1 = run_case("kat1_default_1period", default, Xi4, u0, [[0.3, -0.5, 1.0, -1.0]])
    # KAT 2: ten more periods with random actions
    rng = np.random.default_rng(7)
    acts = [rng.uniform(-1, 1, (1, 4)).astype(np.float32)[0] for _ in range(10)]
    run_case("kat2_default_10periods", default, Xi4, k1["u"][-1], acts, t0=1)
    # KAT 3: large domain, 8 jets
    u0 = np.random.default_rng(2000).uniform(-0.4, 0.4, 256)
    a = np.random.default_rng(8).uniform(-1, 1, (1, 8)).astype(np.float32)
    run_case("kat3_large_1period", large, Xi8, u0, a)

    # developed (attractor) states: many sign changes -> exercises the upwind switch
    ua = attractor_state(default, Xi4, seed=11, periods=120)
    rng = np.random.default_rng(21)
    run_case("attractor_default_random", default, Xi4, ua,
             rng.uniform(-1, 1, (4, 4)).astype(np.float32), t0=120)
    run_case("attractor_default_zero_action", default, Xi4, ua, np.zeros((2, 4), np.float32), t0=10)
    run_case("attractor_default_saturated", default, Xi4, ua,
             np.array([[1, -1, 1, -1], [-1, -1, 1, 1], [1, 1, 1, 1]], np.float32), t0=0)
    run_case("attractor_default_action1d", default, Xi4, ua,
             rng.uniform(-1, 1, (2, 4)).astype(np.float32), action_shape="1d")
    # truncation edge: steps 399 -> 400 -> 401
    run_case("truncation_edge", default, Xi4, ua, rng.uniform(-1, 1, (3, 4)).astype(np.float32), t0=398)
    ual = attractor_state(large, Xi8, seed=12, periods=60)
    run_case("attractor_large_random", large, Xi8, ual, rng.uniform(-1, 1, (2, 8)).astype(np.float32))
    # another lane layout: N=128, L=44, 4 jets; and a short control period
    um = attractor_state(dict(L=44.0, N=128), Xi4, seed=13, periods=60)
    run_case("attractor_n128_random", dict(L=44.0, N=128), Xi4, um,
             rng.uniform(-1, 1, (2, 4)).astype(np.float32))
    run_case("short_period_cfg10", dict(cfg_steps=10), Xi4, ua, rng.uniform(-1, 1, (3, 4)).astype(np.float32))
    # non-power-of-two grid
    un = attractor_state(dict(L=33.0, N=96), Xi4, seed=14, periods=60)
    run_case("attractor_n96_random", dict(L=33.0, N=96), Xi4, un, rng.uniform(-1, 1, (2, 4)).astype(np.float32))

    # rhs / derivative triples on a developed state (kuramoto.py:118-129)
    env = make_reference_env()
    phi = np.squeeze(env.forcing(np.array([[0.7, -0.2, 0.1, -0.9]], np.float32)))
    r, (ux, uxx, uxxxx) = env.rhs(ua, phi)
    np.savez_compressed(os.path.join(HERE, "rhs_default.npz"), u=ua, phi=phi, rhs=r, ux=ux, uxx=uxx,
                        uxxxx=uxxxx, dx=env.dx)

    # reset(): IC stream only (cheap) ...
    ics = {}
    for seed in (0, 1, 5, 123):
        np.random.seed(seed)
        ics[f"seed{seed}"] = np.random.uniform(-0.4, 0.4, size=64)
    np.savez_compressed(os.path.join(HERE, "reset_ic.npz"), **ics)
    # ... and one full reference reset (800 burn-in periods, ~30 s)
    env = make_reference_env()
    obs, info = env.reset(seed=5, return_info=True)
    np.savez_compressed(os.path.join(HERE, "reset_full_seed5.npz"), u=obs[0], step=int(info["step"]),
                        u0=ics["seed5"])
    print(f"reset_full_seed5: |u|={np.linalg.norm(obs):.17g} step={info['step']}")

    # phi = a @ F known answers, incl. batched 2-D calls
    env = make_reference_env()
    A = np.random.default_rng(99).uniform(-1, 1, (256, 4)).astype(np.float32)
    phi_rows = np.stack([np.squeeze(env.forcing(a.reshape(1, 4))) for a in A])
    envl = make_reference_env(Xi=Xi8, **large)
    A8 = np.random.default_rng(98).uniform(-1, 1, (64, 8)).astype(np.float32)
    phi8 = np.stack([np.squeeze(envl.forcing(a.reshape(1, 8))) for a in A8])
    np.savez_compressed(os.path.join(HERE, "forcing_kat.npz"), A=A, phi=phi_rows, F=env.forcing.forcing.numpy(),
                        A8=A8, phi8=phi8, F8=envl.forcing.forcing.numpy())

    import scipy
    manifest = dict(
        generated_by="tests/golden/make_golden.py (reference executed via oracle/ref_loader.py)",
        numpy=np.__version__, scipy=scipy.__version__, torch=torch.__version__,
        python=sys.version.split()[0],
        reference_pins="numpy 1.24.4 / scipy 1.10.1 / torch 1.10.1 (poetry.lock) -- not installable here",
    )
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
