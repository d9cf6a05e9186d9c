"""Golden vectors for the INTENDED reward (``reward_mode="dissipation"``), produced by executing the reference's own code.

    python tests/golden/make_golden_dissipation.py

``KuramotoSivashinskyEnv(objective="")`` selects the ``dissipation`` closure (``kuramoto.py:67-72``), but
``env.step`` then raises ``TypeError``: ``FuncTransform`` hands the closure torch tensors and ``rhs`` mixes them with
NumPy arrays.  The closure itself -- ``env.reward_func.transf`` -- and ``env.rhs`` are intact, so this
script calls THEM with the NumPy arrays ``step`` holds at that point:

* ``reward[i] = closure(U[i], PHI[i])`` for 32 developed / random states (one value per state);
* one control period composed exactly as ``step`` composes it (``kuramoto.py:83-96``: reward of the pre-step state,
  four ``env.rhs`` calls, RK4 update, mean over the sub-steps), every arithmetic call being the reference's own
  function; default grid (250 sub-steps) and the large domain (N = 256, 8 jets, 25 sub-steps).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle.ref_loader import make_reference_env  # noqa: E402


def period(env, closure, u, phi):
    """``KuramotoSivashinskyEnv.step`` (kuramoto.py:83-96) with NumPy arrays all the way."""
    reward = 0.0
    for _ in range(env.cfg_steps):
        reward += closure(u, phi)
        k1, _ = env.rhs(u, phi)
        k2, _ = env.rhs(u + env.dt * k1 / 2.0, phi)
        k3, _ = env.rhs(u + env.dt * k2 / 2.0, phi)
        k4, _ = env.rhs(u + env.dt * k3, phi)
        u = u + env.dt * (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0
    return u, reward / env.cfg_steps


def main():
    env = make_reference_env(objective="")
    closure = env.reward_func.transf
    assert closure.__name__ == "dissipation"
    rng = np.random.default_rng(31)
    ua = np.load(os.path.join(HERE, "attractor_default_random.npz"))["u0"]
    U = np.concatenate([ua[None] * rng.uniform(0.5, 1.5, (16, 1)) + rng.uniform(-0.2, 0.2, (16, 64)),
                        rng.uniform(-2.0, 2.0, (16, 64))])
    A = rng.uniform(-1, 1, (32, 4)).astype(np.float32)
    PHI = np.stack([np.squeeze(env.forcing(a.reshape(1, 4))) for a in A]).astype(np.float32)
    reward = np.array([closure(u, p) for u, p in zip(U, PHI)])
    a1 = rng.uniform(-1, 1, (1, 4)).astype(np.float32)
    phi1 = np.squeeze(env.forcing(a1))
    u1, r1 = period(env, closure, ua.copy(), phi1)

    Xi8 = [k / 8 for k in range(8)]
    envl = make_reference_env(Xi=Xi8, L=88.0, N=256, cfg_steps=25, objective="")
    ual = np.load(os.path.join(HERE, "attractor_large_random.npz"))["u0"]
    a8 = rng.uniform(-1, 1, (1, 8)).astype(np.float32)
    phi8 = np.squeeze(envl.forcing(a8))
    u8, r8 = period(envl, envl.reward_func.transf, ual.copy(), phi8)

    np.savez_compressed(os.path.join(HERE, "dissipation_kat.npz"), U=U, A=A, PHI=PHI, reward=reward,
                        u0=ua, a1=a1, phi1=phi1.astype(np.float32), u1=u1, r1=r1,
                        u0_large=ual, a8=a8, phi8=phi8.astype(np.float32), u1_large=u8, r1_large=r8, cfg_steps_large=25)
    print(f"dissipation_kat: reward[:3]={reward[:3]}, period reward {r1:.17g}, |u1|={np.linalg.norm(u1):.17g}; "
          f"large: {r8:.17g}, |u1|={np.linalg.norm(u8):.17g}")


if __name__ == "__main__":
    main()
