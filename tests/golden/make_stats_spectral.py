"""Long-horizon statistics of the spectral ETDRK4 solver from its NumPy oracle (oracle/ks_etdrk4.py).

    python tests/golden/make_stats_spectral.py 8192            # default grid, ~5 min on 8 cores
    python tests/golden/make_stats_spectral.py 1024 large      # BASELINE configs[3]: N=256, L=88, 8 jets

Same protocol and estimators as tests/golden/make_stats.py (the FD-RK4 / reference fixture):
U(-0.4,0.4)^N initial conditions, 800 no-op control periods of burn-in, one 400-period episode with
i.i.d. actions ~ U(-1,1)^4; spectrum / dissipation / mean u^2 over the 400 period-end states, mean
reward over the 400 periods.  dt = 0.025 x 10 ETDRK4 steps per control period, 2/3 dealiasing.

The result is stored next to the reference-scheme fixture (stats_default.npz) together with the
relative differences between the two discretisations -- the spectral solver integrates the PDE
(whose mean mode is driven only by the mean forcing), the reference's upwind finite differences
at dx = 0.34 do not conserve the mean and damp the low modes differently, so the two attractors
differ by several per cent in the low wavenumber bins (recorded here, asserted nowhere as equal).
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ks_c, ks_etdrk4 as ke, ks_numpy as ko  # noqa: E402

DT, S = 0.025, 10


def config(name):
    if name == "large":
        return ko.KSConfig(L=88.0, N=256, Xi=tuple(k / 8 for k in range(8)))
    return ko.KSConfig()


def worker(args):
    seed, E, name = args
    cfg = config(name)
    F = ko.forcing_matrix(cfg)
    rng = np.random.default_rng(seed)
    c = ke.etd_coefficients(cfg.N, cfg.L, DT)
    u = rng.uniform(-0.4, 0.4, (E, cfg.N))
    zero = np.zeros((E, cfg.N), np.float32)
    for _ in range(cfg.burnin_periods):
        u, _ = ke.step(u, zero, cfg.N, cfg.L, DT, S, coef=c)
    K = cfg.max_episode_steps
    spec = np.zeros((E, cfg.N // 2 + 1)); diss = np.zeros(E); u2 = np.zeros(E); rew = np.zeros(E)
    for _ in range(K):
        a = rng.uniform(-1, 1, (E, cfg.J)).astype(np.float32)
        phi = ks_c.forcing(a, F)
        u, r = ke.step(u, phi, cfg.N, cfg.L, DT, S, coef=c)
        spec += ko.energy_spectrum(u); diss += ko.dissipation_rate(u, phi, cfg.dx); u2 += (u * u).mean(-1); rew += r
    return spec / K, diss / K, u2 / K, rew / K


def main():
    E = int(sys.argv[1])
    name = sys.argv[2] if len(sys.argv) > 2 else "default"
    P = os.cpu_count() or 1
    per = (E + P - 1) // P
    t0 = time.time()
    with mp.get_context("spawn").Pool(P) as pool:
        res = pool.map(worker, [(20261018 + i, per, name) for i in range(P)])
    spec = np.concatenate([r[0] for r in res]); diss = np.concatenate([r[1] for r in res])
    u2 = np.concatenate([r[2] for r in res]); rew = np.concatenate([r[3] for r in res])
    n = len(diss)

    def sem(x):
        return x.std(axis=0, ddof=1) / np.sqrt(n)

    cfg = config(name)
    out = dict(L=cfg.L, N=cfg.N, Xi=np.asarray(cfg.Xi), n_envs=n, periods=cfg.max_episode_steps,
               burnin_periods=cfg.burnin_periods, dt=DT, cfg_steps=S, dealias=True,
               spectrum=spec.mean(0), spectrum_sem=sem(spec), dissipation=diss.mean(), dissipation_sem=sem(diss),
               mean_u2=u2.mean(), mean_u2_sem=sem(u2), mean_reward=rew.mean(), mean_reward_sem=sem(rew),
               generator="oracle/ks_etdrk4.py via tests/golden/make_stats_spectral.py")
    ref = np.load(os.path.join(HERE, f"stats_{name}.npz"))
    out["rel_diff_vs_reference_scheme_spectrum"] = (out["spectrum"] - ref["spectrum"]) / ref["spectrum"]
    out["rel_diff_vs_reference_scheme_dissipation"] = out["dissipation"] / float(ref["dissipation"]) - 1
    out["rel_diff_vs_reference_scheme_mean_u2"] = out["mean_u2"] / float(ref["mean_u2"]) - 1
    np.savez_compressed(os.path.join(HERE, f"stats_spectral_{name}.npz"), **out)
    print(f"{n} envs in {time.time() - t0:.0f} s; mean u^2 {out['mean_u2']:.5f} +- {out['mean_u2_sem']:.5f}; "
          f"dissipation {out['dissipation']:.5f}; vs reference scheme: spectrum[:6] "
          f"{np.round(out['rel_diff_vs_reference_scheme_spectrum'][:6], 4)}, dissipation "
          f"{out['rel_diff_vs_reference_scheme_dissipation']:.4f}, mean u^2 {out['rel_diff_vs_reference_scheme_mean_u2']:.4f}")


if __name__ == "__main__":
    main()
