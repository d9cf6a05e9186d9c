"""Long-horizon reference statistics for the 1 % gate (BASELINE.json north star: "energy spectrum
and mean dissipation must agree within 1%").

    python tests/golden/make_stats.py default 32768     # ~17 min on 8 cores
    python tests/golden/make_stats.py large 4096

(optionally with KS_ORACLE_LIB pointing at a -march=native build of oracle/ks_oracle.c)

Protocol: every env starts from u ~ U(-0.4,0.4)^N, runs the reference's 800
no-op burn-in periods, then one 400-period episode with i.i.d. actions ~ U(-1,1)^J; statistics are
taken over the 400 period-end states of all envs:
  spectrum[k]  = < |rfft(u)_k|^2 > / N^2
  dissipation  = < mean(uxx^2) + mean(ux^2) + mean(u*phi) >      (ux = upwind d/dx of u^2, as rhs())
  mean_u2      = < mean(u^2) >
The trajectories come from the plain-C oracle (oracle/ks_oracle.c), which tests/test_oracle.py
pins to the reference's own outputs at 1e-13 per control period; the unmodified reference needs
~50 ms per env-period in Python, i.e. ~34 core-hours for this many episodes.  Per-env means are
kept so that the standard error of every statistic is recorded next to it.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ks_c, ks_numpy as ko  # noqa: E402

CONFIGS = {
    "default": dict(L=22.0, N=64, Xi=(0.0, 0.25, 0.5, 0.75)),
    "large": dict(L=88.0, N=256, Xi=tuple(k / 8 for k in range(8))),
}


def main():
    name = sys.argv[1]
    E = int(sys.argv[2])
    cfg = ko.KSConfig(**CONFIGS[name])
    F = ko.forcing_matrix(cfg)
    rng = np.random.default_rng(20261018)
    u = rng.uniform(-0.4, 0.4, (E, cfg.N))
    t0 = time.time()
    u, _, _ = ks_c.rollout(cfg, u, None, F, K=cfg.burnin_periods, want_obs=False)
    print(f"burn-in done in {time.time() - t0:.0f} s", flush=True)
    K = cfg.max_episode_steps
    spec = np.zeros((E, cfg.N // 2 + 1))
    diss = np.zeros(E)
    u2 = np.zeros(E)
    rew = np.zeros(E)
    for k in range(K):
        a = rng.uniform(-1, 1, (E, cfg.J)).astype(np.float32)
        phi = ks_c.forcing(a, F)            # == ko.forcing bit for bit (tests/test_oracle.py), much faster
        u, r = ks_c.step(cfg, u, phi)
        spec += ko.energy_spectrum(u)
        diss += ko.dissipation_rate(u, phi, cfg.dx)
        u2 += (u * u).mean(axis=-1)
        rew += r
        if k % 50 == 0:
            print(f"period {k} / {K}  ({time.time() - t0:.0f} s)", flush=True)
    spec /= K; diss /= K; u2 /= K; rew /= K

    def sem(x):
        return x.std(axis=0, ddof=1) / np.sqrt(E)

    out = dict(
        L=cfg.L, N=cfg.N, Xi=np.asarray(cfg.Xi), n_envs=E, periods=K, burnin_periods=cfg.burnin_periods,
        spectrum=spec.mean(0), spectrum_sem=sem(spec), dissipation=diss.mean(), dissipation_sem=sem(diss),
        mean_u2=u2.mean(), mean_u2_sem=sem(u2), mean_reward=rew.mean(), mean_reward_sem=sem(rew),
        generator="oracle/ks_oracle.c via tests/golden/make_stats.py, default_rng(20261018)",
    )
    np.savez_compressed(os.path.join(HERE, f"stats_{name}.npz"), **out)
    print({k: (float(v) if np.ndim(v) == 0 and not isinstance(v, str) else None) for k, v in out.items()})


if __name__ == "__main__":
    main()
