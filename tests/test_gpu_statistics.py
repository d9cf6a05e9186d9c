"""Long-horizon gate of BASELINE.json's north star: "Long-horizon chaotic trajectories must match in
statistics: the energy spectrum and mean dissipation must agree within 1%".

Reference statistics: ``tests/golden/stats_*.npz`` (``tests/golden/make_stats.py``: reference
burn-in + one 400-period random-action episode per env, trajectories from the C oracle that is
pinned to the reference at 1e-13 per period).  GPU side: the same protocol with 16 384 envs through
the C ABI -- device Philox initial conditions, ONE 800-period burn-in launch, 400 ``ks_step``
periods with i.i.d. U(-1,1) actions -- statistics accumulated on the GPU from the fp64 state.
"""
import os

import numpy as np
import pytest

from ks_testutil import GOLDEN

pytestmark = pytest.mark.gpu

TOL = 0.01     # north star: within 1 %


def gpu_statistics(name, B, precision="f64", solver="fd_rk4"):
    import torch
    from model_based_pde_control_b200 import KSVecEnv

    g = np.load(os.path.join(GOLDEN, f"stats_{name}.npz"))
    cfg = dict(L=float(g["L"]), N=int(g["N"]))
    if solver != "fd_rk4":
        cfg.update(dt=float(g["dt"]), cfg_steps=int(g["cfg_steps"]))
    env = KSVecEnv(B, cfg, Xi=list(g["Xi"]), ic="device", precision=precision, solver=solver)
    N, J = env.N, env.J
    env.reset_device(seed=20261018)                     # IC + 800 no-op periods, one launch
    gen = torch.Generator(device="cuda").manual_seed(7)
    F = torch.as_tensor(env.forcing.matrix()).cuda()
    spec = torch.zeros(N // 2 + 1, dtype=torch.float64, device="cuda")
    diss = torch.zeros((), dtype=torch.float64, device="cuda")
    u2 = torch.zeros((), dtype=torch.float64, device="cuda")
    rew = torch.zeros((), dtype=torch.float64, device="cuda")
    K = int(g["periods"])
    assert env.max_episode_steps == K
    for _ in range(K):
        a = torch.rand((B, J), generator=gen, device="cuda", dtype=torch.float32) * 2 - 1
        out = env.step_device(a)
        u, _ = env.get_state_device()
        phi = a @ F
        d = env.evaluate(u, phi, want=("ux", "uxx"))
        spec += (torch.fft.rfft(u, dim=-1).abs() ** 2).mean(0) / N ** 2
        diss += (d["uxx"] ** 2).mean() + (d["ux"] ** 2).mean() + (u * phi.double()).mean()
        u2 += (u * u).mean()
        rew += out["reward"].mean()
    assert not env.nonfinite().any()
    env.close()
    return g, (spec / K).cpu().numpy(), float(diss / K), float(u2 / K), float(rew / K)


@pytest.mark.parametrize("name,B", [("default", 16384), ("large", 4096)])
def test_spectrum_and_dissipation_within_one_percent(name, B):
    if not os.path.exists(os.path.join(GOLDEN, f"stats_{name}.npz")):
        pytest.skip(f"stats_{name}.npz not generated")
    g, spec, diss, u2, rew = gpu_statistics(name, B)
    ref = g["spectrum"]
    big = ref > 0.01 * ref.sum()                     # wavenumbers holding more than 1 % of the energy
    assert big.sum() >= 3
    rel = np.abs(spec[big] - ref[big]) / ref[big]
    # per bin: 1 %, widened to 4 standard errors of the REFERENCE statistic where the fixture's own
    # sampling error is larger than that (stats_large: 1024 reference episodes -> ~1 % per bin)
    tol = np.maximum(TOL, 4.0 * g["spectrum_sem"][big] / ref[big])
    print(f"{name}: spectrum rel. dev (bins {np.nonzero(big)[0].tolist()}): {np.round(rel, 4).tolist()}; "
          f"dissipation {diss:.5f} vs {float(g['dissipation']):.5f}; mean u^2 {u2:.5f} vs {float(g['mean_u2']):.5f}")
    assert (rel <= tol).all(), (rel / tol).max()
    # the energy in those bins taken together (sampling error averages out): strictly 1 %
    assert abs(spec[big].sum() / ref[big].sum() - 1) <= TOL
    assert abs(diss - float(g["dissipation"])) <= TOL * abs(float(g["dissipation"]))
    assert abs(u2 - float(g["mean_u2"])) <= TOL * float(g["mean_u2"])
    assert abs(rew - float(g["mean_reward"])) <= TOL * abs(float(g["mean_reward"]))
    # total energy (Parseval) ties spectrum and mean u^2 together
    assert abs((2 * spec.sum() - spec[0] - spec[-1]) - u2) <= 1e-9 * u2


def test_fp32_mode_statistics():
    if not os.path.exists(os.path.join(GOLDEN, "stats_default.npz")):
        pytest.skip("stats_default.npz not generated")
    g, spec, diss, u2, rew = gpu_statistics("default", 16384, precision="f32")
    ref = g["spectrum"]
    big = ref > 0.01 * ref.sum()
    assert (np.abs(spec[big] - ref[big]) / ref[big]).max() <= TOL
    assert abs(diss - float(g["dissipation"])) <= TOL * abs(float(g["dissipation"]))


def test_fp32_mode_statistics_large_domain():
    """BASELINE configs[3] in the optional fp32 mode: N=256, L=88, 8 jets against the reference-scheme
    fixture (1024 reference episodes; per bin 1 % or 4 standard errors of the fixture, whichever is
    larger -- the same widening as the fp64 test above; the totals strictly 1 %)."""
    if not os.path.exists(os.path.join(GOLDEN, "stats_large.npz")):
        pytest.skip("stats_large.npz not generated")
    g, spec, diss, u2, rew = gpu_statistics("large", 4096, precision="f32")
    ref = g["spectrum"]
    big = ref > 0.01 * ref.sum()
    rel = np.abs(spec[big] - ref[big]) / ref[big]
    tol = np.maximum(TOL, 4.0 * g["spectrum_sem"][big] / ref[big])
    print(f"large fp32: spectrum rel. dev {np.round(rel, 4).tolist()}; dissipation {diss:.5f} vs {float(g['dissipation']):.5f}; "
          f"mean u^2 {u2:.5f} vs {float(g['mean_u2']):.5f}")
    assert (rel <= tol).all(), (rel / tol).max()
    assert abs(spec[big].sum() / ref[big].sum() - 1) <= TOL
    assert abs(diss - float(g["dissipation"])) <= TOL * abs(float(g["dissipation"]))
    assert abs(u2 - float(g["mean_u2"])) <= TOL * float(g["mean_u2"])
    assert abs(rew - float(g["mean_reward"])) <= TOL * abs(float(g["mean_reward"]))


@pytest.mark.parametrize("name,B", [("spectral_default", 32768), ("spectral_large", 8192)])
def test_spectral_solver_statistics_vs_its_oracle(name, B):
    """The ETDRK4 solver against the statistics of its own NumPy oracle (tests/golden/
    make_stats_spectral.py; the reference has no spectral solver).  Same 1 % gate, widened only by
    the fixture's recorded standard error where that is larger (the unforced mean mode performs a
    random walk under the jets' mean forcing, so bin 0 is noisy)."""
    if not os.path.exists(os.path.join(GOLDEN, f"stats_{name}.npz")):
        pytest.skip(f"stats_{name}.npz not generated")
    g, spec, diss, u2, rew = gpu_statistics(name, B, solver="etdrk4")
    ref, sem = g["spectrum"], g["spectrum_sem"]
    big = ref > 0.01 * ref.sum()
    tol = np.maximum(TOL * ref[big], 4.0 * sem[big])
    dev = np.abs(spec[big] - ref[big])
    print(f"{name}: spectrum rel. dev {np.round(dev / ref[big], 4).tolist()} (tol {np.round(tol / ref[big], 4).tolist()}); "
          f"dissipation {diss:.5f} vs {float(g['dissipation']):.5f}; mean u^2 {u2:.5f} vs {float(g['mean_u2']):.5f}")
    assert (dev <= tol).all()
    assert abs(diss - float(g["dissipation"])) <= max(TOL * abs(float(g["dissipation"])), 4 * float(g["dissipation_sem"]))
    assert abs(u2 - float(g["mean_u2"])) <= max(TOL * float(g["mean_u2"]), 4 * float(g["mean_u2_sem"]))
    assert abs(rew - float(g["mean_reward"])) <= max(TOL * abs(float(g["mean_reward"])), 4 * float(g["mean_reward_sem"]))
    # and the recorded distance between the two discretisations is what the fixture says it is:
    # a few per cent in the low wavenumbers, i.e. NOT within the 1 % gate of the reference scheme
    fd = np.load(os.path.join(GOLDEN, "stats_" + name.split("_", 1)[1] + ".npz"))
    assert abs(u2 / float(fd["mean_u2"]) - 1) < 0.05
