"""CPU checks of the spectral ETDRK4 oracle (``oracle/ks_etdrk4.py``), the checker of the CUDA
spectral solver.  The reference contains no spectral code to pin it against, so it is anchored on
(1) fourth-order self-convergence in dt, (2) agreement with the reference's FD-RK4 scheme up to that
scheme's spatial truncation error, which shrinks at the FD order when the grid is refined (C oracle
pinned to the reference, ``tests/test_oracle.py``), and (3) exactness of the linear propagator."""
import numpy as np

from ks_testutil import load_golden, rel_l2
from oracle import ks_c, ks_etdrk4 as ke, ks_numpy as ko


def test_fourth_order_convergence_in_dt():
    g = load_golden("attractor_default_random")
    u0, phi = g["u0"], g["phi"][0]
    fine, _ = ke.step(u0, phi, 64, 22.0, 0.25 / 500, 500)
    err = [rel_l2(ke.step(u0, phi, 64, 22.0, 0.25 / s, s)[0], fine) for s in (10, 20, 40)]
    assert err[0] < 1e-7 and err[1] < 1e-8
    assert 10 < err[0] / err[1] < 20 and 12 < err[1] / err[2] < 20       # -> 2^4 = 16 (stiff order reduction at large dt)


def test_linear_propagator_is_exact():
    """Without the nonlinearity (tiny amplitude) a Fourier mode decays / grows as exp((k^2-k^4) t)."""
    N, L = 64, 22.0
    x = np.arange(N) * L / N
    for m in (1, 2, 3, 6):
        k = 2 * np.pi * m / L
        u0 = 1e-9 * np.cos(k * x)
        u1, _ = ke.step(u0, np.zeros(N), N, L, 0.125, 2)
        assert rel_l2(u1, u0 * np.exp((k ** 2 - k ** 4) * 0.25)) < 1e-7


def test_agrees_with_the_reference_scheme_up_to_its_spatial_error_which_converges():
    """FD-RK4 (the reference scheme, C oracle) on N = 64, 128, 256 points of the same L = 22 domain
    against the spectral solution on the same grid: 1.1e-3, 6.3e-5, 4.5e-6 -- the gap is the FD truncation error."""
    L, T = 22.0, 0.25
    xs = lambda N: np.arange(N) * L / N
    u0f = lambda x: 1.2 * np.cos(2 * np.pi * 2 * x / L + 0.3) + 0.8 * np.sin(2 * np.pi * 3 * x / L) + 0.3 * np.cos(2 * np.pi * x / L)
    gaps = []
    for N, dt in ((64, 1e-3), (128, 5e-5), (256, 3.125e-6)):
        steps = int(round(T / dt))
        cfg = ko.KSConfig(L=L, N=N, dt=dt, cfg_steps=steps)
        u0 = u0f(xs(N))
        u_fd, _ = ks_c.step(cfg, u0[None], np.zeros((1, N), np.float32))
        u_sp, _ = ke.step(u0, np.zeros(N), N, L, T / 50, 50)
        gaps.append(float(rel_l2(u_fd[0], u_sp)))
    assert gaps[0] < 5e-3 and gaps[1] < gaps[0] / 8 and gaps[2] < gaps[1] / 8, gaps


def test_packing_two_real_fields_into_one_complex_transform_is_exact():
    """The CUDA kernel evolves Z = U_a + i U_b without unpacking; every ETDRK4 operator has a real
    kernel, so this equals two independent real evolutions (checked here in NumPy)."""
    rng = np.random.default_rng(0)
    N, L, h = 64, 22.0, 0.025
    c = ke.etd_coefficients(N, L, h)
    ua, ub = rng.uniform(-1, 1, (2, N))
    pa, pb = rng.uniform(-0.5, 0.5, (2, N))
    z = np.fft.fft(ua + 1j * ub)
    ph = np.fft.fft(pa + 1j * pb)

    def nl(w):
        p = np.fft.ifft(w)
        return 1j * c.g * np.fft.fft(p.real ** 2 + 1j * p.imag ** 2) + ph

    Nv = nl(z); a = c.E2 * z + c.Q * Nv; Na = nl(a); b = c.E2 * z + c.Q * Na; Nb = nl(b)
    cc = c.E2 * a + c.Q * (2 * Nb - Nv); Nc = nl(cc)
    z1 = np.fft.ifft(c.E * z + Nv * c.f1 + 2 * (Na + Nb) * c.f2 + Nc * c.f3)
    va, _ = ke.etdrk4_step(np.fft.fft(ua), c, np.fft.fft(pa))
    vb, _ = ke.etdrk4_step(np.fft.fft(ub), c, np.fft.fft(pb))
    assert np.allclose(z1.real, np.fft.ifft(va).real, rtol=0, atol=1e-13)
    assert np.allclose(z1.imag, np.fft.ifft(vb).real, rtol=0, atol=1e-13)


def test_independent_high_accuracy_integration_of_the_same_ode():
    """A second, independent checker of the oracle: the Fourier-Galerkin ODE system the ETDRK4 scheme
    discretises in time -- d v_k / dt = (k^2 - k^4) v_k - (i k / 2) dealias_k FFT(u^2)_k + phi_hat_k --
    integrated by SciPy's adaptive 8th-order Runge-Kutta (DOP853, rtol 1e-12) in physical-space
    variables, jets on, over one control period.  ETDRK4 converges to that solution at 4th order
    (3e-8 at the benchmarked dt = 0.025; 2e-12 at dt / 16), on the default and on the large domain."""
    from scipy.integrate import solve_ivp

    for name, steps_coarse in (("attractor_default_random", 10), ("attractor_large_random", 10)):
        g = load_golden(name)
        N, L = int(g["N"]), float(g["L"])
        u0, phi = g["u0"], g["phi"][0].astype(np.float64)
        c = ke.etd_coefficients(N, L, 1.0)                # only k, lin, mask are used below (h-independent)
        phi_hat = np.fft.fft(phi)

        def rhs(_t, u):
            v = np.fft.fft(u)
            dv = c.lin * v + 1j * c.g * np.fft.fft(u * u) + phi_hat
            return np.real(np.fft.ifft(dv))

        sol = solve_ivp(rhs, (0.0, 0.25), u0, method="DOP853", rtol=1e-12, atol=1e-14)
        assert sol.success
        exact = sol.y[:, -1]
        e10 = float(rel_l2(ke.step(u0, phi, N, L, 0.25 / steps_coarse, steps_coarse)[0], exact))
        e40 = float(rel_l2(ke.step(u0, phi, N, L, 0.25 / 40, 40)[0], exact))
        e160 = float(rel_l2(ke.step(u0, phi, N, L, 0.25 / 160, 160)[0], exact))
        assert e10 < 2e-7 and e160 < 5e-11, (name, e10, e40, e160)
        assert 100 < e10 / e40 < 400, (name, e10, e40)      # 4^4 = 256
