"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the *spectral ETDRK4 mode* (SURVEY.md 8f-3).

The reference (stwerner97/model-based-pde-control) contains NO spectral solver: its scheme is
periodic finite differences + classic RK4 (pdegym/kuramoto/kuramoto.py:83-90,118-129).  The
north star nevertheless names an exponential-integrator / FFT kernel, so the library offers it
as a second solver (``solver="etdrk4"``).  Its algorithm is the published one:

  * Cox & Matthews, "Exponential time differencing for stiff systems", JCP 176 (2002): ETDRK4.
  * Kassam & Trefethen, "Fourth-order time-stepping for stiff PDEs", SISC 26 (2005): the
    contour-integral evaluation of the phi-functions (M points on a unit circle around L*h) and
    the KS example ``u_t = -u u_x - u_xx - u_xxxx`` this file restates for the env's equation
    ``u_t = -u_xxxx - u_xx - 1/2 (u^2)_x + phi`` (kuramoto.py:127).

**Parity unpinned by the reference** (there is nothing in it to pin against): the CUDA spectral
kernel is checked (a) against this restatement on identical inputs (same algorithm, fp64,
<= 1e-10 relative L2 per control period), and (b) against the *reference scheme* only through
convergence (ETDRK4 and the reference FD-RK4 approach each other as both discretisations are
refined; at the default grid they differ by ~3e-3 per control period) and
through the long-horizon statistics fixture (``tests/golden/stats_*.npz``).

Everything that is *not* the time stepper is shared with the reference path and follows it:
  jet forcing phi = a @ F in float32          pdegym/common/transforms.py:250-265
  reward = mean over sub-steps of -mean(u^2) of the PRE-step state   kuramoto.py:64-65,82-84,96
  (dissipation mode: -(mean(uxx^2) + mean(ux^2) + mean(u phi)) with spectral derivatives; as in
   the reference's code "ux" is the derivative of u^2, kuramoto.py:67-70,120-122)
  timestep / truncation / observation cast                            kuramoto.py:92-98
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CONTOUR_POINTS = 32      # Kassam & Trefethen use 16..64; 32 gives full double precision here


@dataclass
class ETDCoefficients:
    """Per-wavenumber tables, natural FFT order (``np.fft.fftfreq`` order), all real."""

    k: np.ndarray        # wavenumbers for odd derivatives (Nyquist entry zeroed)
    k_even: np.ndarray   # wavenumbers for even derivatives (Nyquist keeps |k|)
    lin: np.ndarray      # L(k) = k^2 - k^4 (Nyquist entry uses |k| = pi N / L)
    E: np.ndarray        # exp(h L)
    E2: np.ndarray       # exp(h L / 2)
    Q: np.ndarray        # h * phi1(h L / 2) / 2   (Kassam-Trefethen's Q)
    f1: np.ndarray
    f2: np.ndarray
    f3: np.ndarray
    g: np.ndarray        # nonlinear multiplier: N_hat = 1j * g * fft(u^2) + phi_hat, g = -k/2 * dealias
    mask: np.ndarray     # 2/3-rule dealiasing mask (1 = kept)
    h: float


def etd_coefficients(N: int, L: float, h: float, dealias: bool = True, M: int = CONTOUR_POINTS) -> ETDCoefficients:
    m = np.fft.fftfreq(N, 1.0 / N)                       # 0, 1, ..., N/2-1, -N/2, ..., -1
    kk = 2.0 * np.pi / L * m
    k_even = kk.copy()                                   # even derivatives: Nyquist keeps |k|
    k_odd = kk.copy()
    if N % 2 == 0:
        k_odd[N // 2] = 0.0                              # odd derivative of the Nyquist mode is zero
    lin = k_even ** 2 - k_even ** 4
    E = np.exp(h * lin)
    E2 = np.exp(h * lin / 2.0)
    r = np.exp(1j * np.pi * (np.arange(1, M + 1) - 0.5) / M)    # upper half of the unit circle
    LR = h * lin[:, None] + r[None, :]
    Q = h * np.real(np.mean((np.exp(LR / 2.0) - 1.0) / LR, axis=1))
    f1 = h * np.real(np.mean((-4.0 - LR + np.exp(LR) * (4.0 - 3.0 * LR + LR ** 2)) / LR ** 3, axis=1))
    f2 = h * np.real(np.mean((2.0 + LR + np.exp(LR) * (-2.0 + LR)) / LR ** 3, axis=1))
    f3 = h * np.real(np.mean((-4.0 - 3.0 * LR - LR ** 2 + np.exp(LR) * (4.0 - LR)) / LR ** 3, axis=1))
    if dealias:
        mask = (np.abs(m) <= N // 3).astype(np.float64)  # 2/3 rule: keep |m| <= N/3
    else:
        mask = np.ones(N)
    g = -0.5 * k_odd * mask
    return ETDCoefficients(k=k_odd, k_even=k_even, lin=lin, E=E, E2=E2, Q=Q, f1=f1, f2=f2, f3=f3, g=g, mask=mask, h=h)


def nonlinear(v: np.ndarray, c: ETDCoefficients, phi_hat: np.ndarray):
    """N_hat(v) = -(i k / 2) * dealias * fft(u^2) + phi_hat with u = real(ifft(v)).  Returns (N_hat, u)."""
    u = np.real(np.fft.ifft(v, axis=-1))
    return 1j * c.g * np.fft.fft(u * u, axis=-1) + phi_hat, u


def etdrk4_step(v: np.ndarray, c: ETDCoefficients, phi_hat: np.ndarray):
    """One ETDRK4 step (Cox-Matthews eq. 26-29 / Kassam-Trefethen kursiv.m).  Returns (v_new, u_pre)."""
    Nv, u_pre = nonlinear(v, c, phi_hat)
    a = c.E2 * v + c.Q * Nv
    Na, _ = nonlinear(a, c, phi_hat)
    b = c.E2 * v + c.Q * Na
    Nb, _ = nonlinear(b, c, phi_hat)
    cc = c.E2 * a + c.Q * (2.0 * Nb - Nv)
    Nc, _ = nonlinear(cc, c, phi_hat)
    v_new = c.E * v + Nv * c.f1 + 2.0 * (Na + Nb) * c.f2 + Nc * c.f3
    return v_new, u_pre


def substep_reward(u: np.ndarray, v: np.ndarray, phi: np.ndarray, c: ETDCoefficients, reward_mode: str) -> np.ndarray:
    """Per-sub-step reward of the pre-step state (kuramoto.py:64-70), batched over leading dims."""
    if reward_mode == "l2":
        return -np.mean(u * u, axis=-1)
    # kuramoto.py:67-70 with spectral derivatives.  Literally as the reference: rhs() differentiates
    # u**2 (kuramoto.py:120-122), so "ux" is the derivative of u^2 (not dealiased), uxx that of u.
    ux = np.real(np.fft.ifft(1j * c.k * np.fft.fft(u * u, axis=-1), axis=-1))
    uxx = np.real(np.fft.ifft(-(c.k_even ** 2) * v, axis=-1))
    return -(np.mean(uxx * uxx, axis=-1) + np.mean(ux * ux, axis=-1) + np.mean(u * phi, axis=-1))


def step(u0: np.ndarray, phi: np.ndarray, N: int, L: float, dt: float, cfg_steps: int, reward_mode: str = "l2",
         dealias: bool = True, coef: ETDCoefficients | None = None):
    """One control period of the spectral env: ``cfg_steps`` ETDRK4 steps of size ``dt``.

    ``u0 [..., N]`` float64, ``phi [..., N]`` (float32 jets, promoted).  Returns ``(u1, reward)``
    with ``reward = mean_s r(u_s)`` over the pre-step states (kuramoto.py:82-84,96).
    The state is carried in Fourier space through the whole period (as the CUDA kernel does).
    """
    c = coef if coef is not None else etd_coefficients(N, L, dt, dealias)
    u0 = np.asarray(u0, dtype=np.float64)
    phi64 = np.asarray(phi, dtype=np.float64)
    phi_hat = np.fft.fft(phi64, axis=-1)
    v = np.fft.fft(u0, axis=-1)
    reward = np.zeros(u0.shape[:-1])
    for _ in range(cfg_steps):
        v_new, u_pre = etdrk4_step(v, c, phi_hat)
        reward = reward + substep_reward(u_pre, v, phi64, c, reward_mode)
        v = v_new
    return np.real(np.fft.ifft(v, axis=-1)), reward / cfg_steps
