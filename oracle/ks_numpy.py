"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Kuramoto-Sivashinsky control environment.

An independent NumPy restatement of the reference algorithm (not a copy: periodic stencils
are written with ``np.roll`` from the maths, the jet forcing as an explicit fp32 FMA chain).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it; the product package never does.

Parity pinning: the reference has no tests or golden vectors of its own (its ``tests/`` is an
empty ``__init__.py``), so this oracle is pinned against outputs of the reference itself,
executed in the build container through ``oracle/ref_loader.py`` and committed as fixtures in
``tests/golden/`` by ``tests/golden/make_golden.py`` (``tests/test_oracle.py`` checks them on
every run, and checks the live reference too where ``/root/reference`` exists).

Reference lines followed (all paths relative to the reference root):

=====================  ======================================  =================================
here                   reference                               what
=====================  ======================================  =================================
``KSConfig``           ``pdegym/kuramoto/kuramoto.py:18,29-57``  constants, grid, episode length
``forcing_matrix``     ``pdegym/common/transforms.py:250-260``   fp32 Gaussian jets ``F[J,N]``
``forcing``            ``pdegym/common/transforms.py:262-265``   ``phi = a @ F`` (fp32)
``rhs``                ``pdegym/kuramoto/kuramoto.py:24-27,118-129``  periodic FD right-hand side
``reward_l2`` / ``..`` ``pdegym/kuramoto/kuramoto.py:64-73``     per-sub-step reward
``step``               ``pdegym/kuramoto/kuramoto.py:78-98``     250 x classic RK4 + reward mean
``initial_condition``  ``pdegym/kuramoto/kuramoto.py:101,106``   ``U(-0.4,0.4)^N`` global RNG
``reset``              ``pdegym/kuramoto/kuramoto.py:100-116``   800 no-op periods of burn-in
``KSOracleEnv``        ``kuramoto.py`` + ``pdegym/kuramoto/__init__.py:8-12``  single-env API
=====================  ======================================  =================================
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

# Finite-difference weights in *stencil* order (offset -> weight), i.e. what the reference's
# flipped ``convolve1d`` weight lists (kuramoto.py:24-27) evaluate to:
#   fwd_i = sum_k UPWIND[k] * q[i+k] / dx      (k = 0..4)        used where u_i <  0
#   bwd_i = sum_k -UPWIND[k] * q[i-k] / dx     (k = 0..4)        used where u_i >= 0
UPWIND = (-25.0 / 12.0, 4.0, -3.0, 4.0 / 3.0, -1.0 / 4.0)
# central, offsets -3..3 and -4..4
D2 = (1.0 / 90.0, -3.0 / 20.0, 3.0 / 2.0, -49.0 / 18.0, 3.0 / 2.0, -3.0 / 20.0, 1.0 / 90.0)
D4 = (7.0 / 240.0, -2.0 / 5.0, 169.0 / 60.0, -122.0 / 15.0, 91.0 / 8.0, -122.0 / 15.0,
      169.0 / 60.0, -2.0 / 5.0, 7.0 / 240.0)

BURNIN_TIME = 200.0     # kuramoto.py:103
IC_AMPLITUDE = 0.4      # kuramoto.py:106


@dataclass
class KSConfig:
    """Constants of ``KuramotoSivashinskyEnv.__init__`` (kuramoto.py:29-57)."""

    L: float = 22.0
    N: int = 64
    cfg_steps: int = 250
    Tmax: float = 100.0
    dt: float = 0.001
    sigma: float = 0.4
    Xi: Sequence[float] = (0.0, 0.25, 0.5, 0.75)     # class attribute in the reference (:18)
    reward_mode: str = "l2"       # "l2" = what the reference executes; "dissipation" = intended
    dx: float = field(init=False)
    max_episode_steps: int = field(init=False)
    burnin_periods: int = field(init=False)

    def __post_init__(self):
        self.dx = self.L / self.N                                             # :55
        self.max_episode_steps = math.ceil(self.Tmax / (self.dt * self.cfg_steps))   # :57
        self.burnin_periods = int(BURNIN_TIME / self.dt / self.cfg_steps)     # :103
        if self.reward_mode not in ("l2", "dissipation"):
            raise ValueError(f"unknown reward_mode {self.reward_mode!r}")

    @property
    def J(self) -> int:
        return len(self.Xi)

    @property
    def x(self) -> np.ndarray:
        return np.linspace(0.0, self.L - self.L / self.N, self.N, dtype=np.float32)   # :56


# ------------------------------------------------------------------------------------------------
# Gaussian-jet forcing (float32 end to end)
# ------------------------------------------------------------------------------------------------

def forcing_matrix(cfg: KSConfig) -> np.ndarray:
    """``F[j,i] = exp(-(x_i - L*Xi_j)^2 / (2 sigma^2)) / sqrt(2 pi sigma)`` in float32.

    transforms.py:253-260.  Not periodic, and the normaliser really is ``sqrt(2*pi*sigma)``.
    torch's CPU float32 ``exp`` returns the correctly rounded value on the boxes probed, which
    NumPy's SIMD float32 ``exp`` does not (1-3 ulp off), so the exponential is evaluated in
    float64 and rounded once; every other operation is a plain float32 op as in torch.
    """
    f32 = np.float32
    x = cfg.x
    xi = (f32(cfg.L) * np.asarray(cfg.Xi, dtype=f32)).reshape(-1, 1)
    d = x - xi
    arg = -(d * d) / f32(2.0 * cfg.sigma ** 2)
    e = np.exp(arg.astype(np.float64)).astype(f32)
    return (e / f32(np.sqrt(2.0 * np.pi * cfg.sigma))).astype(f32)


def _fmaf(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """Correctly rounded float32 ``a*b + c`` for arrays (NumPy has no fma).

    ``a*b`` is exact in float64 (24+24 significant bits).  The float64 sum ``s = a*b + c`` may
    round; rounding ``s`` again to float32 differs from the single correct rounding only when
    ``s`` sits exactly on a float32 tie *and* the float64 sum was inexact.  Those entries are
    resolved with the exact TwoSum error term.
    """
    a64, b64, c64 = (np.asarray(v, dtype=np.float32).astype(np.float64) for v in (a, b, c))
    p = a64 * b64                               # exact
    s = p + c64
    bb = s - p                                  # TwoSum (Knuth): s + err == p + c exactly
    err = (p - (s - bb)) + (c64 - bb)
    out = s.astype(np.float32)
    r64 = out.astype(np.float64)
    inexact32 = r64 != s
    if np.any(inexact32 & (err != 0.0)):
        toward = np.where(s > r64, np.float32(np.inf), np.float32(-np.inf)).astype(np.float32)
        nb = np.nextafter(out, toward)
        tie = inexact32 & (err != 0.0) & (((r64 + nb.astype(np.float64)) * 0.5) == s)
        if np.any(tie):
            # exact value = s + err: above the midpoint -> larger candidate, below -> smaller
            hi = np.maximum(out, nb)
            lo = np.minimum(out, nb)
            out = np.where(tie, np.where(err > 0.0, hi, lo), out)
    return out


def forcing(action: np.ndarray, F: np.ndarray) -> np.ndarray:
    """``phi = action @ F`` as torch's CPU sgemm evaluates it for ``[B,J] @ [J,N]``, J <= 8:

    a sequential-k float32 FMA chain ``acc = fmaf(a[k], F[k,i], acc)`` from ``acc = 0``
    (pinned by the ``phi`` arrays in ``tests/golden``).
    ``action``: ``[..., J]`` float32 -> ``[..., N]`` float32.
    """
    a = np.asarray(action, dtype=np.float32)
    F = np.asarray(F, dtype=np.float32)
    J, N = F.shape
    if a.shape[-1] != J:
        raise ValueError(f"action has {a.shape[-1]} jets, forcing matrix has {J}")
    acc = np.zeros(a.shape[:-1] + (N,), dtype=np.float32)
    for k in range(J):
        acc = _fmaf(a[..., k:k + 1], F[k], acc)
    return acc


# ------------------------------------------------------------------------------------------------
# Finite-difference right-hand side
# ------------------------------------------------------------------------------------------------

def derivatives(u: np.ndarray, dx: float):
    """``(ux, uxx, uxxxx)`` exactly as ``rhs`` returns them (kuramoto.py:120-125, 129).

    ``ux`` is the upwind derivative of ``u**2`` (not of ``u``); periodic in the last axis.
    """
    q = u * u
    fwd = UPWIND[0] * q
    bwd = -UPWIND[0] * q
    for k in range(1, 5):
        fwd = fwd + UPWIND[k] * np.roll(q, -k, axis=-1)
        bwd = bwd - UPWIND[k] * np.roll(q, k, axis=-1)
    ux = np.where(u < 0, fwd, bwd) / dx
    uxx = sum(D2[k + 3] * np.roll(u, -k, axis=-1) for k in range(-3, 4)) / dx ** 2
    uxxxx = sum(D4[k + 4] * np.roll(u, -k, axis=-1) for k in range(-4, 5)) / dx ** 4
    return ux, uxx, uxxxx


def rhs(u: np.ndarray, phi: np.ndarray, dx: float) -> np.ndarray:
    """``-uxxxx - uxx - ux/2 + phi`` (kuramoto.py:127); ``phi`` float32 promoted to float64."""
    ux, uxx, uxxxx = derivatives(u, dx)
    return -uxxxx - uxx - 0.5 * ux + phi


def rhs_scipy(u: np.ndarray, phi: np.ndarray, dx: float) -> np.ndarray:
    """Same RHS through ``scipy.ndimage.convolve1d(mode="wrap")`` -- the third-party routine
    the reference calls (kuramoto.py:120-125) -- used to cross-check the roll-based stencils
    and as the cost-faithful path of the CPU baseline."""
    from scipy.ndimage import convolve1d

    # convolve1d flips its weights: pass offsets +4..-4 order, centre at index 4
    w_fwd = [UPWIND[4], UPWIND[3], UPWIND[2], UPWIND[1], UPWIND[0], 0.0, 0.0, 0.0, 0.0]
    w_bwd = [0.0, 0.0, 0.0, 0.0, -UPWIND[0], -UPWIND[1], -UPWIND[2], -UPWIND[3], -UPWIND[4]]
    q = u ** 2
    fwd = convolve1d(q, weights=w_fwd, mode="wrap", axis=-1) / dx
    bwd = convolve1d(q, weights=w_bwd, mode="wrap", axis=-1) / dx
    ux = (u < 0) * fwd + (u >= 0) * bwd
    uxx = convolve1d(u, weights=list(D2), mode="wrap", axis=-1) / dx ** 2
    uxxxx = convolve1d(u, weights=list(D4), mode="wrap", axis=-1) / dx ** 4
    return -uxxxx - uxx - 0.5 * ux + phi


# ------------------------------------------------------------------------------------------------
# Reward, control period, reset
# ------------------------------------------------------------------------------------------------

def reward_l2(u: np.ndarray, N: int) -> np.ndarray:
    """Executed default (kuramoto.py:64-65,72): ``-(1/N) * ||u||_2**2`` per env."""
    return -(1.0 / N) * np.sqrt(np.sum(u * u, axis=-1)) ** 2


def reward_dissipation(u: np.ndarray, phi: np.ndarray, dx: float) -> np.ndarray:
    """Intended reward (kuramoto.py:67-70; unreachable without a TypeError in the reference):
    ``-(mean(uxx^2) + mean(ux^2) + mean(u*phi))`` with ``ux`` = upwind d/dx of ``u**2``."""
    ux, uxx, _ = derivatives(u, dx)
    return -((uxx * uxx).mean(axis=-1) + (ux * ux).mean(axis=-1) + (u * phi).mean(axis=-1))


def step(cfg: KSConfig, u: np.ndarray, phi: np.ndarray, rhs_fn=rhs):
    """One control period for a batch ``u [..., N]`` float64 with forcing ``phi [..., N]`` f32.

    kuramoto.py:82-96: per sub-step the reward of the *pre-step* state is accumulated, then
    one classic RK4 step; the period reward is the plain mean over sub-steps.
    Returns ``(u_next float64, reward float64 [...])``.
    """
    u = np.asarray(u, dtype=np.float64)
    dt, dx = cfg.dt, cfg.dx
    reward = np.zeros(u.shape[:-1], dtype=np.float64)
    for _ in range(cfg.cfg_steps):
        if cfg.reward_mode == "l2":
            reward = reward + reward_l2(u, cfg.N)
        else:
            reward = reward + reward_dissipation(u, phi, dx)
        k1 = rhs_fn(u, phi, dx)
        k2 = rhs_fn(u + dt * k1 / 2.0, phi, dx)
        k3 = rhs_fn(u + dt * k2 / 2.0, phi, dx)
        k4 = rhs_fn(u + dt * k3, phi, dx)
        u = u + dt * (k1 + 2.0 * k2 + 2.0 * k3 + k4) / 6.0
    return u, reward / cfg.cfg_steps


def initial_condition(cfg: KSConfig, seed) -> np.ndarray:
    """``np.random.seed(seed); np.random.uniform(-0.4, 0.4, N)`` (kuramoto.py:101,106) using a
    private legacy generator so the caller's global RNG is left alone (same MT19937 stream)."""
    rs = np.random.RandomState(seed)
    return rs.uniform(-IC_AMPLITUDE, IC_AMPLITUDE, size=cfg.N)


def reset(cfg: KSConfig, u0: np.ndarray, periods: int | None = None, rhs_fn=rhs) -> np.ndarray:
    """Burn-in: ``periods`` (default 800) control periods with zero action (kuramoto.py:108-109).
    Zero action gives ``phi = +0.0f`` exactly under the FMA chain."""
    periods = cfg.burnin_periods if periods is None else periods
    u = np.asarray(u0, dtype=np.float64)
    phi = np.zeros(u.shape, dtype=np.float32)
    for _ in range(periods):
        u, _ = step(cfg, u, phi, rhs_fn=rhs_fn)
    return u


class KSOracleEnv:
    """Single-env API of the reference (``reset``/``step`` incl. ``timestep`` and truncation),
    vectorised over a leading batch axis for convenience: ``u [B,N]``, ``timestep [B]``."""

    def __init__(self, cfg: KSConfig | None = None, num_envs: int = 1, F: np.ndarray | None = None,
                 rhs_fn=rhs):
        self.cfg = cfg or KSConfig()
        self.num_envs = num_envs
        self.F = forcing_matrix(self.cfg) if F is None else np.asarray(F, dtype=np.float32)
        self.rhs_fn = rhs_fn
        self.u = np.zeros((num_envs, self.cfg.N), dtype=np.float64)
        self.timestep = np.zeros(num_envs, dtype=np.int64)

    def set_state(self, u, timestep=0):
        self.u = np.array(u, dtype=np.float64).reshape(self.num_envs, self.cfg.N)
        self.timestep = np.broadcast_to(np.asarray(timestep, dtype=np.int64), (self.num_envs,)).copy()

    def step(self, actions, phi=None):
        """-> ``(obs [B,1,N] f64, reward [B] f64, terminated [B], truncated [B], {"step": [B]})``."""
        a = np.asarray(actions, dtype=np.float32).reshape(self.num_envs, self.cfg.J)
        phi = forcing(a, self.F) if phi is None else np.asarray(phi, dtype=np.float32)
        self.u, reward = step(self.cfg, self.u, phi, rhs_fn=self.rhs_fn)
        self.timestep = self.timestep + 1
        truncated = self.timestep >= self.cfg.max_episode_steps
        obs = self.u.reshape(self.num_envs, 1, self.cfg.N).copy()
        return obs, reward, np.zeros(self.num_envs, dtype=bool), truncated, {"step": self.timestep.copy()}

    def reset(self, seed=None, u0=None, periods=None):
        """Seeds follow gym 0.25.2's vector env: env ``i`` gets ``seed + i`` (``None`` stays None)."""
        if u0 is None:
            seeds = [None] * self.num_envs if seed is None else [seed + i for i in range(self.num_envs)]
            u0 = np.stack([initial_condition(self.cfg, s) for s in seeds])
        self.u = reset(self.cfg, np.asarray(u0, dtype=np.float64).reshape(self.num_envs, self.cfg.N),
                       periods=periods, rhs_fn=self.rhs_fn)
        self.timestep = np.zeros(self.num_envs, dtype=np.int64)
        return self.u.reshape(self.num_envs, 1, self.cfg.N).copy()


# ------------------------------------------------------------------------------------------------
# statistics used by the long-horizon gate (BASELINE.json north star: spectrum + dissipation)
# ------------------------------------------------------------------------------------------------

def energy_spectrum(u: np.ndarray) -> np.ndarray:
    """``|rfft(u)|^2 / N^2`` per wavenumber, last axis = space."""
    n = u.shape[-1]
    return np.abs(np.fft.rfft(u, axis=-1)) ** 2 / n ** 2


def dissipation_rate(u: np.ndarray, phi: np.ndarray, dx: float) -> np.ndarray:
    """``mean(uxx^2) + mean(ux^2) + mean(u*phi)`` (the negated 'dissipation' reward)."""
    return -reward_dissipation(u, phi, dx)
