"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the plain-C oracle (``oracle/ks_oracle.c``).

Used by tests (fast long-horizon statistics), ``smoke()`` and ``bench.py``'s CPU legs.  The
product package never imports this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("KS_ORACLE_LIB") or os.path.join(_HERE, "_build", "libks_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if os.environ.get("KS_ORACLE_LIB"):      # caller supplies its own build (e.g. -march=native)
        return _LIB_PATH
    src = os.path.join(_HERE, "ks_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "all"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.kso_forcing.argtypes = [ctypes.c_int] * 3 + [f32p, f32p, f32p]
        L.kso_forcing.restype = None
        L.kso_step.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double] * 2 + [ctypes.c_int, f64p, f32p, f64p]
        L.kso_step.restype = ctypes.c_int
        L.kso_rollout.argtypes = ([ctypes.c_int] * 5 + [ctypes.c_double] * 2 + [ctypes.c_int, f64p]
                                  + [ctypes.c_void_p] * 4)
        L.kso_rollout.restype = ctypes.c_int
        L.kso_num_threads.restype = ctypes.c_int
        L.kso_set_threads.argtypes = [ctypes.c_int]
        L.kso_set_threads.restype = None
        _lib = L
    return _lib


_MODE = {"l2": 0, "dissipation": 1}


def forcing(actions: np.ndarray, F: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(actions, dtype=np.float32)
    F = np.ascontiguousarray(F, dtype=np.float32)
    J, N = F.shape
    a2 = a.reshape(-1, J)
    phi = np.empty((a2.shape[0], N), dtype=np.float32)
    lib().kso_forcing(a2.shape[0], J, N, a2, F, phi)
    return phi.reshape(a.shape[:-1] + (N,))


def step(cfg, u: np.ndarray, phi: np.ndarray):
    """One control period for ``u [B,N]``; returns ``(u_next, reward [B])`` (inputs untouched)."""
    u = np.array(u, dtype=np.float64, order="C").reshape(-1, cfg.N)
    phi = np.ascontiguousarray(np.broadcast_to(np.asarray(phi, dtype=np.float32), u.shape))
    rew = np.empty(u.shape[0], dtype=np.float64)
    rc = lib().kso_step(u.shape[0], cfg.N, cfg.cfg_steps, cfg.dt, cfg.dx, _MODE[cfg.reward_mode], u, phi, rew)
    if rc:
        raise RuntimeError(f"kso_step failed: {rc}")
    return u, rew


def rollout(cfg, u: np.ndarray, actions, F: np.ndarray, K: int | None = None, want_obs=True):
    """``K`` periods; ``actions [K,B,J]`` (``None`` = no-op burn-in, then ``K`` is required).
    Returns ``(u_final [B,N], obs [K,B,N] f32 or None, reward [K,B])``."""
    u = np.array(u, dtype=np.float64, order="C").reshape(-1, cfg.N)
    B = u.shape[0]
    F = np.ascontiguousarray(F, dtype=np.float32)
    if actions is not None:
        actions = np.ascontiguousarray(actions, dtype=np.float32).reshape(-1, B, cfg.J)
        K = actions.shape[0]
    obs = np.empty((K, B, cfg.N), dtype=np.float32) if want_obs else None
    rew = np.empty((K, B), dtype=np.float64)
    rc = lib().kso_rollout(K, B, cfg.N, cfg.J, cfg.cfg_steps, cfg.dt, cfg.dx, _MODE[cfg.reward_mode], u,
                           actions.ctypes.data if actions is not None else None, F.ctypes.data,
                           obs.ctypes.data if obs is not None else None, rew.ctypes.data)
    if rc:
        raise RuntimeError(f"kso_rollout failed: {rc}")
    return u, obs, rew


def num_threads() -> int:
    return lib().kso_num_threads()


def set_threads(n: int) -> None:
    """``0`` = one thread per online core (default)."""
    lib().kso_set_threads(int(n))
