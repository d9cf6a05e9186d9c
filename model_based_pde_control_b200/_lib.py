"""ctypes binding of ``libks_b200.so`` (C ABI declared in ``include/ks_b200.h``).

There is no CPU fallback and no lazy degradation: if the shared library has not been built
(``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C
model_based_pde_control_b200/csrc``) importing the binding raises.
"""
from __future__ import annotations

import ctypes
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# KS_LIB_PATH: development override (tools/sweep.py times experimental builds of the same ABI side by side)
LIB_PATH = os.environ.get("KS_LIB_PATH") or os.path.join(_PKG_DIR, "libks_b200.so")

KS_ABI_VERSION = 3
KS_F64, KS_F32 = 0, 1
KS_REWARD_L2, KS_REWARD_DISSIPATION = 0, 1
KS_HOST, KS_DEVICE = 0, 1
KS_SOLVER_FD_RK4, KS_SOLVER_ETDRK4 = 0, 1
KS_MAX_WORLD, KS_IPC_HANDLE_BYTES = 16, 64
KS_ERR_ARG, KS_ERR_UNSUPPORTED, KS_ERR_NO_DEVICE, KS_ERR_STATE = -1, -2, -3, -4

PRECISIONS = {"f64": KS_F64, "fp64": KS_F64, "float64": KS_F64, "f32": KS_F32, "fp32": KS_F32, "float32": KS_F32}
REWARD_MODES = {"l2": KS_REWARD_L2, "dissipation": KS_REWARD_DISSIPATION}
SOLVERS = {"fd_rk4": KS_SOLVER_FD_RK4, "rk4": KS_SOLVER_FD_RK4, "reference": KS_SOLVER_FD_RK4,
           "etdrk4": KS_SOLVER_ETDRK4, "spectral": KS_SOLVER_ETDRK4}


class KsConfig(ctypes.Structure):
    """``struct ks_config`` (include/ks_b200.h)."""

    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("num_envs", ctypes.c_int32),
        ("N", ctypes.c_int32),
        ("J", ctypes.c_int32),
        ("cfg_steps", ctypes.c_int32),
        ("max_episode_steps", ctypes.c_int32),
        ("burnin_periods", ctypes.c_int32),
        ("precision", ctypes.c_int32),
        ("reward_mode", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("points_per_lane", ctypes.c_int32),
        ("obs_stride", ctypes.c_int32),
        ("solver", ctypes.c_int32),
        ("dealias", ctypes.c_int32),
        ("env_index_base", ctypes.c_int32),
        ("reserved0", ctypes.c_int32),
        ("L", ctypes.c_double),
        ("dt", ctypes.c_double),
        ("forcing", ctypes.c_void_p),
    ]


class KsCollectArgs(ctypes.Structure):
    """``struct ks_collect_args`` (include/ks_b200.h)."""

    _fields_ = [(n, ctypes.c_void_p) for n in (
        "actions", "obs", "reward", "truncated", "step", "obs_store", "act_store", "vminmax", "agent_obs",
        "rec_obs", "rec_actions", "rec_nxtobs", "rec_reward", "rec_truncated", "rec_step")] + [
        ("lower", ctypes.c_float), ("scale_width", ctypes.c_float), ("frozen", ctypes.c_int32),
        ("agent_stride", ctypes.c_int32), ("slot_index", ctypes.c_void_p)]


# every symbol include/ks_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64
_SIZE5 = ctypes.c_size_t * 5
EXPORTS = {
    "ks_abi_version": (ctypes.c_int, []),
    "ks_last_error": (ctypes.c_char_p, [_vp]),
    "ks_create": (ctypes.c_int, [ctypes.POINTER(KsConfig), ctypes.POINTER(_vp)]),
    "ks_destroy": (ctypes.c_int, [_vp]),
    "ks_set_state": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp]),
    "ks_get_state": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _vp]),
    "ks_reset": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int, _u64, _i32, _vp]),
    "ks_step": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ks_rollout": (ctypes.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ks_step_host": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "ks_out_layout": (ctypes.c_int, [_vp, ctypes.POINTER(_SIZE5), ctypes.POINTER(ctypes.c_size_t)]),
    "ks_status": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(_i32), _vp]),
    "ks_eval": (ctypes.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ks_get_config": (ctypes.c_int, [_vp, ctypes.POINTER(KsConfig)]),
    "ks_launch_info": (ctypes.c_int, [_vp] + [ctypes.POINTER(_i32)] * 5),
    "ks_launch_count": (ctypes.c_uint64, [_vp]),
    "ks_gather_init": (ctypes.c_int, [_vp, _i32, _i32, _vp, ctypes.POINTER(ctypes.c_size_t)]),
    "ks_gather_connect": (ctypes.c_int, [_vp, _vp]),
    "ks_step_gather": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(_vp), _vp]),
    "ks_gather_status": (ctypes.c_int, [_vp, ctypes.POINTER(_i32), _vp]),
    "ks_gather_clear": (ctypes.c_int, [_vp, _vp]),
    "ks_gather_barrier": (ctypes.c_int, [_vp, _vp]),
    "ks_gather_layout": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "ks_gather_attach": (ctypes.c_int, [_vp, _i32, _i32, ctypes.POINTER(_vp), _vp, ctypes.c_size_t]),
    "ks_collect": (ctypes.c_int, [_vp, ctypes.POINTER(KsCollectArgs), _vp]),
    "ks_bench_fp64_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "ks_bench_fp32_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
}

_lib = None


class KsError(RuntimeError):
    """A non-zero status from the C ABI (``code`` < 0: argument/state error, > 0: cudaError_t)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libks_b200 error {code}: {message}")
        self.code = code


def load() -> ctypes.CDLL:
    """Load the in-tree shared library; raise loudly when it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not built. Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C model_based_pde_control_b200/csrc`. There is no CPU fallback for the KS kernels.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.ks_abi_version() != KS_ABI_VERSION:
        raise ImportError(f"libks_b200 ABI {lib.ks_abi_version()} != binding ABI {KS_ABI_VERSION}")
    _lib = lib
    return lib


def check(handle, code: int) -> None:
    if code != 0:
        msg = load().ks_last_error(handle)
        raise KsError(code, msg.decode() if msg else "?")
