"""The C-ABI library loads and exports every symbol include/ks_b200.h declares; argument errors
are reported without touching a device.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from model_based_pde_control_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ks_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ks_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 17
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in ks_b200.h but not exported by libks_b200.so"
        assert name in _lib.EXPORTS, f"{name} has no ctypes prototype in _lib.EXPORTS"
    assert set(_lib.EXPORTS) == set(syms)
    assert lib.ks_abi_version() == _lib.KS_ABI_VERSION == 3


def test_config_struct_layout_matches_header():
    """Field order/types of ks_config in the header == the ctypes mirror."""
    text = open(os.path.join(ROOT, "include", "ks_b200.h")).read()
    body = re.search(r"typedef struct ks_config \{(.*?)\} ks_config;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(int32_t|double|const float \*)\s*(\w+);", body)
    ctype = {"int32_t": ctypes.c_int32, "double": ctypes.c_double, "const float *": ctypes.c_void_p}
    assert [(n, ctype[t]) for t, n in fields] == list(_lib.KsConfig._fields_)
    assert ctypes.sizeof(_lib.KsConfig) == 16 * 4 + 2 * 8 + 8


def test_collect_args_struct_layout_matches_header():
    """ks_collect_args: every pointer / scalar of the header, in order, in the ctypes mirror."""
    text = open(os.path.join(ROOT, "include", "ks_b200.h")).read()
    body = re.search(r"typedef struct ks_collect_args \{(.*?)\} ks_collect_args;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        kind = "ptr" if "*" in decl else ("f32" if decl.startswith("float") else "i32")
        for part in decl.split(","):
            names.append((re.findall(r"(\w+)\s*$", part.strip())[0], kind))
    ctype = {"ptr": ctypes.c_void_p, "f32": ctypes.c_float, "i32": ctypes.c_int32}
    assert [(n, ctype[k]) for n, k in names] == list(_lib.KsCollectArgs._fields_)


def good_config(**kw):
    F = np.zeros((4, 64), np.float32)
    c = dict(abi_version=3, num_envs=8, N=64, J=4, cfg_steps=250, max_episode_steps=400, burnin_periods=800,
             precision=0, reward_mode=0, device=0, points_per_lane=0, obs_stride=1, L=22.0, dt=1e-3,
             forcing=F.ctypes.data)
    c.update(kw)
    return _lib.KsConfig(**c), F


@pytest.mark.parametrize("bad,code", [
    (dict(abi_version=99), _lib.KS_ERR_ARG), (dict(num_envs=0), _lib.KS_ERR_ARG), (dict(N=4), _lib.KS_ERR_ARG),
    (dict(J=0), _lib.KS_ERR_ARG), (dict(J=33), _lib.KS_ERR_ARG), (dict(cfg_steps=0), _lib.KS_ERR_ARG),
    (dict(L=-1.0), _lib.KS_ERR_ARG), (dict(dt=0.0), _lib.KS_ERR_ARG), (dict(precision=7), _lib.KS_ERR_ARG),
    (dict(reward_mode=5), _lib.KS_ERR_ARG), (dict(forcing=None), _lib.KS_ERR_ARG), (dict(obs_stride=-2), _lib.KS_ERR_ARG),
    (dict(N=1031), _lib.KS_ERR_UNSUPPORTED),          # prime > 16 points: no lanes*P layout
    (dict(N=64, points_per_lane=5), _lib.KS_ERR_UNSUPPORTED),
    (dict(N=1024), _lib.KS_ERR_UNSUPPORTED),          # would need more than one warp per env
    (dict(solver=3), _lib.KS_ERR_ARG),
    (dict(solver=1, N=96), _lib.KS_ERR_UNSUPPORTED),       # spectral solver: N = 64, 128, 256 only
])
def test_create_rejects_bad_config_without_a_device(bad, code):
    lib = _lib.load()
    cfg, _keep = good_config(**bad)
    h = ctypes.c_void_p()
    rc = lib.ks_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == code and not h.value
    assert lib.ks_last_error(None)          # a message is available for the failed creation


def test_create_fails_loudly_without_gpu_and_null_handles_are_safe():
    import torch

    lib = _lib.load()
    if not torch.cuda.is_available():
        cfg, _keep = good_config()
        h = ctypes.c_void_p()
        rc = lib.ks_create(ctypes.byref(cfg), ctypes.byref(h))
        assert rc == _lib.KS_ERR_NO_DEVICE and not h.value
        assert b"no CPU path" in lib.ks_last_error(None)
    assert lib.ks_destroy(None) == 0
    assert lib.ks_step(None, None, None, None, None, None, None, None, None) == _lib.KS_ERR_ARG
    assert lib.ks_get_state(None, None, None, 0, None) == _lib.KS_ERR_ARG
    assert lib.ks_launch_count(None) == 0
    with pytest.raises(_lib.KsError):
        _lib.check(None, _lib.KS_ERR_ARG)


def test_library_has_sm100a_code_and_no_torch_dependency():
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "libc10" not in ldd
