"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference KS environment.

Only ``tests/``, ``tests/golden/make_golden.py`` and ``oracle/`` self-checks may import this.
Nothing on the product path (``model_based_pde_control_b200``) may import anything under
``oracle/``.

The reference (``/root/reference``, read-only, present in the build container only; installed
byte-identically into ``baseline/_ref`` by ``oracle/install_ref.py`` so that it travels to the GPU box) cannot be
imported as-is:

* ``pdegym/__init__.py:2`` imports a ``pdegym.burgers`` package that is not in the tree
  -> a bare ``pdegym`` package object with the right ``__path__`` is registered instead;
* ``pdegym/kuramoto/kuramoto.py:4`` imports ``gym`` (0.25.2, not installed, no network)
  -> a stub exposing exactly the attributes the reference touches;
* ``pdegym/common/transforms.py:7`` imports ``pdecontrol.mbrl.types`` which imports
  ``pytorch_lightning`` (``pdecontrol/mbrl/types.py:6``) -> stub module.

With those three stubs ``KuramotoSivashinskyEnv`` constructs and ``step``/``reset``/``rhs``/
``forcing`` execute the reference's own code, byte for byte.  This file contains no reference
code; it only arranges ``sys.modules`` so that the reference's files can be executed where they
lie.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Search order: an explicit override, the install that travels to the GPU box
# (``baseline/_ref``, made by ``oracle/install_ref.py``: byte-identical .py files), the build
# container's read-only tree.
REFERENCE_ROOTS = (
    os.environ.get("KS_REFERENCE_ROOT", ""),
    os.path.join(_REPO, "baseline", "_ref"),
    "/root/reference",
)


def reference_root() -> str | None:
    for root in REFERENCE_ROOTS:
        if root and os.path.isfile(os.path.join(root, "pdegym", "kuramoto", "kuramoto.py")):
            return root
    return None


def reference_available() -> bool:
    return reference_root() is not None


class _Box:
    """Just enough of ``gym.spaces.Box`` for ``kuramoto.py:75-76``."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi, size=self.shape).astype(self.dtype)


class _Env:
    """Just enough of ``gym.Env``."""

    metadata: dict = {}

    @property
    def unwrapped(self):
        return self


class _Wrapper:
    def __init__(self, env, *args, **kwargs):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class _VectorEnvWrapper:
    """The behaviour of ``gym.vector.VectorEnvWrapper`` (0.25.2) the reference's wrappers rely on:
    holds ``env``, forwards unknown attributes, ``step`` = ``step_async`` + ``step_wait``."""

    def __init__(self, env):
        self.env = env

    def reset_async(self, **kwargs):
        return self.env.reset_async(**kwargs)

    def reset_wait(self, **kwargs):
        return self.env.reset_wait(**kwargs)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step_async(self, actions):
        return self.env.step_async(actions)

    def step_wait(self):
        return self.env.step_wait()

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self, **kwargs):
        return self.env.close(**kwargs)

    def __getattr__(self, name):
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped


def _install_stubs(root: str) -> None:
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        gym.Env = _Env
        gym.Wrapper = _Wrapper
        spaces = types.ModuleType("gym.spaces")
        spaces.Box = _Box
        gym.spaces = spaces
        envs = types.ModuleType("gym.envs")
        envs.register = lambda *a, **k: None
        gym.envs = envs
        wrappers = types.ModuleType("gym.wrappers")
        wrappers.TimeLimit = _Wrapper
        wrappers.RescaleAction = _Wrapper
        gym.wrappers = wrappers
        core = types.ModuleType("gym.core")
        core.ObservationWrapper = _Wrapper
        core.ActionWrapper = _Wrapper
        gym.core = core
        gym.ObservationWrapper = _Wrapper
        gym.ActionWrapper = _Wrapper
        vector = types.ModuleType("gym.vector")
        vector.VectorEnv = type("VectorEnv", (), {})
        vector.VectorEnvWrapper = _VectorEnvWrapper
        gym.vector = vector
        sys.modules.update({
            "gym": gym, "gym.spaces": spaces, "gym.envs": envs, "gym.wrappers": wrappers,
            "gym.core": core, "gym.vector": vector,
        })
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.Trainer = type("Trainer", (), {})
        pl.Callback = type("Callback", (), {})
        pl.LightningModule = type("LightningModule", (), {})
        pl.LightningDataModule = type("LightningDataModule", (), {})
        sys.modules["pytorch_lightning"] = pl
    # bare package objects: the reference's own ``__init__`` files are NOT executed
    for pkg in ("pdegym", "pdegym.common", "pdegym.kuramoto", "pdecontrol", "pdecontrol.mbrl"):
        if pkg not in sys.modules:
            mod = types.ModuleType(pkg)
            mod.__path__ = [os.path.join(root, *pkg.split("."))]
            sys.modules[pkg] = mod


_CACHE: dict = {}


def load_reference():
    """Return ``(kuramoto_module, transforms_module)`` executed from the reference tree."""
    if "mods" in _CACHE:
        return _CACHE["mods"]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not available (expected /root/reference)")
    _install_stubs(root)
    import importlib.util

    def _load(name, relpath):
        if name in sys.modules:
            return sys.modules[name]
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    _load("pdecontrol.mbrl.types", "pdecontrol/mbrl/types.py")
    transforms = _load("pdegym.common.transforms", "pdegym/common/transforms.py")
    kuramoto = _load("pdegym.kuramoto.kuramoto", "pdegym/kuramoto/kuramoto.py")
    _CACHE["mods"] = (kuramoto, transforms)
    return _CACHE["mods"]


def load_reference_wrappers():
    """``(vec_wrappers module, transforms module)`` of the reference, executed where they lie.
    ``vec_wrappers.py`` uses ``np.bool8``, which NumPy 2 removed; the alias is restored in the
    NumPy namespace for the import (an environment shim, the reference file is untouched)."""
    _, transforms = load_reference()
    if "pdegym.common.vec_wrappers" not in sys.modules:
        if not hasattr(np, "bool8"):
            np.bool8 = np.bool_
        import importlib.util

        path = os.path.join(reference_root(), "pdegym", "common", "vec_wrappers.py")
        spec = importlib.util.spec_from_file_location("pdegym.common.vec_wrappers", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["pdegym.common.vec_wrappers"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["pdegym.common.vec_wrappers"], transforms


def load_reference_worker():
    """``(worker module, replay module)``: the reference's data-collection loop
    (``pdecontrol/mbrl/worker.py:39-93``) and its replay container, executed where they lie.
    ``worker.py`` imports two modules only for type annotations -- ``pdecontrol.mbrl.callbacks`` (pulls in
    wandb / matplotlib / seaborn) and ``pdecontrol.mbrl.world.wrappers`` (pulls in the surrogate models);
    those two get attribute-only stand-ins, ``Worker`` / ``PDEEnvStack`` / ``ExperienceReplay`` / ``Sample``
    are the reference's own code."""
    load_reference_wrappers()
    import importlib.util

    root = reference_root()
    for name, attr in (("pdecontrol.mbrl.callbacks", "PDECallback"),
                       ("pdecontrol.mbrl.world", None),
                       ("pdecontrol.mbrl.world.wrappers", "BaseWorldVecEnvWrapper")):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            if attr is None:
                mod.__path__ = []
            else:
                setattr(mod, attr, type(attr, (), {}))
            sys.modules[name] = mod

    def _load(name, relpath):
        if name not in sys.modules:
            spec = importlib.util.spec_from_file_location(name, os.path.join(root, relpath))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
        return sys.modules[name]

    replay = _load("pdecontrol.mbrl.replay", "pdecontrol/mbrl/replay.py")
    worker = _load("pdecontrol.mbrl.worker", "pdecontrol/mbrl/worker.py")
    return worker, replay


def make_reference_env(Xi=None, **config):
    """Construct the reference's ``KuramotoSivashinskyEnv`` (optionally with other jets).

    ``Xi`` is a class attribute in the reference (``kuramoto.py:18``); other jet layouts are
    obtained the only way the reference allows: by subclassing.  ``noop`` is hard-coded to 4
    jets (``kuramoto.py:62``) and ``reset`` steps with a 4-vector (``kuramoto.py:109``), so a
    J != 4 subclass can ``step`` but not ``reset``.
    """
    kuramoto, _ = load_reference()
    cls = kuramoto.KuramotoSivashinskyEnv
    if Xi is not None:
        cls = type("KuramotoSivashinskyEnvXi", (cls,), {"Xi": list(Xi)})
    return cls(**config)
