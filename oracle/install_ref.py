"""TEST / BASELINE INFRASTRUCTURE ONLY -- installs the UNMODIFIED reference into ``baseline/_ref``.

``baseline/_ref`` is git-ignored (the reference's sources never enter this repository's history)
but NOT gpurun-ignored, so the installed copy travels to the GPU box, where ``/root/reference``
does not exist.  There it serves two purposes, both through ``oracle/ref_loader.py``:

* ``bench.py --impl reference`` and ``cpu_baseline`` time the reference's own
  ``KuramotoSivashinskyEnv.step`` (``pdegym/kuramoto/kuramoto.py:78-98``), ``kind: "reference"``;
* ``-m gpu`` tests drive the reference's own ``vec_wrappers`` / ``Worker.rollout`` over a real
  ``KSVecEnv``.

Recipe (run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present):

    python -m pip install --no-index --no-build-isolation --no-deps --ignore-requires-python \
        --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference under /tmp>

Two things stand between the stock command and a working install, neither of them source code:
the reference's ``pyproject.toml`` names ``poetry-core`` as its build backend (not installed, not in
``/opt/wheelhouse``) and pins ``python <3.9``.  The copy under ``/tmp`` therefore gets a
four-line ``pyproject.toml`` that selects setuptools as the backend; every ``.py`` file is installed
byte for byte, which ``verify()`` checks against ``/root/reference`` with SHA-256.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
SOURCE = "/root/reference"
PACKAGES = ("pdegym", "pdecontrol")

_PYPROJECT = """[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"

[project]
name = "pdecontrol"
version = "0.1.0"

[tool.setuptools.packages.find]
include = ["pdegym*", "pdecontrol*"]
"""


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def _py_files(root: str):
    for pkg in PACKAGES:
        for dirpath, _dirs, files in os.walk(os.path.join(root, pkg)):
            for name in files:
                if name.endswith(".py"):
                    full = os.path.join(dirpath, name)
                    yield os.path.relpath(full, root), full


def verify(source: str = SOURCE, target: str = TARGET) -> int:
    """Every ``.py`` of the reference is present in ``target`` with identical bytes; returns the count."""
    n = 0
    for rel, full in _py_files(source):
        inst = os.path.join(target, rel)
        if not os.path.isfile(inst) or _sha(inst) != _sha(full):
            raise RuntimeError(f"baseline/_ref: {rel} missing or different from the reference")
        n += 1
    return n


def installed(target: str = TARGET) -> bool:
    return os.path.isfile(os.path.join(target, "pdegym", "kuramoto", "kuramoto.py"))


def install(source: str = SOURCE, target: str = TARGET, force: bool = False) -> str:
    """Install ``source`` into ``target``; returns a one-line outcome (also written to
    ``baseline/_ref/INSTALL_OUTCOME.txt``)."""
    if not os.path.isdir(source):
        return f"skipped: {source} not present (using the prebuilt {target})" if installed(target) \
            else f"skipped: neither {source} nor {target} present"
    if installed(target) and not force:
        try:
            n = verify(source, target)
            return f"up to date: {n} reference .py files in baseline/_ref, byte-identical"
        except RuntimeError:
            pass
    tmp = tempfile.mkdtemp(prefix="ks_refcopy_")
    try:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(source, copy, ignore=shutil.ignore_patterns(".git", "assets"))
        with open(os.path.join(copy, "pyproject.toml"), "w") as f:
            f.write(_PYPROJECT)
        shutil.rmtree(target, ignore_errors=True)
        os.makedirs(os.path.dirname(target), exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--ignore-requires-python", "--find-links", "/opt/wheelhouse", "--target", target, copy]
        subprocess.run(cmd, check=True, cwd=tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    n = verify(source, target)
    outcome = (f"installed: pip --target baseline/_ref from a /tmp copy with a setuptools build backend "
               f"(poetry-core absent); {n} .py files byte-identical to {source}")
    with open(os.path.join(target, "INSTALL_OUTCOME.txt"), "w") as f:
        f.write(outcome + "\n")
    return outcome


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
