"""Host-side mirror of the reference interface: forcing transform, spaces, factories, sharding
arithmetic, and the loud failure without CUDA.  CPU only."""
import numpy as np
import pytest
import torch

import model_based_pde_control_b200 as pkg
from ks_testutil import load_golden
from model_based_pde_control_b200.forcing import GaussianForcing
from model_based_pde_control_b200.sharding import shard_range
from model_based_pde_control_b200.spaces import Box, batch_space


@pytest.mark.parametrize("name", ["kat1_default_1period", "kat3_large_1period", "attractor_n128_random",
                                  "attractor_n96_random"])
def test_forcing_matrix_bit_identical_to_reference(name):
    g = load_golden(name)
    f = GaussianForcing(g["x"], g["Xi"], float(g["sigma"]), float(g["L"]), int(g["N"]))
    assert f.matrix().dtype == np.float32 and np.array_equal(f.matrix(), g["F"])
    assert f.J == len(g["Xi"])


def test_forcing_call_and_inverse():
    g = load_golden("forcing_kat")
    x = np.linspace(0.0, 22.0 - 22.0 / 64, 64, dtype=np.float32)
    f = GaussianForcing(x, [0, .25, .5, .75], 0.4, 22.0, 64)
    phi = f(g["A"][:1])                       # numpy in -> numpy out, as Transform.convert/unconvert
    assert isinstance(phi, np.ndarray) and phi.dtype == np.float32 and np.array_equal(phi[0], g["phi"][0])
    t = f(torch.from_numpy(g["A"][:4]))
    assert isinstance(t, torch.Tensor) and np.array_equal(t.numpy(), g["phi"][:4])
    # Inverse samples the pattern at the jet positions and undoes the 4x4 mixing (transforms.py:267-279)
    back = f.Inverse(phi)
    assert np.allclose(back, g["A"][:1], atol=2e-5)
    assert f.Inverse.Inverse is f
    assert f.Inverse.xpos.tolist() == [0, 16, 32, 48]


def test_spaces_and_batching():
    b = Box(-1.0, 1.0, shape=(1, 4), dtype=np.float32)
    assert b.shape == (1, 4) and b.low.dtype == np.float32 and (b.high == 1).all()
    s = b.sample()
    assert s.shape == (1, 4) and s.dtype == np.float32 and b.contains(s)
    bb = batch_space(b, 10)
    assert bb.shape == (10, 1, 4) and bb.sample().shape == (10, 1, 4)
    o = Box(-np.inf, np.inf, shape=(1, 64), dtype=np.float32)
    assert np.isneginf(o.low).all() and o.sample().shape == (1, 64)


def test_shard_ranges_cover_and_are_contiguous():
    for n, w in [(4096, 1), (4096, 2), (65536, 8), (10, 4), (7, 8), (4097, 8)]:
        ranges = [shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(65536, 3, 8) == (24576, 32768)
    with pytest.raises(ValueError):
        shard_range(8, 8, 8)


def test_package_surface_and_factories_fail_loudly_without_cuda():
    assert pkg.ENV_ID == "KuramotoSivashinskyEnv-v0"
    with pytest.raises(ValueError):
        pkg.vector_make("SomethingElse-v0", num_envs=2)
    with pytest.raises(ValueError):
        pkg.make({}, new_step_api=False)
    with pytest.raises(TypeError):
        pkg.KSVecEnv(2, {"not_a_kwarg": 1})
    with pytest.raises(ValueError):
        pkg.KSVecEnv(2, precision="bf16")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            pkg.make({})
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            pkg.vector_make(pkg.ENV_ID, num_envs=4)


def test_register_uses_the_references_id_and_keywords(monkeypatch):
    """``register()`` = ``gym.envs.register(id="KuramotoSivashinskyEnv-v0", entry_point=..., order_enforce=False,
    new_step_api=True)`` as ``pdegym/kuramoto/__init__.py:26-31`` does; a no-op (False) without gym."""
    import sys
    import types
    from model_based_pde_control_b200 import registration

    if not registration.HAVE_GYM:
        assert registration.register() is False
    calls = []
    gym = types.ModuleType("gym")
    gym.envs = types.SimpleNamespace(registry={}, register=lambda **kw: calls.append(kw))
    monkeypatch.setitem(sys.modules, "gym", gym)
    monkeypatch.setattr(registration, "HAVE_GYM", True)
    assert registration.register() is True
    assert calls == [dict(id="KuramotoSivashinskyEnv-v0", entry_point="model_based_pde_control_b200.registration:make",
                          order_enforce=False, new_step_api=True)]
    gym.envs.registry[registration.ENV_ID] = object()          # e.g. pdegym was imported first
    assert registration.register() is False and len(calls) == 1
    assert registration.register(force=True) is True and len(calls) == 2
    mod, fn = registration.ENTRY_POINT.split(":")
    assert getattr(sys.modules[mod], fn) is registration.make


def test_local_time_limit_wrapper():
    from model_based_pde_control_b200.single_env import TimeLimit

    class Inner:
        observation_space = action_space = None
        metadata, reward_range, marker = {}, (0, 1), 7
        unwrapped = property(lambda self: self)

        def reset(self, **kw):
            return "obs"

        def step(self, a):
            return "obs", 1.0, False, False, {}

        def close(self):
            return "closed"

    env = TimeLimit(Inner(), 3)
    assert env.reset() == "obs" and env.marker == 7 and env.unwrapped is env.env
    assert [env.step(0)[3] for _ in range(4)] == [False, False, True, True]
    env.reset()
    assert env.step(0)[3] is False and env.close() == "closed"
    with pytest.raises(ValueError):
        TimeLimit(Inner(), 3, new_step_api=False)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure; the product path must not route through it."""
    import os
    import re

    root = os.path.dirname(os.path.abspath(pkg.__file__))
    for dirpath, _, files in os.walk(root):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "ks_oracle" not in text and "scipy" not in text, fn


def test_result_blocks_are_reused_only_when_no_result_references_them():
    """KSVecEnv.step returns views of pinned result blocks (zero-copy); a block is recycled only when
    no earlier result, nor any view derived from one, is alive.  Host logic only -- no GPU needed."""
    import types
    from model_based_pde_control_b200.env import KSVecEnv

    B, No = 6, 8
    offs = [0, 48, 48 + 192, 48 + 192 + 32, 48 + 192 + 32 + 16]
    fake = types.SimpleNamespace(num_envs=B, obs_len=No, _out_offsets=offs, _out_total=offs[4] + 16, copy=True,
                                 MAX_RESULT_BLOCKS=3, _block_refs=KSVecEnv._block_refs, _spill=None)
    fake._new_block = lambda: KSVecEnv._new_block(fake)
    fake._blocks = [fake._new_block()]
    free = lambda: KSVecEnv._free_block(fake)

    b0 = free()
    assert b0 is fake._blocks[0] and b0["owned"]
    obs = b0["obs"][...]                    # what step() hands out
    b1 = free()
    assert b1 is not b0 and len(fake._blocks) == 2          # obs alive -> block 0 must not be overwritten
    row = obs[2, 0]                          # a derived view keeps the block busy after `obs` is gone
    del obs
    assert free() is b1
    del row
    assert free() is b0                      # released -> recycled
    held = []
    for _ in range(3):                       # a caller that keeps everything
        held.append(free()["obs"][...])
    assert len(fake._blocks) == 3
    spill = free()
    assert not spill["owned"] and all(spill is not b for b in fake._blocks)   # -> copy-out path
    del held
    assert free()["owned"]


def test_forcing_object_inside_the_references_own_batch_transform():
    """``mbrl.py:157`` / ``evaluation/evaluate.py:88`` wrap ``env.forcing`` in the reference's ``BatchTransform`` (and use
    ``Operation([forcing, pdescaling])`` and ``.Inverse`` on it).  Our ``GaussianForcing`` inside the reference's own
    container classes gives, bit for bit, what the reference's ``GaussianForcing`` gives there."""
    from oracle import ref_loader

    if not ref_loader.reference_available():
        pytest.skip("reference tree not available")
    _, tr = ref_loader.load_reference()
    ref_env = ref_loader.make_reference_env()
    ours = GaussianForcing(ref_env.x, ref_env.Xi, ref_env.sigma, ref_env.L, ref_env.N)
    rng = np.random.default_rng(3)
    acts = rng.uniform(-1, 1, (7, 1, 4)).astype(np.float32)                  # [B, C, A] as the replay stores them
    a, b = tr.BatchTransform(ours), tr.BatchTransform(ref_env.forcing)
    assert np.array_equal(a(acts), b(acts)) and a(acts).shape == (7, 1, 64)
    t = torch.from_numpy(acts)
    assert torch.equal(a(t), b(t))
    phi = b(acts)
    assert np.array_equal(a.Inverse(phi), b.Inverse(phi))                    # BatchTransform._Inverse -> forcing.Inverse per sample
    assert np.allclose(a.Inverse(phi), acts, atol=2e-6)
    scal = tr.Normalize(aggregate=True, batched=True)
    scal.update(phi)
    op_a, op_b = tr.Operation([a, scal]), tr.Operation([b, scal])            # evaluate.py:91
    assert np.array_equal(op_a(acts), op_b(acts))
