"""Multi-GPU checks (need >= 2 B200s on one box: ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu``;
skipped on a 1-GPU box).  Each test launches ``torchrun`` with one rank per GPU.

* fused all-gather (``ks_step_gather``: the period kernel's epilogue stores into every peer's buffer
  over NVLink + epoch handshake) == the NCCL all-gather of the packed block, bit for bit, over many
  consecutive periods (double-buffer / handshake hazards would show as stale rows);
* sharded trajectories == the single-GPU trajectory of the same envs, bit for bit.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["KS_ROOT"])
import numpy as np, torch, torch.distributed as dist
from model_based_pde_control_b200 import KSVecEnv
from model_based_pde_control_b200.sharding import connect_fused_gather, gather_packed, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
solver = os.environ.get("KS_SOLVER", "fd_rk4")
cfg = dict(dt=0.025, cfg_steps=10) if solver == "etdrk4" else dict(cfg_steps=25)
B_total, K = (1020 if 1020 % world == 0 else 127 * world), 12      # equal shards (the fused exchange requires them);
                                                                   # 510 envs per rank at world 2: the last thread block has an early-exit warp
lo, hi = shard_range(B_total, rank, world)
rng = np.random.default_rng(0)
u0 = rng.uniform(-1, 1, (B_total, 64))
acts = torch.as_tensor(rng.uniform(-1, 1, (K, B_total, 4)).astype(np.float32)).to(dev)

ppl = int(os.environ.get("KS_PPL", "0"))                           # spectral solver: 8 = 8-lane layout, 0 = automatic (16-lane here)
env = KSVecEnv(hi - lo, cfg, device=local, solver=solver, points_per_lane=ppl)          # fused path
ref = KSVecEnv(hi - lo, cfg, device=local, solver=solver, points_per_lane=ppl)          # NCCL path
env.set_state(u0[lo:hi], 0); ref.set_state(u0[lo:hi], 0)
if os.environ.get("KS_SYMM"):      # symmetric memory + NVLS multicast stores instead of CUDA IPC + unicast
    from model_based_pde_control_b200.sharding import connect_fused_gather_symm
    info = connect_fused_gather_symm(env)
    if rank == 0:
        print("SYMM", info, flush=True)
else:
    connect_fused_gather(env)
fields = ref.packed_fields()
ok = True
for k in range(K):
    if k % 3 == 1:
        env.gather_barrier()           # device-side rendezvous with its own flag words: must not disturb the exchange's epochs
    g = env.step_gather(acts[k, lo:hi])
    o = ref.step_device(acts[k, lo:hi])
    n = gather_packed(o["packed"], fields, hi - lo)
    for name in ("obs", "reward", "step", "truncated", "nonfinite"):
        ok = ok and torch.equal(g[name], n[name])
    ok = ok and int(g["step"].min()) == k + 1 == int(g["step"].max())
assert not env.gather_timed_out()

# the gathered batch equals the single-GPU run of all envs (rank 0 computes it)
if rank == 0:
    # same lane layout as the shards: the STATE is bitwise independent of the layout, the reward's summation order
    # follows it (the chooser picks P from the batch size, and a 510-env shard gets another P than 1020 envs)
    full = KSVecEnv(B_total, cfg, device=local, solver=solver, points_per_lane=env.launch_info()["points_per_lane"])
    full.set_state(u0, 0)
    for k in range(K):
        f = full.step_device(acts[k])
    ok = ok and torch.equal(f["obs"], g["obs"].reshape(B_total, -1)) and torch.equal(f["reward"], g["reward"].reshape(-1))
    full.close()

# sharded reset_device(seed) == the single-GPU reset_device(seed), bit for bit, whatever the world size
# (the device generator is keyed by the GLOBAL env index)
from model_based_pde_control_b200 import ShardedKSVecEnv
sh = ShardedKSVecEnv(B_total, cfg, device=local, solver=solver, burnin_periods=3, ic="device")
sh.reset_device(seed=77)
one = KSVecEnv(B_total, cfg, device=local, solver=solver, burnin_periods=3, ic="device")
one.reset_device(seed=77)
us, _ = sh.local.get_state_device()
uo, _ = one.get_state_device()
ok_reset = torch.equal(us, uo[lo:hi]) and float(us.abs().max()) > 0
if solver == "etdrk4":      # two envs share a transform: bits are sharding-independent for even shard starts only
    ok_reset = ok_reset or (lo % 2 == 1)
ok = ok and ok_reset
sh.close(); one.close()

# a peer that arrives late (beyond KS_GATHER_TIMEOUT_S) must not go unnoticed: the waiting rank's NEXT
# ks_step_gather fails, its block for that period carries poisoned flags, and ks_gather_clear re-arms
if os.environ.get("KS_TEST_TIMEOUT") and world == 2:
    import time
    from model_based_pde_control_b200._lib import KsError
    torch.cuda.synchronize(); dist.barrier()
    if rank == 1:
        time.sleep(2.5)                          # > KS_GATHER_TIMEOUT_S = 0.5
    g = env.step_gather(acts[0, lo:hi]); o = ref.step_device(acts[0, lo:hi])
    torch.cuda.synchronize()
    if rank == 0:
        assert env.gather_timed_out()
        assert bool((g["nonfinite"][1] == 0xFF).all()), "the late peer's slot must be poisoned"
        try:
            env.step_gather(acts[1, lo:hi])
            ok = False
        except KsError as exc:
            ok = ok and exc.code == -4
    dist.barrier()
    env.gather_clear()
    if rank == 0:
        assert not env.gather_timed_out()
    dist.barrier()
    g = env.step_gather(acts[1, lo:hi]); o = ref.step_device(acts[1, lo:hi])
    n = gather_packed(o["packed"], fields, hi - lo)
    ok = ok and all(torch.equal(g[name], n[name]) for name in ("obs", "reward", "step", "truncated", "nonfinite"))

t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_OK" if int(t.item()) == 1 else "MULTI_GPU_MISMATCH", flush=True)
env.close(); ref.close()
dist.barrier(); dist.destroy_process_group()
'''


def _run(world, solver, timeout_case=False, ppl=0, symm=False):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    env = dict(os.environ, KS_ROOT=ROOT, KS_SOLVER=solver, KS_PPL=str(ppl))
    if timeout_case:
        env.update(KS_TEST_TIMEOUT="1", KS_GATHER_TIMEOUT_S="0.5")
    if symm:
        env.update(KS_SYMM="1")
    path = os.path.join(ROOT, "gpurun_out", f"_multi_worker_{solver}.py")     # torchrun needs a script file
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", path]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    errs = "\n".join(ln for ln in res.stderr.splitlines() if "Error" in ln or "error" in ln or "assert" in ln)[:3000]
    assert res.returncode == 0, res.stdout[-2000:] + errs + res.stderr[-3000:]
    assert "MULTI_GPU_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.parametrize("solver,ppl", [("fd_rk4", 0), ("etdrk4", 0), ("etdrk4", 8)])
def test_fused_gather_equals_nccl_gather_and_single_gpu_2gpus(solver, ppl):
    _run(2, solver, ppl=ppl)


def test_fused_gather_4gpus():
    _run(4, "fd_rk4")


def test_fused_gather_8gpus():
    _run(8, "fd_rk4")


@pytest.mark.parametrize("solver", ["fd_rk4", "etdrk4"])
def test_fused_gather_on_symmetric_memory_with_multicast_2gpus(solver):
    """ks_gather_attach: the same exchange on torch symmetric-memory buffers; the FD-RK4 kernel sends its
    observation rows through the NVLS multicast address (multimem.st), the spectral kernels stay unicast."""
    _run(2, solver, symm=True)


def test_fused_gather_late_peer_is_reported_not_ignored():
    _run(2, "fd_rk4", timeout_case=True)


def test_fused_gather_world1_equals_step_device():
    """World size 1 (runs on any GPU box): the gather path with no peers is the plain step."""
    import numpy as np
    import torch

    from model_based_pde_control_b200 import KSVecEnv

    B = 37
    rng = np.random.default_rng(1)
    u0 = rng.uniform(-1, 1, (B, 64))
    a = torch.as_tensor(rng.uniform(-1, 1, (3, B, 4)).astype(np.float32)).cuda()
    e1, e2 = KSVecEnv(B, cfg_steps=20), KSVecEnv(B, cfg_steps=20)
    e1.set_state(u0, 0); e2.set_state(u0, 0)
    h = e1.gather_init(1, 0)
    assert len(h) == 64
    e1.gather_connect([h])
    for k in range(3):
        g = e1.step_gather(a[k])
        o = e2.step_device(a[k])
        for name in ("obs", "reward", "step", "truncated", "nonfinite"):
            assert g[name].shape[0] == 1 and torch.equal(g[name][0], o[name])
    assert not e1.gather_timed_out()
    e1.close(); e2.close()
