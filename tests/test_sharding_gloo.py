"""world_size-2 (and 3, ragged) gloo tests of the env sharding + all-gather layout on CPU.

The local env is a deterministic stub (the CUDA env needs a GPU); what is tested is exactly what
runs at N>1 on the GPU box around the kernel: which rows each rank steps and that the gathered
full-batch tensors come back in env order on every rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from model_based_pde_control_b200.sharding import ShardedKSVecEnv, shard_range


class StubEnv:
    """step_device: obs[b] = 1000*global_id... derived only from the action rows it was given."""

    def __init__(self, n, N=8):
        self.num_envs, self.N = n, N
        self.calls = 0

    def step_device(self, actions):
        assert actions.shape[0] == self.num_envs
        self.calls += 1
        key = actions[:, 0].double()
        return {"obs": (key[:, None] * 10 + torch.arange(self.N)[None]).float(),
                "reward": -key, "truncated": (key.long() % 2).to(torch.uint8), "nothing": None}

    def set_state(self, u, ts=None):
        self.u = u

    def close(self):
        pass


def _worker(rank, world, port, num_envs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        env = ShardedKSVecEnv(num_envs, env_factory=lambda n: StubEnv(n))
        lo, hi = shard_range(num_envs, rank, world)
        assert (env.lo, env.hi, env.local_num_envs) == (lo, hi, hi - lo)
        actions = torch.arange(num_envs, dtype=torch.float32)[:, None].repeat(1, 4)   # row b carries id b
        out = env.step_device(actions)
        ids = torch.arange(num_envs, dtype=torch.float64)
        ok = (torch.equal(out["reward"], -ids)
              and torch.equal(out["obs"], (ids[:, None] * 10 + torch.arange(8)[None]).float())
              and torch.equal(out["truncated"], (ids.long() % 2).to(torch.uint8))
              and out["nothing"] is None and env.local.calls == 1)
        local = env.step_device(actions, gather=False)
        ok = ok and local["reward"].shape[0] == hi - lo and torch.equal(local["reward"], -ids[lo:hi])
        if num_envs % world == 0:
            # single-collective path: per-rank packed block [reward f64 | obs f32 | flag u8], 16-byte aligned parts
            n = hi - lo
            off_obs = (8 * n + 15) // 16 * 16
            off_flag = off_obs + (4 * n * 8 + 15) // 16 * 16
            block = torch.zeros(off_flag + (n + 15) // 16 * 16, dtype=torch.uint8)
            block[:8 * n].view(torch.float64).copy_(local["reward"])
            block[off_obs:off_obs + 4 * n * 8].view(torch.float32).copy_(local["obs"].reshape(-1))
            block[off_flag:off_flag + n].copy_(local["truncated"])
            from model_based_pde_control_b200.sharding import gather_packed
            g = gather_packed(block, {"reward": (0, torch.float64, ()), "obs": (off_obs, torch.float32, (8,)),
                                      "truncated": (off_flag, torch.uint8, ())}, n)
            ok = ok and g["reward"].shape == (world, n) and torch.equal(g["reward"].reshape(-1), -ids)
            ok = ok and torch.equal(g["obs"].reshape(num_envs, 8), out["obs"])
            ok = ok and torch.equal(g["truncated"].reshape(-1), out["truncated"])
        env.set_state(torch.arange(num_envs * 8.0).reshape(num_envs, 8))
        ok = ok and torch.equal(env.local.u, torch.arange(num_envs * 8.0).reshape(num_envs, 8)[lo:hi])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,num_envs", [(2, 16), (2, 4096), (3, 10)])
def test_sharded_step_and_gather(world, num_envs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=5) for _ in range(world))
    assert results == {r: True for r in range(world)}


def test_single_process_passthrough():
    env = ShardedKSVecEnv(6, env_factory=lambda n: StubEnv(n))
    assert (env.rank, env.world_size, env.lo, env.hi) == (0, 1, 0, 6)
    out = env.step_device(torch.arange(6.0)[:, None].repeat(1, 4))
    assert out["reward"].tolist() == [-0.0, -1.0, -2.0, -3.0, -4.0, -5.0]


# ---- fused-gather orchestration (host side): handle exchange in rank order, equal-shard check --------------
class StubGatherEnv(StubEnv):
    """Records what the host-side set-up of the fused gather hands to the C ABI (the peer stores
    themselves need NVLink; tests/test_gpu_multi.py covers them on 2 GPUs)."""

    def gather_init(self, world, rank):
        self.world, self.rank = world, rank
        return bytes([rank]) * 64                     # a 64-byte "IPC handle" that names its owner

    def gather_connect(self, handles):
        self.handles = [bytes(h) for h in handles]

    def step_gather(self, actions):
        return {"rows": actions.shape[0], "handles": self.handles}


def _fused_worker(rank, world, port, num_envs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        env = ShardedKSVecEnv(num_envs, env_factory=lambda n: StubGatherEnv(n))
        actions = torch.zeros(num_envs, 4)
        try:
            out = env.step_device(actions, gather="fused")
            ok = (env.local.world, env.local.rank) == (world, rank) and out["rows"] == env.local_num_envs
            ok = ok and out["handles"] == [bytes([r]) * 64 for r in range(world)]
            env.step_device(actions, gather="fused")      # connects only once
            q.put((rank, bool(ok)))
        except ValueError as exc:                          # ragged shards are refused on every rank
            q.put((rank, "equal shards" in str(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,num_envs", [(2, 16), (3, 12), (3, 10)])
def test_fused_gather_setup(world, num_envs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fused_worker, args=(r, world, port, num_envs, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=5) for _ in range(world))
    assert results == {r: True for r in range(world)}


class FailingGatherEnv(StubGatherEnv):
    def __init__(self, n, fail_rank):
        super().__init__(n)
        self.fail_rank = fail_rank

    def gather_connect(self, handles):
        if self.rank == self.fail_rank:
            raise OSError("no peer access")
        super().gather_connect(handles)


def _failing_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from model_based_pde_control_b200.sharding import connect_fused_gather
        try:
            connect_fused_gather(FailingGatherEnv(8, fail_rank=1))
            q.put((rank, False))
        except RuntimeError as exc:       # every rank, not only the one that failed
            q.put((rank, "rank 1" in str(exc) and "no peer access" in str(exc)))
        dist.barrier()                    # and nobody is stuck in a half-finished collective
    finally:
        dist.destroy_process_group()


def test_fused_gather_setup_failure_is_raised_on_every_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_failing_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(q.get(timeout=5) for _ in range(3)) == {0: True, 1: True, 2: True}


class SymmLayoutFailsEnv(StubGatherEnv):
    """For ``connect_fused_gather_symm``: the local allocation stage fails on one rank only."""

    device = torch.device("cpu")

    def __init__(self, n, fail_rank):
        super().__init__(n)
        self.fail_rank = fail_rank

    def gather_layout(self, world):
        if dist.get_rank() == self.fail_rank:
            raise OSError("no memory for the gather buffer")
        return 256, 2 * world * 256 + 256


def _symm_failing_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from model_based_pde_control_b200.sharding import connect_fused_gather_symm
        try:
            connect_fused_gather_symm(SymmLayoutFailsEnv(8, fail_rank=2))
            q.put((rank, False))
        except RuntimeError as exc:       # raised on EVERY rank before anybody enters the rendezvous collective
            q.put((rank, "allocation" in str(exc) and "rank 2" in str(exc)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_symmetric_memory_setup_failure_is_agreed_on_before_the_rendezvous():
    """One rank failing in the local allocation stage must not leave the others alone inside
    ``symm_mem.rendezvous`` (a collective): the outcome of every stage is exchanged first."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_symm_failing_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(q.get(timeout=5) for _ in range(3)) == {0: True, 1: True, 2: True}
