"""Env sharding across GPUs: one process and one ``ks_handle`` per GPU, contiguous env ranges.

Every environment is independent, so the solver needs no exchange; the only collective is the
all-gather of observations / rewards / flags after each control period so that every learner rank
sees the whole batch.  The reference has no counterpart: it runs one
process per env through gym's ``AsyncVectorEnv`` (``pdecontrol/mbrl/mbrl.py:81-86``).

The sharding arithmetic and the gather layout are backend-agnostic (they work on CPU tensors
over ``gloo``, which is how they are unit-tested); the env itself needs CUDA.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def shard_range(num_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous ``[lo, hi)`` env range of ``rank``; the first ``num_envs % world_size`` ranks
    get one extra env.  ``world_size`` divides 4096 / 65536 in all BASELINE configs."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(num_envs, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_batch(local: dict, num_envs: int, group=None) -> dict:
    """All-gather per-env tensors (leading dim = local envs) into full-batch tensors, rank order
    = env order.  Equal shards use ``all_gather_into_tensor`` (one NCCL all-gather per tensor);
    ragged shards fall back to ``all_gather`` on padded buffers."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(num_envs, r, world)[1] - shard_range(num_envs, r, world)[0] for r in range(world)]
    out = {}
    for key, t in local.items():
        if t is None:
            out[key] = None
            continue
        assert t.shape[0] == sizes[rank], (key, t.shape, sizes[rank])
        if len(set(sizes)) == 1:
            full = torch.empty((num_envs,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(full, t.contiguous(), group=group)
        else:
            pad = max(sizes)
            buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            buf[: t.shape[0]] = t
            parts = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(parts, buf, group=group)
            full = torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
        out[key] = full
    return out


def gather_packed(packed_local: torch.Tensor, fields: dict, local_envs: int, group=None) -> dict:
    """ONE all-gather of every rank's packed output block (``KSVecEnv.step_device()["packed"]``,
    layout of ``ks_out_layout``) -- obs, reward, step, truncated and flags travel in a single NCCL
    collective.  Equal shards only.  Returns ``{name: tensor [world, local_envs, ...]}``: strided
    views into the gathered buffer in rank (= env) order; ``.reshape(world * local_envs, ...)``
    gives the flat batch (a copy)."""
    world = dist.get_world_size(group)
    total = packed_local.numel()
    full = torch.empty(world * total, dtype=torch.uint8, device=packed_local.device)
    dist.all_gather_into_tensor(full, packed_local, group=group)
    rows = full.view(world, total)
    out = {}
    for name, (off, dtype, shape) in fields.items():
        n = local_envs
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        out[name] = rows[:, off:off + nbytes].view(dtype).reshape((world, local_envs) + tuple(shape))
    return out


def connect_fused_gather(env, group=None) -> None:
    """Set up the fused all-gather of ``env`` (a local ``KSVecEnv`` shard) across ``group``: every
    rank allocates its gather buffer, the 64-byte CUDA-IPC handles travel through one host-side
    ``all_gather_object``, every rank maps its peers' buffers.  Equal shards only (checked).
    A failure on any rank (no peer access, different IPC namespaces, ...) raises on EVERY rank, after
    the ranks have agreed on it -- nobody is left waiting in a collective."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    err, handle = None, b""
    try:
        handle = env.gather_init(world, rank)
    except Exception as exc:      # noqa: BLE001 - agreed on below
        err = f"rank {rank}: {type(exc).__name__}: {exc}"
    infos = [None] * world
    dist.all_gather_object(infos, (env.num_envs, handle, err), group=group)
    errors = [e for _, _, e in infos if e]
    if errors:
        raise RuntimeError("fused gather: buffer allocation failed: " + "; ".join(errors))
    if len({n for n, _, _ in infos}) != 1:
        raise ValueError(f"fused gather needs equal shards, got {[n for n, _, _ in infos]}")
    try:
        env.gather_connect([h for _, h, _ in infos])
    except Exception as exc:      # noqa: BLE001
        err = f"rank {rank}: {type(exc).__name__}: {exc}"
    outcomes = [None] * world
    dist.all_gather_object(outcomes, err, group=group)
    errors = [e for e in outcomes if e]
    if errors:
        raise RuntimeError("fused gather: peer mapping failed: " + "; ".join(errors))
    dist.barrier(group)      # nobody launches a peer-writing kernel before everyone is mapped


def connect_fused_gather_symm(env, group=None, multicast: bool = True) -> dict:
    """The fused exchange on ``torch.distributed._symmetric_memory`` buffers instead of CUDA IPC: every rank
    allocates its gather buffer with ``symm_mem.empty``, ``symm_mem.rendezvous`` maps all of them into every
    process and -- on an NVSwitch system -- binds them to one NVLS multicast address.  With ``multicast`` the
    period kernel then stores its observation rows once to that address (``multimem.st``; the switch replicates
    them into every GPU) instead of once per peer.  Returns ``{"multicast": bool, ...}``; raises on every rank
    if any rank fails.  PyTorch is plumbing here (allocation + mapping); the stores and the handshake are the
    library's own kernels."""
    import torch.distributed._symmetric_memory as symm_mem

    group = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def agree(stage, err, payload=None):
        """Every rank learns every rank's outcome of `stage`; any failure raises on ALL ranks (nobody is left
        alone inside the next collective)."""
        outcomes = [None] * world
        dist.all_gather_object(outcomes, (err, payload), group=group)
        errors = [e for e, _ in outcomes if e]
        if errors:
            raise RuntimeError(f"fused gather (symmetric memory), {stage}: " + "; ".join(errors))
        return [p for _, p in outcomes]

    # 1. local allocation (no collective inside)
    err, buf, total = None, None, 0
    try:
        _slot, total = env.gather_layout(world)
        buf = symm_mem.empty(total, dtype=torch.uint8, device=env.device)
    except Exception as exc:      # noqa: BLE001 - agreed on below
        err = f"rank {rank}: {type(exc).__name__}: {exc}"
    sizes = agree("allocation", err, (env.num_envs, total))
    if len(set(sizes)) != 1:
        raise ValueError(f"fused gather needs equal shards, got {[n for n, _ in sizes]}")
    # 2. rendezvous (collective: every rank is known to arrive) + attach
    err, mc = None, 0
    try:
        hdl = symm_mem.rendezvous(buf, group)
        mc = int(hdl.multicast_ptr) if multicast else 0
        env.gather_attach(world, rank, [int(p) for p in hdl.buffer_ptrs], mc, total)
        env._symm = (buf, hdl)            # keep the mapping alive as long as the env
    except Exception as exc:      # noqa: BLE001
        err = f"rank {rank}: {type(exc).__name__}: {exc}"
    flags = agree("rendezvous / attach", err, bool(mc))
    torch.cuda.synchronize(env.device)
    dist.barrier(group)      # every buffer zero-filled and attached before anybody's kernel writes into a peer
    return {"multicast": bool(multicast) and all(flags), "bytes": total}


class ShardedKSVecEnv:
    """``num_envs`` environments spread over the ranks of a ``torch.distributed`` group.

    ``step_device(actions)`` takes the FULL action batch ``[num_envs, J]`` (every rank holds the
    replicated policy output), steps the local shard, and all-gathers the results so that every
    rank returns full-batch tensors.  States and observations equal the single-GPU run bit for bit because
    an env's arithmetic does not depend on where it lives; so do the rewards when ``points_per_lane=`` pins the
    lane layout (their summation order follows the layout, which the library otherwise picks from the LOCAL
    batch size -- a shard may get another layout than the whole batch: equal to 1e-15 then).

    ``env_factory(local_num_envs)`` builds the local env (default: ``KSVecEnv``).
    """

    def __init__(self, num_envs: int, config: Optional[dict] = None, group=None,
                 env_factory: Optional[Callable] = None, **kwargs):
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.num_envs = num_envs
        self.lo, self.hi = shard_range(num_envs, self.rank, self.world_size)
        if env_factory is None:
            from .env import KSVecEnv

            # env_index_base: the device generator's counter uses the GLOBAL env index, so a sharded
            # reset_device(seed) equals the single-GPU one whatever the world size
            env_factory = lambda n: KSVecEnv(n, config, env_index_base=self.lo, **kwargs)  # noqa: E731
        self.local = env_factory(self.hi - self.lo)
        self._fused = False

    @property
    def local_num_envs(self) -> int:
        return self.hi - self.lo

    def local_slice(self, full: torch.Tensor) -> torch.Tensor:
        """This rank's rows of a full-batch tensor (actions, initial conditions, ...)."""
        return full[self.lo:self.hi]

    def gather(self, local: dict) -> dict:
        if self.world_size == 1:
            return dict(local)
        return gather_batch(local, self.num_envs, self.group)

    def step_device(self, actions: torch.Tensor, gather: bool = True) -> dict:
        """Step the local shard with this rank's rows of the full action batch.  ``gather=True``
        returns full-batch tensors ``[num_envs, ...]``; ``gather="packed"`` uses the single-
        collective path and returns ``[world, local_envs, ...]`` views (equal shards);
        ``gather="fused"`` / ``"fused_mc"`` return the same views filled by the kernel epilogues themselves
        (peer stores over NVLink + epoch handshake, ``ks_step_gather``) -- no NCCL call per period."""
        if gather in ("fused", "fused_mc") and self.world_size > 1:
            # kernel epilogue stores straight into every peer's buffer over NVLink (no collective call).
            # "fused": CUDA-IPC mapped peer buffers, one store per peer; "fused_mc": torch symmetric memory with an
            # NVLS multicast address, observation rows sent once (7 us per period less on 8 GPUs).  The transport is
            # fixed by the first call.  A peer that did not arrive within KS_GATHER_TIMEOUT_S makes the NEXT call
            # raise KsError (the handshake kernel poisons the incomplete block and sets a host-visible sticky word).
            if not self._fused:
                if gather == "fused_mc":
                    connect_fused_gather_symm(self.local, self.group)
                else:
                    connect_fused_gather(self.local, self.group)
                self._fused = True
            return self.local.step_gather(self.local_slice(actions.reshape(self.num_envs, -1)))
        out = self.local.step_device(self.local_slice(actions.reshape(self.num_envs, -1)))
        if not gather:
            return out
        if gather in ("packed", "fused", "fused_mc"):
            if self.world_size > 1:
                return gather_packed(out["packed"], self.local.packed_fields(), self.local_num_envs, self.group)
            # world of one: the same [world, local_envs, ...] view shape as the multi-rank paths
            return {k: v.unsqueeze(0) for k, v in out.items()}
        out = {k: v for k, v in out.items() if k != "packed"}
        return self.gather(out)

    def reset_device(self, seed: Optional[int] = None, **kwargs) -> None:
        """Device-drawn initial conditions + burn-in on every shard.  The generator is keyed by (seed;
        point, GLOBAL env index), so with one shared ``seed`` the batch equals the single-GPU
        ``reset_device(seed)`` bit for bit, for every world size.  ``seed=None``: rank 0 draws one from the
        OS and broadcasts it."""
        if seed is None:
            import os

            box = [int.from_bytes(os.urandom(8), "little")]
            if self.world_size > 1:
                dist.broadcast_object_list(box, src=0, group=self.group)
            seed = box[0]
        self.local.reset_device(seed=seed, **kwargs)

    def set_state(self, u_full, timestep_full=None) -> None:
        ts = None if timestep_full is None else timestep_full[self.lo:self.hi]
        self.local.set_state(u_full[self.lo:self.hi], ts)

    def close(self) -> None:
        self.local.close()
