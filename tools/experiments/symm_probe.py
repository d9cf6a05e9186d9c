"""Feasibility probe (torchrun, >= 2 GPUs): does torch's symmetric memory give peer + multicast pointers here?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    from torch._C._distributed_c10d import _SymmetricMemory
    print(rank, "has_multicast_support:", _SymmetricMemory.has_multicast_support(torch.device("cuda").type if False else "cuda", local) if True else None, flush=True)
except Exception as exc:
    print(rank, "has_multicast_support failed:", type(exc).__name__, exc, flush=True)
try:
    buf = symm_mem.empty(1 << 22, dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "multicast_ptr", hex(hdl.multicast_ptr), "signal pads", len(hdl.signal_pad_ptrs),
          "world", hdl.world_size, flush=True)
except Exception as exc:
    import traceback
    traceback.print_exc()
    print(rank, "symmetric memory failed:", type(exc).__name__, exc, flush=True)
dist.barrier()
dist.destroy_process_group()
