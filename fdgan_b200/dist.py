"""Multi-GPU plumbing of the training step: the path shards by batch only.

One process per GPU (torchrun), each rank holds a full replica and a contiguous shard of the global batch;
the only exchange is ONE all-reduce (sum) per network per optimiser step over the flat fp32 gradient buffer
of the used parameters, followed by the 1/world scale folded into the fused Adam.  BatchNorm statistics stay
local, exactly as under the reference's nn.DataParallel (demo.py:89).  Device-agnostic on purpose: the same
functions run under gloo on CPU tensors in tests/ and under NCCL (NVLink 5 / NVSwitch) on the B200 box.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_range(global_batch: int, rank: int, world: int):
    """Contiguous images [lo, hi) of the global batch owned by ``rank``; the batch must divide evenly so that
    the mean over ranks of per-rank mean losses equals the global mean."""
    if global_batch % world != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def allreduce_flat_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a flat gradient buffer (a no-op in a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_flat_(flat: torch.Tensor, src=0, group=None) -> torch.Tensor:
    """Initial parameter broadcast from rank 0."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


def world_size(group=None) -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1
