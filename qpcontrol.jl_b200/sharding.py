"""Instance sharding across the GPUs of one box.

The reference has no cross-instance coupling of any kind (one mutable controller = one state, one QP:
reference src/lowlevel/momentum.jl:1-13,41-81), so the batch splits into contiguous blocks, one per rank / GPU, with
NO collective on the data path.  `torch.distributed` is used only as plumbing: the barrier around the timed region,
the max-over-ranks of the device time, and (tests, optional) gathering results on rank 0.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly."""
    def _i(name, default):
        try:
            return int(os.environ.get(name, default))
        except ValueError:
            return default
    return _i("RANK", 0), _i("WORLD_SIZE", 1), _i("LOCAL_RANK", 0)


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of a batch of B instances owned by `rank`: the first B % world ranks get one extra."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    base, extra = divmod(int(B), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_process_group(backend: str, local_rank: int = 0):
    """One process per GPU (backend "nccl") or per CPU worker (backend "gloo", tests); rendezvous on 127.0.0.1."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend)


def barrier(device=None):
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and torch.cuda.is_available():
        torch.cuda.synchronize(device)


def max_over_ranks(values, device=None) -> np.ndarray:
    """Element-wise maximum of a small float vector over all ranks (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(np.asarray(values, dtype=np.float64), dtype=torch.float64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()


def gather_rows(local: np.ndarray, B: int, dst: int = 0) -> Optional[np.ndarray]:
    """Concatenate every rank's block (in rank order) on `dst`; used by tests and by callers that want the whole
    batch in one place -- not part of the timed path."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = [None] * world if rank == dst else None
    dist.gather_object(np.ascontiguousarray(local), parts, dst=dst)
    if rank != dst:
        return None
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == B, (out.shape, B)
    return out
