"""One-process-per-GPU plumbing: torch.distributed only carries the 128-byte NCCL id (and the bench's
barriers); the data path's single collective — ncclAllReduce(sum, 1 x fp64) of the squared error norm per
step attempt — is issued by libb200rk.so itself on the context stream.

The state vector shards contiguously: rank r owns [r*chunk, min(N, (r+1)*chunk)) with
chunk = ceil(N/world) rounded up to 4 elements; every work vector (y, k1..kS, tmp, yNew, lambda) uses the
same partition and never moves. Fixed-step methods need no collective at all.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _capi as capi


def shard_range(n_global: int, rank: int, world: int) -> tuple[int, int]:
    """(offset, length) of rank's contiguous shard — b200rk_shard_range (host-only, no GPU needed)."""
    off, ln = C.c_size_t(0), C.c_size_t(0)
    capi.check(capi.lib().b200rk_shard_range(n_global, rank, world, C.byref(off), C.byref(ln)))
    return off.value, ln.value


def exchange_unique_id(make_id, group=None) -> bytes:
    """Rank 0 calls make_id() (normally Context.nccl_unique_id); everyone returns the same 128 bytes.
    Works on any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
    import torch.distributed as dist

    box = [make_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise ValueError("NCCL unique id must be 128 bytes")
    return bytes(uid)


def init_context(local_rank: int | None = None):
    """Build the Context of this rank from the torch.distributed default group (or a single-GPU one)."""
    from .ode import Context

    local_rank = int(os.environ.get("LOCAL_RANK", "0")) if local_rank is None else local_rank
    try:
        import torch.distributed as dist
        active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except ImportError:
        active = False
    if not active:
        return Context(local_rank)
    uid = exchange_unique_id(Context.nccl_unique_id)
    return Context(local_rank, dist.get_rank(), dist.get_world_size(), uid)


def global_error_norm(local_sumsq: float, n_global: int, group=None) -> float:
    """Host-side statement of what the library does on the device: allreduce(sum) of the shard partials,
    then sqrt((1/N) * S) (ode.nim:64-65). Used by the gloo tests to pin the sharded arithmetic."""
    import math

    import torch
    import torch.distributed as dist

    t = torch.tensor([local_sumsq], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return math.sqrt(1.0 / float(n_global) * float(t.item()))
