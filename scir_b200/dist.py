"""One-process-per-GPU front end: channels (rows) sharded across ranks, no collective on the data path.

Rows of a batched FIR are independent (crates/scir-gpu/src/lib.rs:1138-1140 carries no cross-row
state), so rank r of `world` owns the contiguous block scir_b200_shard_rows(batch, world, r) -- the
same integer rule the in-process scir_b200_mg_* front end uses -- and filters it on its own GPU.
The only collective is OPTIONAL: `gather_rows` (NCCL all-gather over NVLink on GPUs, gloo on CPU
tensors in the tests) for callers that want the whole output everywhere.  It costs ~18x the
per-GPU compute at BASELINE config 2, so it is never on the timed path.
"""
from __future__ import annotations

import ctypes as C

from . import _lib as L
from .gpu import _check, _load


def shard_rows(batch: int, world: int, rank: int) -> tuple[int, int]:
    a, b = C.c_int64(), C.c_int64()
    _check(_load().scir_b200_shard_rows(int(batch), int(world), int(rank), C.byref(a), C.byref(b)))
    return a.value, b.value


def world_rank(group=None) -> tuple[int, int]:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def local_block(x_global, group=None):
    """The row block of a replicated (batch, n) array this rank owns."""
    world, rank = world_rank(group)
    r0, r1 = shard_rows(x_global.shape[0], world, rank)
    return x_global[r0:r1]


def fir1d_batched_f32_sharded(x_local, taps, **kw):
    """Filter this rank's rows on this rank's GPU.  No communication."""
    from . import gpu
    return gpu.fir1d_batched_f32_cuda(x_local, taps, **kw)


def gather_rows(y_local, batch: int, group=None):
    """All-gather the row blocks back into the full (batch, n) array on every rank (opt-in)."""
    import torch
    import torch.distributed as dist
    world, rank = world_rank(group)
    if world == 1:
        return y_local
    n = y_local.shape[1]
    blocks = [shard_rows(batch, world, r) for r in range(world)]
    cap = max(b - a for a, b in blocks)
    padded = torch.zeros((cap, n), dtype=y_local.dtype, device=y_local.device)
    padded[: y_local.shape[0]] = y_local
    buf = torch.empty((world * cap, n), dtype=y_local.dtype, device=y_local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * cap: r * cap + (b - a)] for r, (a, b) in enumerate(blocks)], dim=0)
