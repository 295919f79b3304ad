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


def gather_rows(y_local, batch: int, group=None, dst: int | None = None):
    """The OPTIONAL collective of the path (north_star (d)): put the row blocks back together.

    dst=None: all-gather -- every rank gets the full (batch, n) array.
    dst=r:    gather to ONE rank ("the caller requests the output on one device"): rank r returns the full array,
              every other rank returns None.  Over NCCL this is send/recv fan-in across NVLink; it moves
              (world-1)/world of the output once, where the all-gather moves it world-1 times.
    Never on the timed hot path: filtering needs no communication at all."""
    import torch
    import torch.distributed as dist
    world, rank = world_rank(group)
    if world == 1:
        return y_local
    n = y_local.shape[1]
    blocks = [shard_rows(batch, world, r) for r in range(world)]
    if dst is None:
        cap = max(b - a for a, b in blocks)
        padded = torch.zeros((cap, n), dtype=y_local.dtype, device=y_local.device)
        padded[: y_local.shape[0]] = y_local
        buf = torch.empty((world * cap, n), dtype=y_local.dtype, device=y_local.device)
        dist.all_gather_into_tensor(buf, padded, group=group)
        return torch.cat([buf[r * cap: r * cap + (b - a)] for r, (a, b) in enumerate(blocks)], dim=0)
    if not 0 <= dst < world:
        raise ValueError(f"dst rank {dst} outside the group of {world}")
    # uneven shards: point-to-point fan-in straight into the destination's rows (no padding, no extra copy)
    if rank == dst:
        full = torch.empty((batch, n), dtype=y_local.dtype, device=y_local.device)
        a, b = blocks[rank]
        full[a:b] = y_local
        reqs = [dist.irecv(full[blocks[r][0]: blocks[r][1]], src=dist.get_global_rank(group, r) if group else r, group=group)
                for r in range(world) if r != dst and blocks[r][1] > blocks[r][0]]
        for q in reqs:
            q.wait()
        return full
    if y_local.shape[0] > 0:
        dist.isend(y_local.contiguous(), dst=dist.get_global_rank(group, dst) if group else dst, group=group).wait()
    return None
