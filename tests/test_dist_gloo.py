"""world_size-2 gloo test of the N>1 host logic on CPU: row sharding + the optional gather.

No CUDA here, so each rank filters its shard with the ORACLE (test infrastructure standing in for
the device) -- what is under test is scir_b200.dist: the shard rule (shared with the C ABI's
scir_b200_shard_rows / scir_b200_mg_*) and the all-gather reassembly, including uneven shards.
"""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist          # noqa: E402
import torch.multiprocessing as mp        # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, n, k, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from scir_b200 import dist as sdist
        rng = np.random.RandomState(123)
        x = (rng.rand(batch, n).astype(np.float32) * 2 - 1)          # replicated input
        taps = rng.randn(k).astype(np.float32)
        xl = sdist.local_block(torch.from_numpy(x))
        r0, r1 = sdist.shard_rows(batch, world, rank)
        assert xl.shape[0] == r1 - r0
        yl = torch.from_numpy(O.fir1d_batched_f32(xl.numpy(), taps)) if xl.shape[0] else torch.zeros((0, n))
        y = sdist.gather_rows(yl, batch)
        want = O.fir1d_batched_f32(x, taps)
        ok = bool(np.array_equal(y.numpy(), want))
        # gather to ONE rank (north_star (d): "when the caller requests the output on one device")
        for dst in range(world):
            y1 = sdist.gather_rows(yl, batch, dst=dst)
            ok = ok and ((y1 is None) if rank != dst else bool(np.array_equal(y1.numpy(), want)))
        q.put((rank, ok, (r0, r1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7, 1])
def test_row_sharding_and_gather_world2(batch):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, 257, 9, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    blocks = sorted(b for _, _, b in res)
    assert blocks[0][0] == 0 and blocks[-1][1] == batch and blocks[0][1] == blocks[1][0]
