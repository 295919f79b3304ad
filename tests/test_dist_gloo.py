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


# ---- bench.py's cross-rank agreement helpers (world_size 2, gloo) ---------------------------------------------------
# Round 2 lost an 8-GPU run to a rank-local `if rejected(clocks): barrier()`: one GPU's clock sample re-measured alone and
# the job deadlocked.  Every decision that changes how many collectives a rank calls goes through Env.any_rank now.
def _agree_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        import bench
        env = bench.Env.__new__(bench.Env)                  # the helpers only need these attributes (no CUDA here)
        env.torch, env.dist, env.dev, env.world, env.rank = torch, dist, torch.device("cpu"), world, rank
        ok = env.any_rank(rank == 1) is True                # true on ONE rank => true on every rank
        ok = ok and env.any_rank(False) is False
        ok = ok and env.all_ranks(10.0 + rank) == [10.0, 11.0]
        ok = ok and env.max_over_ranks(float(rank)) == 1.0 and env.min_over_ranks(float(rank)) == 0.0
        ok = ok and env.sum_over_ranks(1.5) == 3.0
        # the control flow of measure_config's re-measure: only rank 1 "sees a bad clock", both ranks take the branch
        calls = 0
        if env.any_rank(rank == 1):
            dist.barrier()
            calls += 1
        dist.barrier()
        ok = ok and calls == 1
        # Env.guarded: a failure every rank hits is reported on every rank; a failure on one rank (after the same
        # collectives) turns into an error on all of them; success passes the result through
        def boom():
            raise RuntimeError("same on every rank")
        res, err = env.guarded(boom)
        ok = ok and res is None and "same on every rank" in err
        res, err = env.guarded(lambda: (_ for _ in ()).throw(ValueError("rank 1 only")) if rank == 1 else 7)
        ok = ok and res is None and err is not None and (("rank 1 only" in err) == (rank == 1))
        res, err = env.guarded(lambda: 40 + rank)
        ok = ok and err is None and res == 40 + rank
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_bench_rank_agreement_helpers_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_agree_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def test_bench_has_no_rank_local_branch_around_a_collective():
    """Static check of bench.py: inside measure_config / measure_e2e / main, a call to env.barrier() or a cross-rank
    reduction must not sit under an `if` / `try` whose outcome can differ between ranks.  Allowed guards: conditions built
    only from run-wide facts (world size, args, config names, `env.any_rank(..)` results)."""
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    collective = {"barrier", "max_over_ranks", "min_over_ranks", "sum_over_ranks", "all_ranks", "any_rank"}
    run_wide = ("env.world", "world", "args.", "env.any_rank", "want_e2e", "div", "name in", "cfg[", "c[", "h2d", "d2h", "k > 0",
                "ok_all", "self.world")

    def has_collective(node):
        for n in ast.walk(node):
            if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr in collective:
                return True
            if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id in ("timed", "copies", "measure_config"):
                return True
        return False

    bad = []
    for fn in ast.walk(tree):
        if not isinstance(fn, ast.FunctionDef) or fn.name not in ("measure_config", "measure_e2e", "main"):
            continue
        for node in ast.walk(fn):
            if isinstance(node, ast.If) and has_collective(ast.Module(body=node.body + node.orelse, type_ignores=[])):
                cond = ast.unparse(node.test)
                if not any(tok in cond for tok in run_wide):
                    bad.append((fn.name, node.lineno, cond))
            if isinstance(node, ast.Try) and has_collective(ast.Module(body=node.body, type_ignores=[])):
                # a try whose body holds collectives is only safe when the handler cannot swallow a rank-local failure:
                # the rank-0-only mg leg sits between two barriers all ranks reach regardless
                src = ast.unparse(node)
                if "scir_b200_mg_create" not in src:
                    bad.append((fn.name, node.lineno, "try"))
    assert not bad, bad
