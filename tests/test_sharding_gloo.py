"""World-size-2 (and 3, uneven) gloo tests of the sample-axis sharding used by the multi-GPU sweep."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jitr_b200.sharding import gather_to_root, shard_bounds  # noqa: E402


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 416, 100000):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_samples, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_samples, world, rank)
        # the "observables" of sample s at (energy e, angle t): a known function of the GLOBAL sample index
        s = torch.arange(lo, hi, dtype=torch.float64)[:, None, None]
        e = torch.arange(3, dtype=torch.float64)[None, :, None]
        t = torch.arange(5, dtype=torch.float64)[None, None, :]
        shard = 1000.0 * s + 10.0 * e + t
        full = gather_to_root(shard, n_samples, dst=0)
        # per-rank work counters: summed like bench.py does (all_reduce SUM) and max-over-ranks timing
        tt = torch.tensor([float(hi - lo), float(rank + 1)], dtype=torch.float64)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        if rank == 0:
            S = torch.arange(n_samples, dtype=torch.float64)[:, None, None]
            assert full.shape == (n_samples, 3, 5)
            assert torch.equal(full, 1000.0 * S + 10.0 * e + t)
            assert sm[0].item() == n_samples and mx[1].item() == world
            out.put("ok")
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_samples", [(2, 416), (3, 10)])
def test_gather_to_root_gloo(world, n_samples):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_samples, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get(timeout=5) == "ok"


def test_gather_single_process_is_identity():
    x = torch.arange(12.0).reshape(4, 3)
    assert gather_to_root(x, 4) is x


def _worker_chunks(rank, world, port, block, nchunk, out):
    """bench.py's N > 1 step: the block is processed in chunks, every chunk gathered asynchronously into the root's preallocated
    slices (gather_chunk_async) and into `out=` (gather_to_root) -- both must reproduce the global array."""
    import torch.distributed as dist

    from jitr_b200.sharding import gather_chunk_async

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = block * world
        s = torch.arange(rank * block, (rank + 1) * block, dtype=torch.float64)[:, None]
        shard = 7.0 * s + torch.arange(4, dtype=torch.float64)[None, :]
        full = torch.full((n, 4), -1.0, dtype=torch.float64) if rank == 0 else None
        works = []
        for c in range(nchunk):
            c0, c1 = c * block // nchunk, (c + 1) * block // nchunk
            works.append(gather_chunk_async(shard[c0:c1], full, [r * block for r in range(world)], c0, c1))
        for w in works:
            w.wait()
        pre = torch.empty((n, 4), dtype=torch.float64) if rank == 0 else None
        got = gather_to_root(shard, n, dst=0, out=pre)
        if rank == 0:
            want = 7.0 * torch.arange(n, dtype=torch.float64)[:, None] + torch.arange(4, dtype=torch.float64)[None, :]
            assert torch.equal(full, want)
            assert got is pre and torch.equal(pre, want)
            out.put("ok")
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,block,nchunk", [(2, 12, 4), (2, 5, 3), (3, 4, 1)])
def test_chunked_async_gather_gloo(world, block, nchunk):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_chunks, args=(r, world, port, block, nchunk, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get(timeout=5) == "ok"
