#!/usr/bin/env python
"""Probe of sharding.PeerWindow under torchrun (2+ ranks): bandwidth of a push alone, and whether it overlaps a kernel that
fills every SM.   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/p2p_probe.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jitr_b200.sharding import PeerWindow  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rows, width = 1 << 18, 1024  # 2 GiB per rank
win = PeerWindow((rows * world, width), torch.float64, dev)
src = torch.randn(rows, width, dtype=torch.float64, device=dev)
out = {"rank": rank, "buf_device": str(win.buf.device)}


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def push_only():
    win.push(src, rank * rows)
    torch.cuda.current_stream().wait_stream(win.stream)


x = torch.randn(8192, 8192, device=dev)


def busy():
    y = x
    for _ in range(12):
        y = y @ x
    return y


def busy_and_push():
    win.push(src, rank * rows)
    busy()
    torch.cuda.current_stream().wait_stream(win.stream)


dist.barrier()
t_push = timed(push_only)
dist.barrier()
t_busy = timed(busy)
dist.barrier()
t_both = timed(busy_and_push)
dist.barrier()
win.flush()
torch.cuda.synchronize()
ok = None
if rank == 0:
    ok = bool(torch.equal(win.buf[:rows], src))
out.update(push_ms=t_push, push_GBps=src.numel() * 8 / t_push / 1e6, busy_ms=t_busy, both_ms=t_both, root_rows_ok=ok,
           p2p=[torch.cuda.can_device_access_peer(local, d) for d in range(torch.cuda.device_count())])
# the peer checks what it wrote by reading the window back through the mapping
if rank != 0:
    out["readback_ok"] = bool(torch.equal(win.buf[rank * rows: (rank + 1) * rows].to(dev), src))
print(json.dumps(out), flush=True)
dist.destroy_process_group()
