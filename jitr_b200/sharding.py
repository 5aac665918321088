"""Sample-axis sharding of a UQ sweep over ranks (one process per GPU).

Every (sample, energy, l, j) system is independent: rank r owns a contiguous block of
samples, all energies and partial waves stay local, nothing is exchanged during the solve, and the
observables are gathered to the root rank ONCE at the end with a single ``torch.distributed.gather``
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  The reference has no counterpart: its user
loop over samples is serial (examples/notebooks/kduq_uq_demo.ipynb cell 16).
"""
from __future__ import annotations


def shard_bounds(n_samples: int, world: int, rank: int):
    """[lo, hi) of rank's contiguous block; blocks differ by at most one sample and cover 0..n_samples."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    if n_samples < 0:
        raise ValueError("n_samples must be >= 0")
    return rank * n_samples // world, (rank + 1) * n_samples // world


def gather_to_root(shard, n_samples: int, dst: int = 0, group=None, out=None):
    """Gather the per-rank blocks ``shard`` (leading axis = this rank's samples, as cut by
    ``shard_bounds(n_samples, world, rank)``) into one tensor of leading size ``n_samples`` on rank
    ``dst``; returns None on the other ranks.  One collective; uneven blocks are padded to the
    largest block for the transfer and trimmed on the root.  ``out`` (root only): a preallocated tensor of leading size
    ``n_samples`` to receive into -- with equal blocks every rank's message lands directly in its slice (no staging copy:
    what a 10^5-sample sweep needs, whose angular distributions are 9 GB per rank)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return shard
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_samples, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if shard.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: shard has {shard.shape[0]} samples, expected {hi - lo}")
    send = shard.contiguous()
    if out is not None and all(h - l == mx for l, h in sizes):
        bufs = [out[l:h] for l, h in sizes] if rank == dst else None
        dist.gather(send, bufs, dst=dst, group=group)
        return out if rank == dst else None
    if hi - lo < mx:
        pad = torch.zeros((mx - (hi - lo),) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
        send = torch.cat([send, pad], dim=0)
    bufs = None
    if rank == dst:
        bufs = [torch.empty_like(send) for _ in range(world)]
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[: h - l] for b, (l, h) in zip(bufs, sizes)], dim=0)


def gather_chunk_async(chunk, out, rank_offsets, c0: int, c1: int, dst: int = 0, group=None):
    """Pipelined form of ``gather_to_root`` for a sweep processed in chunks of samples: every rank passes the rows [c0, c1) of ITS
    block (``chunk``, contiguous) as soon as they are finished, the root receives rank r's rows in ``out[rank_offsets[r] + c0 :
    rank_offsets[r] + c1]`` (``out`` is ignored elsewhere).  The collective is launched asynchronously (``async_op=True``: NCCL's own
    stream, ordered after the work already queued on the current stream) so that the next chunk's kernels run beside the transfer;
    returns the work handle -- ``handle.wait()`` orders the current stream after the transfer without blocking the host.
    Blocks must be equal (``rank_offsets[r] = r * block``)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bufs = [out[rank_offsets[r] + c0: rank_offsets[r] + c1] for r in range(world)] if rank == dst else None
    return dist.gather(chunk.contiguous(), bufs, dst=dst, group=group, async_op=True)


class PeerWindow:
    """A buffer on the root rank's GPU that every rank of the node writes with COPY-ENGINE peer copies (CUDA IPC mapping of the
    root's allocation + ``cudaMemcpyAsync`` over NVLink): no SM takes part, so a chunk of observables travels to the root while
    the persistent sweep kernel -- which fills every SM -- solves the next chunk.  (An NCCL gather needs SMs for its copy
    kernels on both ends; launched behind a chunk it waits for the next chunk's CTAs to drain and holds SMs meanwhile: measured
    SLOWER than the one-shot gather)  Single node only; every rank must see the root's device.

        win = PeerWindow((n_samples, E, n_angles), torch.float64, device)      # collective: all ranks
        win.push(chunk, row0)        # after the kernels queued on the current stream, on a side stream
        win.flush()                  # current stream waits for this rank's copies; then one tiny all_reduce = "all have landed"
        win.buf                      # the root's tensor (on the other ranks: the IPC-mapped view of it)
    """

    def __init__(self, shape, dtype, device, dst: int = 0, group=None):
        import torch
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor

        self.group, self.dst = group, dst
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        # every step that can fail on ONE rank is followed by a vote before the next collective, so that a failure raises on all
        # ranks (the caller falls back to the gather) instead of leaving the others waiting
        err = None
        payload = [None]
        if self.rank == dst:
            try:
                self.buf = torch.empty(shape, dtype=dtype, device=self.device)
                payload = [reduce_tensor(self.buf)]
            except Exception as exc:  # noqa: BLE001
                err = exc
        dist.broadcast_object_list(payload, src=dst, group=group)
        if self.rank != dst:
            if payload[0] is None:
                err = RuntimeError("the root could not export its buffer")
            else:
                try:
                    fn, args = payload[0]
                    # The IPC handle is opened in THIS rank's device context (argument 6 of rebuild_cuda_tensor is the device the
                    # handle is opened on): the mapping is then a peer pointer this GPU's copy engines write through NVLink directly.
                    # Opened on the root's device instead (the default), the memory belongs to another context of this process and
                    # the copy is staged through the host: measured 37 GB/s and no overlap (tools/p2p_probe.py).
                    args = list(args)
                    if not isinstance(args[6], int):
                        raise RuntimeError("unexpected layout of torch's CUDA tensor reduction")
                    args[6] = self.device.index
                    self.buf = fn(*args)
                except Exception as exc:  # noqa: BLE001
                    err = exc
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) != 1:
            self.buf = None
            raise RuntimeError(f"PeerWindow unavailable on this node: {err!r}")
        self.stream = torch.cuda.Stream(device=self.device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._self_check(dist.get_world_size(group))

    def _self_check(self, world):
        """Every rank writes a marker through its mapping, the root reads them back: a mapping that does not reach the root's
        memory (an IPC handle rebuilt on a different layout of torch's reduction tuple, say) raises here on EVERY rank, and the
        caller falls back to the collective."""
        import torch
        import torch.distributed as dist

        flat = self.buf.view(-1)
        ok = torch.ones(1, dtype=torch.int32, device=self.device)
        if flat.numel() >= world:
            saved = flat[:world].clone() if self.rank == self.dst else None
            dist.barrier(group=self.group)
            flat[self.rank: self.rank + 1].copy_(torch.full((1,), float(self.rank + 1), dtype=flat.dtype, device=self.device))
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
            if self.rank == self.dst:
                want = torch.arange(1, world + 1, dtype=flat.dtype, device=self.device)
                ok[0] = int(torch.equal(flat[:world], want))
                flat[:world].copy_(saved)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) != 1:
            raise RuntimeError("PeerWindow: the peers' writes did not reach the root's buffer")

    def push(self, chunk, row0: int):
        import torch

        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.buf[row0: row0 + chunk.shape[0]].copy_(chunk, non_blocking=True)

    def flush(self):
        import torch
        import torch.distributed as dist

        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        dist.all_reduce(self._flag, group=self.group)  # ordered after the copies on every rank: when it completes, all have landed

    def close(self):
        """Collective: the mapped views are dropped before the root releases its allocation."""
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize(self.device)
        if self.rank != self.dst:
            self.buf = None
            torch.cuda.ipc_collect()  # give the mapping back now, not at the next allocator sweep
        dist.barrier(group=self.group)
        self.buf = None
        if self.rank == self.dst:
            torch.cuda.ipc_collect()
