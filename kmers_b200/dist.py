"""Multi-GPU plumbing: how a batch / a long sequence shards across ranks, and the
one optional collective of the path (final histogram + digest reduction).

The extraction itself needs no communication: every window is a pure function
of its own K bytes and reads are independent.  One process per
GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

MASK64 = (1 << 64) - 1


def shard_reads(n_reads: int, rank: int, world: int) -> Tuple[int, int]:
    """Reads [start, stop) owned by `rank`: contiguous, balanced, covering every read exactly once."""
    return n_reads * rank // world, n_reads * (rank + 1) // world


def shard_sequence(n_bases: int, k: int, rank: int, world: int) -> Tuple[int, int, int]:
    """One long sequence cut into contiguous ranges with a K-1 halo on the right.

    Returns (start, stop, load_stop): the rank owns the windows whose FIRST base lies in [start, stop) and
    must load bases [start, load_stop) with load_stop = min(n_bases, stop + k - 1).  A window belongs to
    exactly one rank; windows crossing `stop` are produced by the rank that holds their first base."""
    start, stop = n_bases * rank // world, n_bases * (rank + 1) // world
    return start, stop, min(n_bases, stop + k - 1)


def _to_i64(x: int) -> int:
    x &= MASK64
    return x - (1 << 64) if x >= (1 << 63) else x


def allreduce_histogram(hist, digest: Sequence[int], group=None):
    """Sum [bins || n_valid || checksum_canon || checksum_hash] over all ranks in ONE collective.

    `hist` is an int64 torch tensor (CUDA with NCCL, CPU with gloo) holding u64 bin counts; sums wrap mod
    2^64 exactly like the reference's `.sum()` in release builds (benches/simple_benchmark.rs:21).
    Returns (global_hist, (n_valid, checksum_canon, checksum_hash)) on every rank."""
    import torch
    import torch.distributed as dist

    tail = torch.tensor([_to_i64(int(d)) for d in digest], dtype=torch.int64, device=hist.device)
    buf = torch.cat([hist.reshape(-1).to(torch.int64), tail])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out_hist = buf[:-3].reshape(hist.shape)
    out_digest = tuple(int(v) & MASK64 for v in buf[-3:].tolist())
    return out_hist, out_digest


def sharded_histogram(ctx, seed: int, n_bases: int, k: int, hist_bits: int, rank: int, world: int,
                      n_thresh20: int = 0, group=None):
    """BASELINE config 5 on this rank's GPU: generate this rank's range of ONE synthetic sequence (with halo) on the
    device, fused extraction -> histogram + digest, then the single all-reduce.  Returns (global_hist, global_digest,
    local_windows)."""
    start, stop, load_stop = shard_sequence(n_bases, k, rank, world)
    n = load_stop - start
    batch = ctx.generate(seed, 1, n, n_thresh20=n_thresh20, first_index=start)
    hist, digest = batch.histogram(k, hist_bits, to="device")
    # windows whose first base is in [stop, load_stop) belong to the next rank: none exist, because the
    # loaded range ends at stop + k - 1, so the last window starts at stop - 1 (or earlier at the sequence end)
    g_hist, g_digest = allreduce_histogram(hist, digest, group=group)
    return g_hist, g_digest, max(0, n - k + 1)


def sharded_histogram_fused(ctx, batch, k: int, hist_bits: int, buf, group=None):
    """One timed step of config 5 on this rank: the fused histogram kernel accumulates [bins | n_valid | checksum_canon |
    checksum_hash] into `buf` (int64 CUDA tensor of 2^hist_bits + 3 words, zeroed by the call), then ONE in-place
    all-reduce sums it over the ranks -- nothing is read back to the host in between ("the histogram kernel
    writes straight into the send buffer, in-place all-reduce on the same stream").  Returns buf."""
    import torch.distributed as dist

    batch.histogram(k, hist_bits, hist=buf, accumulate=False, digest_in_hist=True)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def allreduce_single_process(ctxs, bufs):
    """The same reduction for ONE process that drives several GPUs (the shape a Rust host has): in-place wrapping-u64
    sum of the int64 CUDA tensors `bufs[i]` (one per context / GPU) through the C ABI's kmb_allreduce_u64 (NCCL)."""
    import ctypes as C

    from . import _native as nv
    n = len(ctxs)
    assert n == len(bufs) and n >= 1
    count = bufs[0].numel()
    assert all(b.numel() == count and b.is_cuda and b.is_contiguous() for b in bufs)
    handles = (C.c_void_p * n)(*[c._h for c in ctxs])
    ptrs = (C.c_void_p * n)(*[b.data_ptr() for b in bufs])
    nv.check(ctxs[0]._h, nv.lib().kmb_allreduce_u64(handles, n, ptrs, count))
    return bufs
