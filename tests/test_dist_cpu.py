"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: shard arithmetic and the one collective of
the path (histogram + digest all-reduce).  The per-shard numbers come from the CPU oracle here (test
infrastructure); on GPUs the same `kmers_b200.dist` functions run over NCCL (tests/test_gpu_dist.py, bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as ko
from kmers_b200 import dist as kd

K, BITS, G, SEED, THR = 31, 10, 120_000, 44, 105


def test_shard_reads_partition():
    for n in (0, 1, 7, 1000, 10_000_001):
        for world in (1, 2, 3, 8):
            spans = [kd.shard_reads(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_sequence_halo_covers_every_window_once():
    for n, k, world in ((1000, 31, 4), (100, 31, 8), (64, 5, 3), (31, 31, 2), (10, 31, 2)):
        starts = []
        for r in range(world):
            s, e, load = kd.shard_sequence(n, k, r, world)
            assert load == min(n, e + k - 1)
            starts += [g for g in range(s, e) if g + k <= load]
        assert starts == list(range(max(0, n - k + 1)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop, load_stop = kd.shard_sequence(G, K, rank, world)
        seq = ko.generate_bases(SEED, start, load_stop - start, THR)
        r = ko.extract_canonical(seq, K, n_reads=1, fixed_len=seq.size, hist_bits=BITS, materialize=False)
        hist = torch.from_numpy(r["hist"].view(np.int64).copy())
        g_hist, g_dig = kd.allreduce_histogram(hist, (r["n_valid"], r["checksum_canon"], r["checksum_hash"]))
        if rank == 0:
            out.put((g_hist.numpy().view(np.uint64).copy(), g_dig))
    finally:
        dist.destroy_process_group()


def test_histogram_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    g_hist, g_dig = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = ko.extract_canonical(ko.generate_bases(SEED, 0, G, THR), K, n_reads=1, fixed_len=G, hist_bits=BITS,
                                 materialize=False)
    assert np.array_equal(g_hist, whole["hist"])
    assert g_dig == (whole["n_valid"], whole["checksum_canon"], whole["checksum_hash"])


def test_allreduce_single_process_is_identity():
    hist = torch.arange(8, dtype=torch.int64)
    h, d = kd.allreduce_histogram(hist, (5, 2**64 - 1, 2**63 + 7))
    assert torch.equal(h, hist) and d == (5, 2**64 - 1, 2**63 + 7)
