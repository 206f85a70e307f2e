"""Config 5 on GPUs: one synthetic sequence sharded with a K-1 halo, fused histogram + digest per shard, one
all-reduce.  Single-GPU form here (virtual ranks, summed by hand); with >= 2 GPUs the same code runs under
torchrun over NCCL."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_histogram_virtual_ranks():
    import kmers_b200 as kb
    import oracle as ko
    from kmers_b200 import dist as kd
    G, k, bits, seed, thr = 3_000_000, 31, 16, 44, 105
    whole = ko.extract_canonical(ko.generate_bases(seed, 0, G, thr), k, n_reads=1, fixed_len=G, hist_bits=bits,
                                 materialize=False, n_threads=1)
    with kb.Context(0) as ctx:
        for world in (1, 3, 8):
            total = np.zeros(1 << bits, dtype=np.uint64)
            dig = [0, 0, 0]
            nwin = 0
            for r in range(world):
                h, d, w = kd.sharded_histogram(ctx, seed, G, k, bits, r, world, n_thresh20=thr)
                total += h.cpu().numpy().view(np.uint64)
                dig = [(a + b) % 2**64 for a, b in zip(dig, d)]
                nwin += w
            assert nwin == G - k + 1
            assert np.array_equal(total, whole["hist"])
            assert tuple(dig) == (whole["n_valid"], whole["checksum_canon"], whole["checksum_hash"])


_TORCHRUN_BODY = r'''
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import kmers_b200 as kb
from kmers_b200 import dist as kd
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
with kb.Context(lr) as ctx:
    h, d, w = kd.sharded_histogram(ctx, 44, 3_000_000, 31, 16, rank, world, n_thresh20=105)
if rank == 0:
    np.save({out!r}, h.cpu().numpy().view(np.uint64))
    json.dump(list(d), open({out!r} + ".json", "w"))
dist.destroy_process_group()
'''


def test_sharded_histogram_nccl_multi_gpu(tmp_path):
    import json
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    import oracle as ko
    out = str(tmp_path / "hist.npy")
    script = tmp_path / "run.py"
    script.write_text(_TORCHRUN_BODY.format(root=ROOT, out=out))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 2)}",
                           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], timeout=600)
    whole = ko.extract_canonical(ko.generate_bases(44, 0, 3_000_000, 105), 31, n_reads=1, fixed_len=3_000_000,
                                 hist_bits=16, materialize=False)
    assert np.array_equal(np.load(out), whole["hist"])
    assert tuple(json.load(open(out + ".json"))) == (whole["n_valid"], whole["checksum_canon"], whole["checksum_hash"])


def test_single_process_allreduce_through_the_c_abi():
    """kmb_allreduce_u64: one process, one context per GPU (the Rust host's shape).  With one GPU the call is a
    synchronising no-op; with >= 2 it must equal the sum of the per-GPU [histogram | digest] vectors."""
    import torch
    import kmers_b200 as kb
    import oracle as ko
    from kmers_b200 import dist as kd
    n = min(torch.cuda.device_count(), 4)
    G, k, bits, seed, thr = 2_000_000, 31, 12, 44, 105
    whole = ko.extract_canonical(ko.generate_bases(seed, 0, G, thr), k, n_reads=1, fixed_len=G, hist_bits=bits,
                                 materialize=False)
    ctxs = [kb.Context(d) for d in range(n)]
    try:
        bufs = []
        for r, ctx in enumerate(ctxs):
            start, stop, load_stop = kd.shard_sequence(G, k, r, n)
            with torch.cuda.device(r):
                batch = ctx.generate(seed, 1, load_stop - start, n_thresh20=thr, first_index=start)
                hist, dig = batch.histogram(k, bits, to="device")
                tail = torch.tensor([kd._to_i64(int(d)) for d in dig], dtype=torch.int64, device=f"cuda:{r}")
                bufs.append(torch.cat([hist.reshape(-1).to(torch.int64), tail]).contiguous())
        kd.allreduce_single_process(ctxs, bufs)
        for b in bufs:   # every GPU holds the global result
            got = b.cpu().numpy().view(np.uint64)
            assert np.array_equal(got[:-3], whole["hist"])
            assert tuple(int(v) for v in got[-3:]) == (whole["n_valid"], whole["checksum_canon"], whole["checksum_hash"])
    finally:
        for c in ctxs:
            c.close()
