#!/usr/bin/env python
"""Does the distance between the two output arrays matter for the bench kernel?  (DRAM channel imbalance.)
canon at offset 0 of one big allocation, hash at n_slots * 8 + delta: kernel time per delta."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import kmers_b200 as kb

n, L, K = 10_000_000, 150, 31
W = L - K + 1
ns = n * W
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = kb.Context(0, stream=stream.cuda_stream)
batch = ctx.generate(42, n, L)
big = torch.empty(2 * ns + (1 << 26), dtype=torch.int64, device="cuda")
for delta_bytes in (0, 32, 256, 1024, 4096, 8192, 64 << 10, 1 << 20, (1 << 20) + 4096, 32 << 20, (128 << 20) + 2048):
    d = delta_bytes // 8
    out = kb.CanonicalKmers(k=K, n_slots=ns, canon=big[:ns], hash=big[ns + d: 2 * ns + d])
    for _ in range(3):
        batch.extract_canonical(K, out=out)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(15)]
    torch.cuda.synchronize()
    for a, b in evs:
        a.record(stream)
        batch.extract_canonical(K, out=out)
        b.record(stream)
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    print(f"hash base = canon end + {delta_bytes:>10d} B: median {ts[len(ts) // 2]:.4f} ms  best {ts[0]:.4f} ms  ({20.7 / ts[len(ts) // 2]:.2f} TB/s)", flush=True)
# separate allocations (what bench.py does)
c, h = torch.empty(ns, dtype=torch.int64, device="cuda"), torch.empty(ns, dtype=torch.int64, device="cuda")
out = kb.CanonicalKmers(k=K, n_slots=ns, canon=c, hash=h)
for _ in range(3):
    batch.extract_canonical(K, out=out)
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(15)]
torch.cuda.synchronize()
for a, b in evs:
    a.record(stream)
    batch.extract_canonical(K, out=out)
    b.record(stream)
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in evs)
print(f"separate torch allocations (canon {c.data_ptr():#x}, hash {h.data_ptr():#x}): median {ts[len(ts) // 2]:.4f} ms best {ts[0]:.4f} ms", flush=True)
