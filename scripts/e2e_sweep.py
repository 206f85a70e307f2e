#!/usr/bin/env python
"""Sweep of the host pipeline's knobs on one GPU (not the driver's bench): packer threads, chunk size, packed-only /
raw-only / hybrid feed, output mode.  Prints one line per setting: ms per call, G k-mers/s, chunks sent raw.

  python scripts/e2e_sweep.py [--reads 10000000]
"""
from __future__ import annotations

import argparse
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import kmers_b200 as kb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    args = ap.parse_args()
    n, L, K = args.reads, 150, 31
    W = L - K + 1
    torch.cuda.set_device(0)
    cores = len(os.sched_getaffinity(0))
    print(f"host: {cores} usable cpus, pack isa {kb.host_pack_isa()}", flush=True)

    # ---- host packer alone: threads x GB/s (python threads; ctypes releases the GIL)
    ctx = kb.Context(0)
    batch = ctx.generate(42, n, L)
    host = torch.empty(n * L, dtype=torch.uint8, pin_memory=True)
    hnp = host.numpy()
    hnp[:] = batch.download()
    nw = (n * L + 15) // 16
    bits = torch.empty(nw, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    inv = torch.empty(nw, dtype=torch.int16, pin_memory=True).numpy().view(np.uint16)
    for T in (1, 2, 4, 8, 16, 32):
        if T > 2 * cores:
            break
        def work(i):
            per = (n * L // T) // 64 * 64
            s = i * per
            e = n * L if i == T - 1 else s + per
            kb.host_pack(hnp[s:e], bits[s // 16:], inv[s // 16:])
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(i,)) for i in range(T)]
            [t.start() for t in th]
            [t.join() for t in th]
            best = min(best, time.perf_counter() - t0)
        print(f"host_pack threads={T:2d}: {n * L / best / 1e9:7.1f} GB/s of ASCII in", flush=True)
    for T in (1, 4, 8, 16):
        print(f"host_read_probe threads={T:2d}: {n * L / min(kb.host_read_probe(hnp, T) for _ in range(3)) / 1e9:7.1f} GB/s", flush=True)
    # memcpy reference
    dst = np.empty_like(hnp)
    t0 = time.perf_counter(); dst[:] = hnp; t1 = time.perf_counter() - t0
    print(f"numpy memcpy 1 thread: {n * L / t1 / 1e9:.1f} GB/s", flush=True)
    del dst

    canon = torch.empty(n * W, dtype=torch.int64, device="cuda")
    hsh = torch.empty(n * W, dtype=torch.int64, device="cuda")
    dev = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    raw_ms = 1e3 * (time.perf_counter() - t0) / 3
    print(f"raw H2D of the ASCII reads: {raw_ms:.2f} ms ({n * L / raw_ms / 1e6:.1f} GB/s)", flush=True)
    del dev

    def run(label, env, threads, outs="dev", src=hnp, reps=5):
        for k_, v in env.items():
            os.environ[k_] = v
        ctx.set_host_threads(threads)
        kw = {"dev": dict(out_canon=canon, out_hash=hsh), "none": {}}[outs]
        for _ in range(2):
            d = ctx.extract_canonical_host(src, n, L, K, **kw)
        t0 = time.perf_counter()
        for _ in range(reps):
            d = ctx.extract_canonical_host(src, n, L, K, **kw)
        ms = 1e3 * (time.perf_counter() - t0) / reps
        st = ctx.host_stats()
        print(f"{label:58s} thr={threads:2d} {ms:8.2f} ms {n * W / ms / 1e6:8.1f} G kmers/s  chunks {st['chunks']:4d} raw {st['raw_chunks']:4d} "
              f"h2d {st['h2d_bytes'] / 1e6:8.1f} MB  n_valid {d[0]}", flush=True)
        for k_ in env:
            os.environ.pop(k_, None)

    tset = sorted({max(1, cores // 4), max(1, cores // 2), max(1, cores - 1), cores})
    for t in tset:
        run("hybrid (default) chunk 8 MB, device arrays", {}, t)
    for mb in ("4", "16", "32"):
        run(f"hybrid chunk {mb} MB, device arrays", {"KMB_PIPE_CHUNK_MB": mb}, cores)
    for t in tset:
        run("packed only, device arrays", {"KMB_PIPE_RAW": "0"}, t)
    run("raw ASCII only, device arrays", {"KMB_PIPE_PACK": "0"}, cores)
    run("hybrid, digest only", {}, cores, outs="none")
    pageable = np.array(hnp)  # pageable copy
    for t in tset[-2:]:
        run("pageable input (packed only by construction)", {}, t, src=pageable)
    ctx.close()


if __name__ == "__main__":
    main()
