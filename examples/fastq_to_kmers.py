#!/usr/bin/env python
"""FASTQ text -> canonical k-mers on the GPU, the way a user of COMBINE-lab/kmers would write

    for read in fastq:  it = CanonicalKmerIterator::from_u8_slice(read, k);  while !it.exhausted() { use(it.get()); it.inc(); }

here as three batched calls.  Usage:  python examples/fastq_to_kmers.py [reads.fastq] [k]
Without a file a synthetic FASTQ (200 000 reads of 100..150 bases, 0.2 % N) is generated in memory."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import kmers_b200 as kb


def synthetic_fastq(n_reads=200_000, seed=1):
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b"ACGTN", dtype=np.uint8)
    out = []
    for i, n in enumerate(rng.integers(100, 151, size=n_reads)):
        seq = letters[rng.choice(5, size=int(n), p=[.2495, .2495, .2495, .2495, .002])].tobytes()
        out.append(b"@read%d\n%s\n+\n%s\n" % (i, seq, b"I" * int(n)))
    return b"".join(out)


def main():
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
    text = open(sys.argv[1], "rb").read() if len(sys.argv) > 1 else synthetic_fastq()
    import torch
    torch.zeros(1, device="cuda")                            # CUDA context + allocator warm-up stay out of the timings
    with kb.Context(0) as ctx:
        ctx.ingest_fastx(text[: text.index(b"\n@", len(text) // 50) + 1] if len(text) > 10_000 else text).extract_compact(k, to="device")
        t0 = time.perf_counter()
        batch = ctx.ingest_fastx(text)                       # parse on the host (multi-threaded), one pinned upload
        t1 = time.perf_counter()
        out = batch.extract_compact(k, to="device")          # exactly the iterator's (pos, canonical word, LexHash) sequence
        ctx.sync()
        t2 = time.perf_counter()
        mm, mpos = batch.minimizers(k, min(k, 15), to="device")   # (k, w) minimizers, one per k-mer window
        ctx.sync()
        t3 = time.perf_counter()
        offs = out["emit_offsets"].cpu().numpy()
        print(f"{batch.n_reads} reads, {batch.n_bytes} bases, k={k}: {out['n']} canonical k-mers")
        print(f"  parse + upload   {1e3 * (t1 - t0):8.2f} ms  ({len(text) / (t1 - t0) / 1e9:.2f} GB/s of text)")
        print(f"  compact extract  {1e3 * (t2 - t1):8.2f} ms  ({out['n'] / (t2 - t1) / 1e9:.1f} G k-mers/s)")
        print(f"  minimizers       {1e3 * (t3 - t2):8.2f} ms")
        r = 0
        a, b = int(offs[r]), int(offs[r + 1])
        print(f"  read 0 emits {b - a} k-mers; first: pos={int(out['pos'][a])} canon={int(out['canon'][a]) & (2**64 - 1):#x}")


if __name__ == "__main__":
    main()
