#!/usr/bin/env python
"""Device-resident timing of every kernel family on the BASELINE.json configs (one GPU).

Not the driver's bench (that is bench.py, configs[1] only): this fills the per-config table of BASELINE.md /
DESIGN.md.  Each row: CUDA-event time of the launch(es) on the stream they run on, algorithmic bytes
(SURVEY.md 8d), achieved GB/s and the fraction of the measured copy peak.

  python scripts/bench_configs.py [--out profiles/configs_r01.json] [--quick]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import kmers_b200 as kb


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def time_ms(stream, fn, reps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    scale = 10 if args.quick else 1
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = kb.Context(0, stream=stream.cuda_stream)
    pk = peak()
    rows = []

    def row(name, units, unit_name, alg_bytes, avg, best, note=""):
        r = {"config": name, "units": units, "unit": unit_name, "algorithmic_bytes": alg_bytes, "avg_ms": avg, "best_ms": best,
             "units_per_s": units / (avg / 1e3), "achieved_GBps": alg_bytes / (avg / 1e3) / 1e9,
             "frac_of_measured_copy_peak": alg_bytes / (avg / 1e3) / 1e9 / pk, "note": note}
        rows.append(r)
        print(f"{name:58s} {avg:9.4f} ms  {r['units_per_s'] / 1e9:8.2f} G {unit_name}/s  {r['achieved_GBps']:8.1f} GB/s  "
              f"{100 * r['frac_of_measured_copy_peak']:6.1f} %", flush=True)

    i64 = lambda n: torch.empty(n, dtype=torch.int64, device="cuda")

    # ---------------- configs 1/2: K=31, 150 bp
    n, L, K = 10_000_000 // scale, 150, 31
    W = L - K + 1
    batch = ctx.generate(42, n, L)
    out = kb.CanonicalKmers(k=K, n_slots=n * W, canon=i64(n * W), hash=i64(n * W))
    avg, best = time_ms(stream, lambda: batch.extract_canonical(K, out=out))
    row("config2 K=31 150bp canon+hash", n * W, "kmers", n * (L + W * 16), avg, best)
    avg, best = time_ms(stream, lambda: batch.extract_canonical(K, out=out, digest=True))
    row("config2 + digest (sync D2H of 24 B inside)", n * W, "kmers", n * (L + W * 16), avg, best)
    out1 = kb.CanonicalKmers(k=K, n_slots=n * W, canon=out.canon, hash=None)
    avg, best = time_ms(stream, lambda: batch.extract_canonical(K, out=out1))
    row("config2 canon only", n * W, "kmers", n * (L + W * 8), avg, best)
    outf = kb.CanonicalKmers(k=K, n_slots=n * W, canon=out.canon, hash=out.hash, fw=i64(n * W), rc=i64(n * W))
    avg, best = time_ms(stream, lambda: batch.extract_canonical(K, out=outf))
    row("config2 canon+hash+fw+rc", n * W, "kmers", n * (L + W * 32), avg, best)
    del outf
    avg, best = time_ms(stream, lambda: batch.extract_canonical(K, out=out, validate=False))
    row("config2 no-validate (Path-E semantics)", n * W, "kmers", n * (L + W * 16), avg, best)
    for k in (15, 21):
        wk = L - k + 1
        ok = kb.CanonicalKmers(k=k, n_slots=n * wk, canon=i64(n * wk), hash=i64(n * wk))
        avg, best = time_ms(stream, lambda: batch.extract_canonical(k, out=ok))
        row(f"150bp K={k} canon+hash", n * wk, "kmers", n * (L + wk * 16), avg, best)
        del ok
    hist, _ = batch.histogram(K, 16, to="device")
    avg, best = time_ms(stream, lambda: batch.histogram(K, 16, hist=hist, accumulate=False, digest=False))
    row("config2 reads, fused histogram 2^16 bins (nothing materialised)", n * W, "kmers", n * L + 8 * 65536, avg, best,
        "ALU/atomic-bound by construction; % of HBM peak is low by design")

    # ragged path on the same bytes
    offs = (torch.arange(n + 1, dtype=torch.int64, device="cuda") * L)
    bases_dev = torch.from_numpy(batch.download()).cuda()
    rb = ctx.attach(bases_dev, dev_offsets=offs)
    avg, best = time_ms(stream, lambda: rb.extract_canonical(K, out=out))
    row("config2 bytes through the CSR (ragged) path", n * W, "kmers", n * (L + W * 16 + 16), avg, best)

    # ---------------- config 3: K=63 two words
    batch = ctx.attach(bases_dev, fixed_len=L)
    w63 = L - 63 + 1
    wide = kb.CanonicalKmers(k=63, n_slots=n * w63, canon=i64(2 * n * w63), hash=None, words_per_kmer=2)

    def run_wide(want_hash):
        d = kb.Digest()
        import ctypes as C
        from kmers_b200.context import _ptr
        ctx._ck(ctx._lib.kmb_extract_canonical_wide(ctx._h, 63, kb.ENC_ACGT, 0, _ptr(wide.canon), _ptr(wide.hash) if want_hash else None, None))

    avg, best = time_ms(stream, lambda: run_wide(False))
    row("config3 K=63 2xu64 canon only", n * w63, "kmers", n * (L + w63 * 16), avg, best, "extension: parity unpinned above K=32")
    wide.hash = i64(2 * n * w63)
    avg, best = time_ms(stream, lambda: run_wide(True))
    row("config3 K=63 2xu64 canon+hash", n * w63, "kmers", n * (L + w63 * 32), avg, best, "extension")
    del wide

    # ---------------- pack / unpack / revcomp / word ops
    for wb in (64, 8):
        img, _ = batch.pack(kb.ENC_ACGT, wb, to="device")
        from kmers_b200.context import _ptr
        avg, best = time_ms(stream, lambda: ctx._ck(ctx._lib.kmb_pack(ctx._h, kb.ENC_ACGT, wb, _ptr(img), None)))
        row(f"pack 150bp reads -> u{wb} words (Encoding::encode bulk)", n * L, "bases", n * L + img.numel(), avg, best)
    nk = 100_000_000 // scale
    words = torch.randint(0, 2**62, (nk,), dtype=torch.int64, device="cuda")
    dst = torch.empty_like(words)
    from kmers_b200.context import _ptr
    avg, best = time_ms(stream, lambda: ctx._ck(ctx._lib.kmb_reverse_complement_words(ctx._h, 31, _ptr(words), _ptr(dst), nk)))
    row("reverse_complement_words K=31 (u64)", nk, "kmers", nk * 16, avg, best)
    avg, best = time_ms(stream, lambda: ctx._ck(ctx._lib.kmb_lexhash_words(ctx._h, 31, _ptr(words), _ptr(dst), nk)))
    row("lexhash_words K=31", nk, "kmers", nk * 16, avg, best)
    avg, best = time_ms(stream, lambda: ctx._ck(ctx._lib.kmb_revcomp_words(ctx._h, kb.ENC_ACGT, 63, 64, 2, _ptr(words), _ptr(dst), nk // 2)))
    row("Encoding::rev_comp::<63> on [u64;2]", nk // 2, "kmers", nk * 16, avg, best)
    del words, dst, out, out1, bases_dev, offs

    # ---------------- config 4: long reads with N
    n4, L4 = 100_000 // scale, 10_000
    w4 = L4 - K + 1
    batch = ctx.generate(43, n4, L4, n_thresh20=1049)
    out = kb.CanonicalKmers(k=K, n_slots=n4 * w4, canon=i64(n4 * w4), hash=i64(n4 * w4))
    avg, best = time_ms(stream, lambda: batch.extract_canonical(K, out=out))
    row("config4 K=31 10kbp reads, 0.1% N", n4 * w4, "slots", n4 * (L4 + w4 * 16), avg, best)
    del out

    # ---------------- config 5: 1 Gbp single sequence, fused histogram
    g = 1_000_000_000 // scale
    batch = ctx.generate(44, 1, g, n_thresh20=105)
    hist, dig = batch.histogram(K, 16, to="device")
    avg, best = time_ms(stream, lambda: batch.histogram(K, 16, hist=hist, accumulate=False, digest=True))
    row("config5 1 Gbp, fused histogram 2^16 bins + digest", g - K + 1, "kmers", g + 8 * 65536, avg, best,
        "ALU/atomic-bound by construction")

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"hbm_peak_gbs": pk, "rows": rows}, open(args.out, "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
