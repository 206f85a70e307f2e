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
    ap.add_argument("--gpu-only", action="store_true", help="skip the CPU-port rows")
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

    # ragged reads of 100..150 bases cut from the same stream (items straddle read boundaries everywhere)
    lens = torch.from_numpy(np.random.default_rng(1).integers(100, 151, size=n)).cuda()
    voffs = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    voffs[1:] = torch.cumsum(lens, 0)
    nb, ns = int(voffs[-1].item()), int((lens - (K - 1)).sum().item())
    vb = ctx.attach(bases_dev[:nb], dev_offsets=voffs)
    vout = kb.CanonicalKmers(k=K, n_slots=ns, canon=out.canon[:ns], hash=out.hash[:ns])
    avg, best = time_ms(stream, lambda: vb.extract_canonical(K, out=vout))
    row("ragged reads of 100..150 bp, K=31 canon+hash", ns, "kmers", nb + ns * 16 + n * 16, avg, best)
    del vout, voffs, lens

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
    mmw, mmo = torch.empty_like(words), torch.empty(nk, dtype=torch.int32, device="cuda")
    avg, best = time_ms(stream, lambda: ctx._ck(ctx._lib.kmb_minimizer_words(ctx._h, 31, 15, 15, _ptr(words), nk, _ptr(mmw), _ptr(mmo))))
    row("Kmer::minimizer_word k=31 w=15 on u64 words", nk, "kmers", nk * 20, avg, best)
    del mmw, mmo
    ni = nk // 4  # decode 25 M packed 31-mers (one u64 each) back to ASCII: 8 B in, 31 B out
    txt = torch.empty(ni * 31, dtype=torch.uint8, device="cuda")
    avg, best = time_ms(stream, lambda: ctx._ck(ctx._lib.kmb_unpack(ctx._h, kb.ENC_ACGT, 64, _ptr(words), ni, 1, 31, _ptr(txt))))
    row("Encoding::decode of u64 31-mers (bulk unpack)", ni * 31, "bases", ni * (8 + 31), avg, best)
    del txt
    del words, dst, out, out1, bases_dev, offs

    # ---------------- "next" rows: compacted output, minimizers, packed store (config-2 reads)
    import ctypes as C
    batch = ctx.generate(42, n, L, n_thresh20=1049)  # 0.1 % N so that compaction has something to drop
    cnt = C.c_uint64()
    ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, K, 0, None, None, None, None, 0, C.byref(cnt)))
    m = int(cnt.value)
    cc, ch, cp = i64(m), i64(m), torch.empty(m, dtype=torch.int32, device="cuda")
    ce = i64(n + 1)
    avg, best = time_ms(stream, lambda: ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, K, 0, _ptr(cc), _ptr(ch), _ptr(cp), _ptr(ce), m, C.byref(cnt))))
    row("config2 reads + 0.1% N, compacted (pos,canon,hash) output, one launch (look-back)", m, "kmers", n * L + m * 20 + (n + 1) * 8, avg, best,
        "iterator-identical output; single pass")
    del cc, ch, cp, ce
    batch = ctx.generate(42, n, L)
    mm, mp = i64(n * W), torch.empty(n * W, dtype=torch.int32, device="cuda")
    avg, best = time_ms(stream, lambda: ctx._ck(ctx._lib.kmb_minimizers(ctx._h, 31, 15, 15, 0, _ptr(mm), _ptr(mp))))
    row("config2 reads, minimizers (k=31, w=15): lmer word + pos per k-mer", n * W, "kmers", n * (L + W * 12), avg, best, "next row N1")
    del mm, mp
    out = kb.CanonicalKmers(k=K, n_slots=n * W, canon=i64(n * W), hash=i64(n * W))
    pb = batch.to_packed()
    avg, best = time_ms(stream, lambda: pb.extract_canonical(K, out=out))
    row("config2 reads from the 2-bit packed store (0.27 B/base in)", n * W, "kmers", n * (40 + W * 16), avg, best, "next row N2")
    del out

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

    if args.gpu_only:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump({"hbm_peak_gbs": pk, "rows": rows}, open(args.out, "w"), indent=1)
        ctx.close()
        return

    # ---------------- config 1: the reference's own bench sizes (benches/simple_benchmark.rs:58-102), GPU vs CPU port
    import time
    import oracle as ko
    cpu_rows = []
    cores = os.cpu_count() or 1
    for logn in (8, 12, 15, 24):
        nb = 1 << logn
        host = ko.generate_bases(1, 0, nb)
        t0 = time.perf_counter(); reps = max(1, (1 << 22) // nb)
        for _ in range(reps):
            r = ko.bench_windows(host, K, n_reads=1, fixed_len=nb, materialize=False)
        cpu_faithful = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            r2 = ko.extract_canonical(host, K, n_reads=1, fixed_len=nb, materialize=False)
        cpu_iter = (time.perf_counter() - t0) / reps
        b = ctx.upload(host, fixed_len=nb)
        o = kb.CanonicalKmers(k=K, n_slots=nb - K + 1, canon=i64(nb - K + 1), hash=i64(nb - K + 1))
        avg, best = time_ms(stream, lambda: b.extract_canonical(K, out=o), reps=20)
        d = b.extract_canonical(K, out=o, digest=True).digest
        assert d == (r2["n_valid"], r2["checksum_canon"], r2["checksum_hash"]) and r["checksum_canon"] == r2["checksum_canon"]
        cpu_rows.append({"config": f"config1 2^{logn} random ACGT bytes, K=31", "kmers": nb - K + 1,
                         "cpu_bench_faithful_1thread_ms": cpu_faithful * 1e3, "cpu_iterator_1thread_ms": cpu_iter * 1e3,
                         "gpu_kernel_ms": avg, "note": "GPU time is launch-latency-bound below ~2^20 bytes"})
        print(f"config1 2^{logn:<2d} B: cpu bench-faithful {cpu_faithful * 1e3:9.4f} ms, cpu iterator {cpu_iter * 1e3:9.4f} ms, gpu {avg:7.4f} ms", flush=True)
    # CPU port on bounded samples of configs 2-5 (all cores), same synthetic bytes
    def cpu_time(fn, reps=2):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps
    ns = 2_000_000 // scale
    hb = ko.generate_bases(42, 0, ns * 150)
    c_out, h_out = np.zeros(ns * 120, dtype=np.uint64), np.zeros(ns * 120, dtype=np.uint64)
    t = cpu_time(lambda: ko.extract_canonical(hb, 31, n_reads=ns, fixed_len=150, n_threads=cores, canon_out=c_out, hash_out=h_out))
    cpu_rows.append({"config": "config2 sample (2M reads) iterator + canonical + LexHash, materialised", "cores": cores, "kmers_per_s": ns * 120 / t})
    t = cpu_time(lambda: ko.bench_windows(hb, 31, n_reads=ns, fixed_len=150, n_threads=cores, materialize=False))
    cpu_rows.append({"config": "config2 sample (2M reads) bench-faithful per-window O(K) re-encode", "cores": cores, "kmers_per_s": ns * 120 / t})
    n3 = 200_000 // scale
    t = cpu_time(lambda: ko.extract_canonical_wide(hb[:n3 * 150], 63, n_reads=n3, fixed_len=150, want_hash=False, n_threads=cores), reps=1)
    cpu_rows.append({"config": "config3 sample (200k reads) K=63 encode + swap-loop rev_comp per window", "cores": cores, "kmers_per_s": n3 * 88 / t})
    n4 = 20_000 // scale
    hb4 = ko.generate_bases(43, 0, n4 * 10_000, 1049)
    t = cpu_time(lambda: ko.extract_canonical(hb4, 31, n_reads=n4, fixed_len=10_000, n_threads=cores, materialize=False))
    cpu_rows.append({"config": "config4 sample (20k x 10 kbp) iterator, digest only", "cores": cores, "kmers_per_s": n4 * 9970 / t})
    g5 = 200_000_000 // scale
    hb5 = ko.generate_bases(44, 0, g5, 105)
    t = cpu_time(lambda: ko.extract_canonical(hb5, 31, n_reads=1000, fixed_len=g5 // 1000, n_threads=cores, materialize=False, hist_bits=16), reps=1)
    cpu_rows.append({"config": "config5 sample (200 Mbp as 1000 chunks) iterator + 2^16-bin histogram", "cores": cores, "kmers_per_s": (g5 - 30 * 1000) / t})
    for r in cpu_rows[4:]:
        print(f"CPU {r['config']:90s} {r['kmers_per_s'] / 1e6:10.1f} M kmers/s on {r['cores']} core(s)", flush=True)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"hbm_peak_gbs": pk, "host_cores": cores, "rows": rows, "cpu_rows": cpu_rows}, open(args.out, "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
