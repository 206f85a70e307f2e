#!/usr/bin/env python
"""Run ONE kernel family a few times (for ncu), or time it with CUDA events:
    python scripts/prof_one.py wide|hist|compact|minimizers|csr|canon|full|pack8|pack64|csr_wide|csr_min|csr_var|revcomp [--time]
4M x 150 bp reads, K=31 (wide: K=63).  --time prints ms per launch and algorithmic GB/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import kmers_b200 as kb
from kmers_b200.context import _ptr

what = sys.argv[1]
timed = "--time" in sys.argv
scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0  # problem size, times the default
n, L, K = int(4_000_000 * scale), 150, (int(sys.argv[sys.argv.index("--k") + 1]) if "--k" in sys.argv else 31)
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
ctx = kb.Context(0, stream=stream.cuda_stream)
i64 = lambda m: torch.empty(m, dtype=torch.int64, device="cuda")
batch = ctx.generate(42, n, L, n_thresh20=1049 if "compact" in what else 0)
W = L - K + 1
n_slots = n * W
if what in ("csr_var", "csr_var_compact"):  # ragged reads of 100..150 bases cut from the same stream
    import numpy as np
    lens = np.random.default_rng(1).integers(100, 151, size=n).astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    batch = ctx.upload(batch.download()[:int(offs[-1])], offsets=offs)
    n_slots, n_bases = int((lens - (K - 1)).sum()), int(offs[-1])
elif what.startswith("csr"):  # same reads, ragged (CSR) geometry
    offs = (torch.arange(0, n + 1, dtype=torch.int64) * L).numpy().astype("uint64")
    batch = ctx.upload(batch.download(), offsets=offs)

if what in ("wide", "csr_wide"):
    w = L - 63 + 1
    c = i64(2 * n * w)
    alg = n * (L + w * 16)
    step = lambda: ctx._ck(ctx._lib.kmb_extract_canonical_wide(ctx._h, 63, kb.ENC_ACGT, 0, _ptr(c), None, None))
elif what == "hist":
    alg = n * L
    hb = i64(65536 + 3)
    step = lambda: batch.histogram(K, 16, hist=hb, digest=False)
elif what == "histd":  # with the digest accumulated behind the bins (config 5's step)
    alg = n * L
    hb = i64(65536 + 3)
    step = lambda: batch.histogram(K, 16, hist=hb, digest_in_hist=True)
elif what == "compact":
    alg = None
    step = lambda: batch.extract_compact(K, to="device")
elif what in ("minimizers", "csr_min"):
    alg = n * (L + W * 12)
    mm, pp = i64(n * W), torch.empty(n * W, dtype=torch.int32, device="cuda")
    step = lambda: ctx._ck(ctx._lib.kmb_minimizers(ctx._h, 31, 15, 15, 0, _ptr(mm), _ptr(pp)))
elif what in ("csr", "full", "csr_var"):
    alg = n * (L + W * 16) if what != "csr_var" else n_bases + n_slots * 16
    out = kb.CanonicalKmers(k=K, n_slots=n_slots, canon=i64(n_slots), hash=i64(n_slots))
    step = lambda: batch.extract_canonical(K, out=out)
elif what == "revcomp":
    import numpy as np
    ni = int(32_000_000 * scale)
    words = torch.randint(0, 2**62, (2 * ni,), dtype=torch.int64, device="cuda")
    outw = torch.empty_like(words)
    alg = ni * 32
    step = lambda: ctx._ck(ctx._lib.kmb_revcomp_words(ctx._h, kb.ENC_ACGT, 63, 64, 2, _ptr(words), _ptr(outw), ni))
elif what == "canon":
    alg = n * (L + W * 8)
    out = kb.CanonicalKmers(k=K, n_slots=n * W, canon=i64(n * W), hash=None)
    step = lambda: batch.extract_canonical(K, out=out)
elif what == "minword":
    nk = int(50_000_000 * scale)
    words = torch.randint(0, 2**62, (nk,), dtype=torch.int64, device="cuda")
    mmw, mmo = torch.empty_like(words), torch.empty(nk, dtype=torch.int32, device="cuda")
    alg = nk * 20
    step = lambda: ctx._ck(ctx._lib.kmb_minimizer_words(ctx._h, 31, 15, 15, _ptr(words), nk, _ptr(mmw), _ptr(mmo)))
elif what == "unpack":
    ni = int(25_000_000 * scale)
    words = torch.randint(0, 2**62, (ni,), dtype=torch.int64, device="cuda")
    txt = torch.empty(ni * 31, dtype=torch.uint8, device="cuda")
    alg = ni * (8 + 31)
    step = lambda: ctx._ck(ctx._lib.kmb_unpack(ctx._h, kb.ENC_ACGT, 64, _ptr(words), ni, 1, 31, _ptr(txt)))
elif what in ("csr_compact", "csr_var_compact"):  # the ragged emit launch(es) with worst-case arrays
    import ctypes as C
    cc, ch, cp, ce = i64(n_slots), i64(n_slots), torch.empty(n_slots, dtype=torch.int32, device="cuda"), i64(n + 1)
    cnt = C.c_uint64()
    step = lambda: ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, K, 0, _ptr(cc), _ptr(ch), _ptr(cp), _ptr(ce), n_slots, C.byref(cnt)))
    step()
    alg = (n_bases if what == "csr_var_compact" else n * L) + 20 * int(cnt.value) + 8 * n
elif what == "compact1":  # the single emit launch with worst-case arrays
    import ctypes as C
    alg = None
    cc, ch, cp, ce = i64(n_slots), i64(n_slots), torch.empty(n_slots, dtype=torch.int32, device="cuda"), i64(n + 1)
    cnt = C.c_uint64()
    step = lambda: ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, K, 0, _ptr(cc), _ptr(ch), _ptr(cp), _ptr(ce), n_slots, C.byref(cnt)))
elif what == "copy":  # a plain device copy of the same size class, for the size dependence of the copy peak itself
    nb = int(512_000_000 * scale)
    src = torch.empty(nb, dtype=torch.uint8, device="cuda")
    dst = torch.empty_like(src)
    alg = 2 * nb
    def step():
        with torch.cuda.stream(stream):
            dst.copy_(src)
elif what in ("pack8", "pack64"):
    alg = n * L + n * ((L + 31) // 32) * 8
    bits = int(what[4:])
    step = lambda: batch.pack(kb.ENC_ACGT, bits, to="device")
else:
    raise SystemExit(f"unknown kernel family {what}")

for _ in range(3):
    step()
ctx.sync()
if timed:
    reps = 20
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(stream)
        step()
        b.record(stream)
    ctx.sync()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    avg = sum(ms) / len(ms)
    row = {"kernel": what, "avg_ms": round(avg, 4), "best_ms": round(ms[0], 4)}
    if alg:
        row.update(algorithmic_GB=alg / 1e9, GBps=round(alg / avg / 1e6, 1), frac_of_6535=round(alg / avg / 1e6 / 6535.7, 3))
    print(json.dumps(row), flush=True)
ctx.close()
