#!/usr/bin/env python
"""Run ONE kernel family a few times so that ncu can capture it:  python scripts/prof_one.py wide|hist|compact|minimizers|csr|canon|full|pack8|pack64"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import kmers_b200 as kb
from kmers_b200.context import _ptr

what = sys.argv[1]
n, L, K = 4_000_000, 150, 31
torch.cuda.set_device(0)
ctx = kb.Context(0)
i64 = lambda m: torch.empty(m, dtype=torch.int64, device="cuda")
batch = ctx.generate(42, n, L, n_thresh20=1049 if what == "compact" else 0)
for _ in range(3):
    if what == "wide":
        w = L - 63 + 1
        c = i64(2 * n * w)
        ctx._ck(ctx._lib.kmb_extract_canonical_wide(ctx._h, 63, kb.ENC_ACGT, 0, _ptr(c), None, None))
    elif what == "hist":
        batch.histogram(K, 16, digest=False)
    elif what == "compact":
        r = batch.extract_compact(K, to="device")
    elif what == "minimizers":
        batch.minimizers(31, 15, to="device")
    elif what == "csr":
        if _ == 0:
            offs = torch.arange(0, n + 1, dtype=torch.int64, device="cuda") * L   # same reads, ragged (CSR) geometry
            bases = batch.download()
            batch = ctx.upload(bases, offsets=offs.cpu().numpy().astype("uint64"))
            out = kb.CanonicalKmers(k=K, n_slots=n * (L - K + 1), canon=i64(n * (L - K + 1)), hash=i64(n * (L - K + 1)))
        batch.extract_canonical(K, out=out)
    elif what == "full":
        if _ == 0:
            out = kb.CanonicalKmers(k=K, n_slots=n * (L - K + 1), canon=i64(n * (L - K + 1)), hash=i64(n * (L - K + 1)))
        batch.extract_canonical(K, out=out)
    elif what == "pack8":
        batch.pack(kb.ENC_ACGT, 8, to="device")
    elif what == "pack64":
        batch.pack(kb.ENC_ACGT, 64, to="device")
    elif what == "canon":
        out = kb.CanonicalKmers(k=K, n_slots=n * (L - K + 1), canon=i64(n * (L - K + 1)), hash=None)
        batch.extract_canonical(K, out=out)
    ctx.sync()
ctx.close()
