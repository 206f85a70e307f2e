"""Randomised differential test: the CUDA path against the oracle over random batch geometries (read lengths around
and below K, ragged / fixed / packed, invalid-base rates, every K class), one seeded case per parameter.  The fixed
parametrised tests pin the known corners; this sweep looks for the unknown ones."""
import numpy as np
import pytest

from golden_util import random_reads

import os

pytestmark = pytest.mark.gpu
SENT = np.uint64(2**64 - 1)
SCALE = int(os.environ.get("KMB_FUZZ_SCALE", "1"))  # KMB_FUZZ_SCALE=10 runs ten times as many seeds
BASES = int(os.environ.get("KMB_FUZZ_BASES", "400000"))  # approximate size of a case


@pytest.fixture(scope="module")
def ctx():
    import kmers_b200 as kb
    c = kb.Context(0)
    yield c
    c.close()


def _case(seed):
    rng = np.random.default_rng(seed)
    k = int(rng.choice([1, 2, 5, 11, 15, 16, 17, 21, 27, 31, 32, int(rng.integers(1, 33))]))
    ragged = bool(rng.integers(0, 2))
    span = int(rng.choice([3, 9, 20, 40, 130, 400, 3000]))
    lo = max(0, k - int(rng.integers(0, 4)))
    n = int(rng.integers(1, max(2, BASES // (lo + span))))
    p_bad = float(rng.choice([0.0, 0.0005, 0.01, 0.2]))
    if ragged:
        bases, offs = random_reads(rng, n, lo, lo + span, p_bad=p_bad)
    else:
        L = lo + int(rng.integers(0, span + 1))
        bases, offs = random_reads(rng, n, L, L, p_bad=p_bad)
    return rng, k, ragged, bases, offs


def _upload(ctx, ragged, bases, offs):
    if ragged:
        return ctx.upload(bases, offsets=offs)
    n = offs.size - 1
    return ctx.upload(bases, fixed_len=int(offs[1]) if n else 0, n_reads=n)


@pytest.mark.parametrize("seed", range(40 * SCALE))
def test_fuzz_extract_compact_histogram(ctx, seed):
    import oracle as ko
    rng, k, ragged, bases, offs = _case(seed)
    if bases.size == 0:
        pytest.skip("empty batch")
    ref = ko.extract_canonical(bases, k, offsets=offs, want_fw_rc=True, n_threads=4)
    batch = _upload(ctx, ragged, bases, offs)
    res = batch.extract_canonical(k, want_fw_rc=True, digest=True, to="host")
    for name in ("canon", "hash", "fw", "rc"):
        assert np.array_equal(getattr(res, name), ref[name]), (seed, name, k, ragged)
    assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"]), seed
    # compacted stream = the dense arrays without their sentinel slots, positions from the window offsets
    c = batch.extract_compact(k)
    keep = ref["canon"] != SENT if k < 32 else np.ones(ref["canon"].size, dtype=bool)
    if k < 32:
        assert c["n"] == int(keep.sum()), seed
        assert np.array_equal(c["canon"], ref["canon"][keep]) and np.array_equal(c["hash"], ref["hash"][keep]), seed
        lens = np.diff(offs.astype(np.int64))
        w = np.maximum(lens - k + 1, 0)
        pos = np.concatenate([np.arange(x) for x in w]) if w.sum() else np.zeros(0, dtype=np.int64)
        assert np.array_equal(c["pos"], pos[keep].astype(np.int32)), seed
    bits = int(rng.integers(1, min(2 * k, 18) + 1))
    hist, dig = batch.histogram(k, bits, to="host")
    want = ko.extract_canonical(bases, k, offsets=offs, hist_bits=bits, materialize=False, n_threads=4)
    assert np.array_equal(hist, want["hist"]) and dig[0] == ref["n_valid"], (seed, bits)


@pytest.mark.parametrize("seed", range(10_000, 10_000 + 30 * SCALE))
def test_fuzz_minimizers_and_wide(ctx, seed):
    import kmers_b200 as kb
    import oracle as ko
    rng, k, ragged, bases, offs = _case(seed)
    if bases.size == 0:
        pytest.skip("empty batch")
    w = int(rng.integers(1, k + 1))
    hk = int(rng.choice([w, int(rng.integers(1, 33))]))
    batch = _upload(ctx, ragged, bases, offs)
    mm, pos = batch.minimizers(k, w, hk)
    rmm, rpos = ko.minimizers_batch(bases, k, w, hk, offsets=offs)
    assert np.array_equal(mm, rmm) and np.array_equal(pos, rpos), (seed, k, w, hk, ragged)
    kw = int(rng.integers(1, 65))
    enc_name = str(rng.choice(["ACGT", "ACTG", "TGCA", "GATC"]))
    res = _upload(ctx, ragged, bases, offs).extract_canonical_wide(kw, ko.NAIVE[enc_name], digest=True, to="host")
    ref = ko.extract_canonical_wide(bases, kw, enc=ko.NAIVE[enc_name], offsets=offs)
    assert np.array_equal(res.host("canon"), ref["canon"]) and np.array_equal(res.host("hash"), ref["hash"]), (seed, kw, enc_name, ragged)
    assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"]), seed


@pytest.mark.parametrize("seed", range(20_000, 20_000 + 15 * SCALE))
def test_fuzz_packed_store(ctx, seed):
    """Valid-only reads through the 2-bit packed store == the ASCII batch (narrow, compact, minimizers)."""
    rng, k, ragged, bases, offs = _case(seed)
    if bases.size == 0:
        pytest.skip("empty batch")
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=bases.size)]
    a = _upload(ctx, ragged, bases, offs).extract_canonical(k, want_fw_rc=True, digest=True, to="host")
    w = int(rng.integers(1, k + 1))
    am = _upload(ctx, ragged, bases, offs).minimizers(k, w)
    p = _upload(ctx, ragged, bases, offs).to_packed()
    b = p.extract_canonical(k, want_fw_rc=True, digest=True, to="host")
    for name in ("canon", "hash", "fw", "rc"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), (seed, name)
    assert a.digest == b.digest
    bm = p.minimizers(k, w)
    assert np.array_equal(am[0], bm[0]) and np.array_equal(am[1], bm[1]), seed
