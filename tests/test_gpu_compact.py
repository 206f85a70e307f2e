"""Iterator-identical output (kmb_extract_compact): the exact (pos, canonical word, LexHash) sequence
CanonicalKmerIterator yields (naive_impl/canonical_kmer_iterator.rs:13-16, 42-101), read after read."""
import ctypes as C

import numpy as np
import pytest

from golden_util import load_goldens, random_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import kmers_b200 as kb
    c = kb.Context(0)
    yield c
    c.close()


def _reference_sequence(ko, bases, offs, k):
    """while !it.exhausted() { it.get(); it.inc(); } over every read, with the restated iterator."""
    L = ko.lib()
    pos, canon, hsh, starts = [], [], [], [0]
    for r in range(len(offs) - 1):
        it = ko.Iter(bases[int(offs[r]):int(offs[r + 1])].tobytes(), k)
        while not it.exhausted():
            km = it.km
            c = L.ko_ck_get_canonical_word(C.byref(km))
            pos.append(it.pos)
            canon.append(c)
            hsh.append(L.ko_lexhash_word(c, k))
            it.inc()
        starts.append(len(pos))
    return (np.array(pos, dtype=np.int32), np.array(canon, dtype=np.uint64), np.array(hsh, dtype=np.uint64),
            np.array(starts, dtype=np.uint64))


def _from_dense(ko, bases, k, offsets=None, n_reads=None, fixed_len=0):
    ref = ko.extract_canonical(bases, k, offsets=offsets, n_reads=n_reads, fixed_len=fixed_len, n_threads=4)
    if offsets is None:
        w = max(0, fixed_len - k + 1)
        woff = np.arange(n_reads + 1, dtype=np.int64) * w
    else:
        lens = np.diff(np.asarray(offsets).astype(np.int64))
        woff = np.concatenate([[0], np.cumsum(np.maximum(0, lens - k + 1))])
    ok = ref["canon"] != ko.SENTINEL
    slots = np.flatnonzero(ok)
    read_of = np.searchsorted(woff, slots, side="right") - 1
    pos = (slots - woff[read_of]).astype(np.int32)
    emit = np.concatenate([[0], np.cumsum(ok)])[woff].astype(np.uint64)
    return pos, ref["canon"][ok], ref["hash"][ok], emit


def _check(got, want):
    pos, canon, hsh, emit = want
    assert got["n"] == pos.size
    assert np.array_equal(got["pos"], pos)
    assert np.array_equal(got["canon"], canon)
    assert np.array_equal(got["hash"], hsh)
    assert np.array_equal(got["emit_offsets"], emit)


def test_compact_reference_reads(ctx):
    """The reference's own iterator test reads (canonical_kmer_iterator.rs:123-206), incl. the N cases."""
    import oracle as ko
    reads = [c["read"].encode() for c in load_goldens()["iterator"]["cases"]] + [b"ACGT" * 5, b"", b"N" * 40]
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8)
    got = ctx.upload(bases, offsets=offs).extract_compact(31)
    _check(got, _reference_sequence(ko, bases, offs, 31))
    # N at index 35 of the 5th read: positions 0..4 then 36.. (canonical_kmer_iterator.rs:178-189)
    s, e = int(got["emit_offsets"][4]), int(got["emit_offsets"][5])
    assert got["pos"][s:e].tolist() == [0, 1, 2, 3, 4] + list(range(36, 121 - 31 + 1))


@pytest.mark.parametrize("k", [1, 5, 16, 17, 31, 32])
def test_compact_ragged_vs_iterator(ctx, k):
    import oracle as ko
    rng = np.random.default_rng(40 + k)
    bases, offs = random_reads(rng, 300, 0, 130, p_bad=0.03)
    got = ctx.upload(bases, offsets=offs).extract_compact(k)
    _check(got, _reference_sequence(ko, bases, offs, k))


@pytest.mark.parametrize("L,k,p_bad", [(150, 31, 0.01), (150, 21, 0.02), (33, 31, 0.01), (31, 31, 0.0), (10007, 31, 0.002), (60, 31, 0.5)])
def test_compact_fixed_vs_dense(ctx, L, k, p_bad):
    import oracle as ko
    rng = np.random.default_rng(L + k)
    n = max(4, 300000 // L)
    bases, _ = random_reads(rng, n, L, L, p_bad=p_bad)
    got = ctx.upload(bases, fixed_len=L).extract_compact(k)
    _check(got, _from_dense(ko, bases, k, n_reads=n, fixed_len=L))


def test_compact_ragged_large_and_device_outputs(ctx):
    import oracle as ko
    rng = np.random.default_rng(99)
    b1, o1 = random_reads(rng, 20000, 20, 300, p_bad=0.004)
    b2, o2 = random_reads(rng, 3000, 0, 40, p_bad=0.0)       # a stretch of reads mostly too short for a window
    b3, o3 = random_reads(rng, 5, 20000, 50000, p_bad=0.001)
    bases = np.concatenate([b1, b2, b3])
    offs = np.concatenate([o1, o2[1:] + o1[-1], o3[1:] + o1[-1] + o2[-1]])
    want = _from_dense(ko, bases, 31, offsets=offs)
    host = ctx.upload(bases, offsets=offs).extract_compact(31, to="host")
    _check(host, want)
    dev = ctx.upload(bases, offsets=offs).extract_compact(31, to="device")
    assert np.array_equal(dev["canon"].cpu().numpy().view(np.uint64), want[1])
    assert np.array_equal(dev["pos"].cpu().numpy(), want[0])
    assert np.array_equal(dev["emit_offsets"].cpu().numpy().view(np.uint64), want[3])


def test_compact_capacity_and_empty(ctx):
    import kmers_b200 as kb
    n = C.c_uint64()
    b = ctx.upload(b"ACGTACGTAC" * 10, fixed_len=100)
    lib, h = ctx._lib, ctx._h
    assert lib.kmb_extract_compact(h, 31, 0, None, None, None, None, 0, C.byref(n)) == 0 and n.value == 70
    buf = np.zeros(10, dtype=np.uint64)
    assert lib.kmb_extract_compact(h, 31, 0, buf.ctypes.data, None, None, None, 10, C.byref(n)) == kb._native.ERR_INVALID_ARG
    assert n.value == 70  # the count is still reported so the caller can re-allocate
    got = ctx.upload(b"NNNN" * 20, fixed_len=80).extract_compact(31)
    assert got["n"] == 0 and got["emit_offsets"].tolist() == [0, 0]
    got = ctx.upload(b"ACGT" * 5, fixed_len=20).extract_compact(31)  # reads shorter than k
    assert got["n"] == 0 and got["emit_offsets"].tolist() == [0, 0]
    with pytest.raises(kb.KmbPanic):
        b.extract_compact(33)


def test_counting_call_then_emit_reuses_counts(ctx):
    """The sizing protocol (count with NULL outputs, then emit) in every order: the emit call is self-contained (it counts
    again from its staged tiles), so nothing a counting call left behind -- another k, another batch -- can leak into it."""
    import oracle as ko
    rng = np.random.default_rng(77)
    bases, _ = random_reads(rng, 3000, 150, 150, p_bad=0.01)
    lib, h = ctx._lib, ctx._h
    batch = ctx.upload(bases, fixed_len=150)

    def count(k):
        n = C.c_uint64()
        ctx._ck(lib.kmb_extract_compact(h, k, 0, None, None, None, None, 0, C.byref(n)))
        return int(n.value)

    def emit(k, cap):
        canon, pos, n = np.zeros(cap, dtype=np.uint64), np.zeros(cap, dtype=np.int32), C.c_uint64()
        ctx._ck(lib.kmb_extract_compact(h, k, 0, canon.ctypes.data, None, pos.ctypes.data, None, cap, C.byref(n)))
        return canon[: n.value], pos[: n.value]

    n31 = count(31)
    c_a, p_a = emit(31, n31)            # reuses the counts
    c_b, p_b = emit(31, n31)            # counts again
    assert np.array_equal(c_a, c_b) and np.array_equal(p_a, p_b)
    n21 = count(21)
    c31, p31 = emit(31, n31)            # different k: the pending counts of k=21 must not be used
    assert np.array_equal(c31, c_a) and np.array_equal(p31, p_a)
    count(31)
    bases2, _ = random_reads(rng, 3000, 150, 150, p_bad=0.05)
    ctx.upload(bases2, fixed_len=150)    # new batch: pending counts dropped
    n2 = count(31)
    c2, p2 = emit(31, n2)
    r = ko.extract_canonical(bases2, 31, n_reads=3000, fixed_len=150)
    keep = r["canon"] != np.uint64(2**64 - 1)
    assert n2 == int(keep.sum()) and np.array_equal(c2, r["canon"][keep])
    assert n21 > n31


def test_worst_case_capacity_needs_no_counting_call(ctx):
    """Arrays that hold one entry per slot: a single emit call (one launch, decoupled look-back) returns the stream and its
    length; thousands of tiles, so every look-back pattern (aggregate chains, prefixes, the 32-wide window) occurs."""
    import torch
    import oracle as ko
    n, L, k = 400_000, 150, 31
    W = L - k + 1
    bases = ko.generate_bases(5, 0, n * L, n_thresh20=3000)
    b = ctx.upload(bases, fixed_len=L)
    cap = n * W
    canon = torch.empty(cap, dtype=torch.int64, device="cuda")
    hsh = torch.empty(cap, dtype=torch.int64, device="cuda")
    pos = torch.empty(cap, dtype=torch.int32, device="cuda")
    offs = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    m = C.c_uint64()
    l0 = ctx.launch_count
    ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, k, 0, canon.data_ptr(), hsh.data_ptr(), pos.data_ptr(), offs.data_ptr(), cap, C.byref(m)))
    assert ctx.launch_count == l0 + 1  # ONE kernel (fixed-length reads: it writes final emit_offsets too): no counting launch, no scan launch
    m = int(m.value)
    ref = ko.extract_canonical(bases, k, n_reads=n, fixed_len=L, n_threads=8)
    keep = ref["canon"] != np.uint64(2**64 - 1)
    assert m == int(keep.sum()) == ref["n_valid"]
    assert np.array_equal(canon[:m].cpu().numpy().view(np.uint64), ref["canon"][keep])
    assert np.array_equal(hsh[:m].cpu().numpy().view(np.uint64), ref["hash"][keep])
    assert np.array_equal(pos[:m].cpu().numpy(), np.tile(np.arange(W, dtype=np.int32), n)[keep])
    per_read = keep.reshape(n, W).sum(axis=1)
    want_offs = np.zeros(n + 1, dtype=np.uint64)
    want_offs[1:] = np.cumsum(per_read)
    assert np.array_equal(offs.cpu().numpy().view(np.uint64), want_offs)
    # repeated launches reuse the descriptor array: it must be reset every time
    for _ in range(3):
        m2 = C.c_uint64()
        l1 = ctx.launch_count
        ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, k, 0, canon.data_ptr(), None, None, None, cap, C.byref(m2)))
        assert int(m2.value) == m and ctx.launch_count == l1 + 1  # without emit_offsets: ONE kernel
    assert np.array_equal(canon[:m].cpu().numpy().view(np.uint64), ref["canon"][keep])


def test_count_then_repack_then_emit(ctx):
    """A counting call on the ASCII batch (validation on), then the batch is switched to the packed store without
    validation (every window becomes valid), then the emit call: it must describe the batch as it is NOW."""
    rng = np.random.default_rng(3)
    n, L, k = 2000, 150, 31
    bases, _ = random_reads(rng, n, L, L, p_bad=0.02)
    b = ctx.upload(bases, fixed_len=L)
    m = C.c_uint64()
    ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, k, 0, None, None, None, None, 0, C.byref(m)))
    n_ascii = int(m.value)
    assert n_ascii < n * (L - k + 1)
    b.to_packed(strict=False)
    cap = n * (L - k + 1)
    canon = np.zeros(cap, dtype=np.uint64)
    ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, k, 0, canon.ctypes.data, None, None, None, cap, C.byref(m)))
    assert int(m.value) == cap  # a packed store holds no invalid base: every window is emitted, and none beyond the arrays
    dense = b.extract_canonical(k, to="host")
    assert np.array_equal(canon, dense.canon)


@pytest.mark.parametrize("n_reads", [1, 60, 69, 137, 300, 341])  # 120 windows per read, 8192 per tile: 1, 1, 2, 3, 5 and 5 tiles
def test_pipelined_kernel_tile_counts_subsets_and_alignment(ctx, n_reads):
    """Fixed-length reads run through the persistent, software-pipelined kernel: its prologue (first tile of a CTA), the
    steady state and the last tile; any subset of the three arrays; arrays that are not 16-byte aligned (scalar write-out);
    a capacity that ends inside a tile."""
    import torch
    import oracle as ko
    rng = np.random.default_rng(n_reads)
    L, k = 150, 31
    bases, _ = random_reads(rng, n_reads, L, L, p_bad=0.004)
    want = _from_dense(ko, bases, k, n_reads=n_reads, fixed_len=L)
    m = want[0].size
    ctx.upload(bases, fixed_len=L)
    lib, h = ctx._lib, ctx._h
    nn = C.c_uint64()

    def run(canon, hsh, pos, emit, cap):
        torch.cuda.synchronize()
        rc = lib.kmb_extract_compact(h, k, 0, _dp(canon), _dp(hsh), _dp(pos), _dp(emit), cap, C.byref(nn))
        torch.cuda.synchronize()
        return rc

    def _dp(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def fresh(off=0):
        cc = torch.full((m + 8,), -7, dtype=torch.int64, device="cuda")
        ch = torch.full((m + 8,), -7, dtype=torch.int64, device="cuda")
        cp = torch.full((m + 8,), -7, dtype=torch.int32, device="cuda")
        ce = torch.full((n_reads + 1,), -7, dtype=torch.int64, device="cuda")
        return cc[off:], ch[off:], cp[off:], ce

    def same(t, ref, dtype):
        return np.array_equal(t[:m].cpu().numpy().view(dtype), ref) and bool((t[m:] == -7).all())  # nothing written past the end

    # every subset of the arrays, aligned (16-byte vector write-out, with and without the "all three" fast path)
    for use in [(1, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1)]:
        cc, ch, cp, ce = fresh()
        assert run(cc if use[0] else None, ch if use[1] else None, cp if use[2] else None, ce, m) == 0 and nn.value == m
        assert not use[0] or same(cc, want[1], np.uint64)
        assert not use[1] or same(ch, want[2], np.uint64)
        assert not use[2] or same(cp, want[0], np.int32)
        assert np.array_equal(ce.cpu().numpy().view(np.uint64), want[3])
    # arrays 8 (4) bytes off a 16-byte (8-byte) boundary: the scalar write-out
    cc, ch, cp, ce = fresh(off=1)
    assert run(cc, ch, cp, ce, m) == 0 and nn.value == m
    assert same(cc, want[1], np.uint64) and same(ch, want[2], np.uint64) and same(cp, want[0], np.int32)
    # a capacity that ends inside a tile: the call fails, reports the count, and nothing is written at or beyond the capacity
    if m > 10:
        import kmers_b200 as kb
        cap = m - min(m // 2, 5000) - 3
        cc, ch, cp, ce = fresh()
        assert run(cc, ch, cp, ce, cap) == kb._native.ERR_INVALID_ARG and nn.value == m
        assert np.array_equal(cc[:cap].cpu().numpy().view(np.uint64), want[1][:cap]) and bool((cc[cap:] == -7).all())
        assert np.array_equal(cp[:cap].cpu().numpy(), want[0][:cap]) and bool((cp[cap:] == -7).all())
