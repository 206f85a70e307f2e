""""Next" row N2: the batch as a 2-bit packed sequence store (SeqVector twin, naive_impl/seq_vector.rs)."""
import numpy as np
import pytest

from golden_util import random_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import kmers_b200 as kb
    c = kb.Context(0)
    yield c
    c.close()


def _acgt(rng, n):
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].copy()


def test_seq_vector_goldens(ctx):
    """seq_vector.rs:342-358 iter_kmers on 'ACTTGAT'; :328-339 push_chars 30 A + 40 C; get_kmer_u64 == oracle."""
    import oracle as ko
    s = b"ACTTGAT"
    b = ctx.upload(s, fixed_len=7).to_packed()
    res = b.extract_canonical(3, want_fw_rc=True, to="host")
    assert [ko.kmer_str(ko.Kmer(3, int(w))) for w in res.fw] == ["act", "ctt", "ttg", "tga", "gat"]
    long = b"A" * 30 + b"C" * 40
    b = ctx.upload(long, fixed_len=70).to_packed()
    assert "".join("ACGT"[int(v)] for v in b.get_kmers(1, np.arange(70))).encode() == long
    words = ko.sv_from_bytes(long)
    pos = np.array([0, 1, 29, 30, 31, 38, 69], dtype=np.uint64)
    for k in (1, 7, 31, 32):
        got = b.get_kmers(k, pos)
        for p, g in zip(pos.tolist(), got.tolist()):
            want = ko.sv_get_kmer_u64(words, 70, p, k) if p + k <= 70 else ko.SENTINEL
            assert g == want, (k, p)


@pytest.mark.parametrize("L,k", [(150, 31), (150, 21), (64, 31), (33, 31), (31, 31), (1000, 32), (10007, 15), (7, 3)])
def test_packed_fixed_matches_ascii(ctx, L, k):
    """Every op that accepts a packed batch gives the same result as on the ASCII bytes (pure-ACGT reads)."""
    rng = np.random.default_rng(L * 3 + k)
    n = max(3, 40000 // L)
    bases = _acgt(rng, n * L)
    a = ctx.upload(bases, fixed_len=L)
    ra = a.extract_canonical(k, want_fw_rc=True, digest=True, to="host")
    ca = a.extract_compact(k)
    ma = a.minimizers(k, max(1, k // 2))
    ha, _ = a.histogram(k, min(2 * k, 10), to="host")
    p = ctx.upload(bases, fixed_len=L).to_packed()
    rp = p.extract_canonical(k, want_fw_rc=True, digest=True, to="host")
    for name in ("canon", "hash", "fw", "rc"):
        assert np.array_equal(getattr(ra, name), getattr(rp, name)), name
    assert ra.digest == rp.digest
    cp = p.extract_compact(k)
    assert all(np.array_equal(ca[x], cp[x]) for x in ("pos", "canon", "hash", "emit_offsets"))
    mp = p.minimizers(k, max(1, k // 2))
    assert np.array_equal(ma[0], mp[0]) and np.array_equal(ma[1], mp[1])
    hp, _ = p.histogram(k, min(2 * k, 10), to="host")
    assert np.array_equal(ha, hp)


def test_packed_ragged_matches_ascii_and_oracle(ctx):
    import oracle as ko
    rng = np.random.default_rng(12)
    lens = np.concatenate([rng.integers(0, 200, size=700), [5000, 0, 31, 32, 33, 64, 65]])
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    bases = _acgt(rng, int(offs[-1]))
    p = ctx.upload(bases, offsets=offs).to_packed()
    for k in (31, 16, 5):
        rp = p.extract_canonical(k, digest=True, to="host")
        ref = ko.extract_canonical(bases, k, offsets=offs)
        assert np.array_equal(rp.canon, ref["canon"]) and np.array_equal(rp.hash, ref["hash"])
        assert rp.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])
    mm, pos = p.minimizers(31, 15)
    rmm, rpos = ko.minimizers_batch(bases, 31, 15, 15, offsets=offs)
    assert np.array_equal(mm, rmm) and np.array_equal(pos, rpos)
    got = p.get_kmers(9, np.array([0, 100, 4990], dtype=np.uint64), reads=np.array([700, 700, 700], dtype=np.uint64))
    start = int(offs[700])
    for g, q in zip(got.tolist(), (0, 100, 4990)):
        assert g == ko.kmer_from(bases[start + q:start + q + 9].tobytes()).data


def test_attach_packed_device_words(ctx):
    """kmb_pack output handed back as a borrowed packed batch (zero copy)."""
    import torch
    import kmers_b200 as kb
    rng = np.random.default_rng(4)
    n, L, k = 500, 150, 31
    bases = _acgt(rng, n * L)
    a = ctx.upload(bases, fixed_len=L)
    want = a.extract_canonical(k, to="host")
    img, _ = a.pack(kb.ENC_ACGT, 64, to="device")
    lib, h = ctx._lib, ctx._h
    ctx._ck(lib.kmb_batch_attach_packed(h, img.data_ptr(), img.numel() // 8, None, None, n, L))
    got = kb.ReadBatch(ctx, img.numel(), n, L, False).extract_canonical(k, to="host")
    assert np.array_equal(got.canon, want.canon) and np.array_equal(got.hash, want.hash)
    del img


def test_repack_strict_panics_on_invalid_bases(ctx):
    import kmers_b200 as kb
    bases, _ = random_reads(np.random.default_rng(1), 50, 100, 100, p_bad=0.02)
    with pytest.raises(kb.KmbPanic):
        ctx.upload(bases, fixed_len=100).to_packed(strict=True)
    # one resident batch per context: take the ASCII result first, then repack
    a = ctx.upload(bases, fixed_len=100).extract_canonical(31, validate=False, to="host")
    p = ctx.upload(bases, fixed_len=100).to_packed(strict=False)  # Encoding::encode semantics: (c >> 1) & 3
    r = p.extract_canonical(31, to="host")
    assert np.array_equal(a.canon, r.canon) and np.array_equal(a.hash, r.hash)


@pytest.mark.parametrize("enc_name", ["ACGT", "TGCA", "XOR10"])
def test_two_word_path_reads_the_packed_store(ctx, enc_name):
    """kmb_extract_canonical_wide on a packed batch (fixed and ragged) == the same call on the ASCII batch, for the
    store's own code and for other encodings (the staged words are re-coded on the fly)."""
    import kmers_b200 as kb
    import oracle as ko
    rng = np.random.default_rng(len(enc_name) + 40)
    enc = kb.ENC_XOR10 if enc_name == "XOR10" else ko.NAIVE[enc_name]
    for k in (31, 47, 63):
        bases = _acgt(rng, 700 * 150)
        a = ctx.upload(bases, fixed_len=150).extract_canonical_wide(k, enc, digest=True, to="host")
        p = ctx.upload(bases, fixed_len=150).to_packed().extract_canonical_wide(k, enc, digest=True, to="host")
        assert np.array_equal(a.canon, p.canon) and np.array_equal(a.hash, p.hash) and a.digest == p.digest, k
    lens = rng.integers(0, 200, size=900)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    bases = _acgt(rng, int(offs[-1]))
    a = ctx.upload(bases, offsets=offs).extract_canonical_wide(63, enc, digest=True, to="host")
    p = ctx.upload(bases, offsets=offs).to_packed().extract_canonical_wide(63, enc, digest=True, to="host")
    assert np.array_equal(a.canon, p.canon) and np.array_equal(a.hash, p.hash) and a.digest == p.digest


def test_push_chars_grows_a_sequence_like_seqvector(ctx):
    """SeqVector::with_capacity + push_chars (seq_vector.rs:135-161), pieces of every length class (the reference pushes a
    partial first word and then whole words; the bit image is the concatenation either way): after every push the store
    equals the oracle's SeqVector of everything pushed so far, and extraction / minimizers see the grown sequence."""
    import kmers_b200 as kb
    import oracle as ko
    rng = np.random.default_rng(21)
    b = ctx.new_packed(16)
    total = b""
    for n in (1, 31, 32, 33, 5, 64, 1000, 3, 4097):
        piece = bytes(rng.choice(list(b"ACGTacgt"), size=n).astype(np.uint8))
        b.push_chars(piece)
        total += piece
        words = b.download().view(np.uint64)
        assert np.array_equal(words[: (len(total) + 31) // 32], ko.sv_from_bytes(total.upper()))
    k = 31
    res = b.extract_canonical(k, want_fw_rc=True, to="host")
    ref = ko.extract_canonical(np.frombuffer(total, dtype=np.uint8), k, n_reads=1, fixed_len=len(total), want_fw_rc=True)
    assert np.array_equal(res.canon, ref["canon"]) and np.array_equal(res.fw, ref["fw"]) and np.array_equal(res.hash, ref["hash"])
    mm, pos = b.minimizers(31, 15)
    rm, rp = ko.minimizers_batch(np.frombuffer(total, dtype=np.uint8), 31, 15, 15, n_reads=1, fixed_len=len(total))
    assert np.array_equal(mm, rm) and np.array_equal(pos, rp)
    with pytest.raises(kb.KmbPanic):  # Kmer::from panics on N: nothing is appended
        b.push_chars(b"ACGTNACGT")
    assert b.fixed_len == len(total) and np.array_equal(b.extract_canonical(k, to="host").canon, ref["canon"])
    with pytest.raises(kb.KmbError):  # a batch of several reads is not a SeqVector
        ctx.upload(b"ACGT" * 20, fixed_len=40).to_packed().push_chars(b"AC")


@pytest.mark.parametrize("ragged", [False, True])
def test_slice_is_a_view_of_one_read(ctx, ragged):
    """SeqVector::slice / SeqVectorSlice (seq_vector.rs:24-90): get_kmer_u64, iter_kmers and iter_minimizers of a slice equal
    those of the sub-sequence; unslice restores the batch."""
    import kmers_b200 as kb
    import oracle as ko
    rng = np.random.default_rng(8 + ragged)
    if ragged:
        lens = rng.integers(200, 700, size=12)
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        bases = np.frombuffer(bytes(rng.choice(list(b"ACGT"), size=int(offs[-1])).astype(np.uint8)), dtype=np.uint8)
        batch = ctx.upload(bases, offsets=offs).to_packed()
    else:
        lens = np.full(12, 333)
        offs = (np.arange(13) * 333).astype(np.uint64)
        bases = np.frombuffer(bytes(rng.choice(list(b"ACGT"), size=12 * 333).astype(np.uint8)), dtype=np.uint8)
        batch = ctx.upload(bases, fixed_len=333).to_packed()
    whole = batch.extract_canonical(31, to="host").canon.copy()
    for read, start, length in ((0, 0, int(lens[0])), (3, 17, 100), (7, 1, 31), (11, int(lens[11]) - 40, 40), (5, 50, 0), (2, 33, 30)):
        sub = bases[int(offs[read]) + start: int(offs[read]) + start + length]
        batch.slice(read, start, length)
        res = batch.extract_canonical(31, want_fw_rc=True, to="host")
        ref = ko.extract_canonical(sub, 31, n_reads=1, fixed_len=length, want_fw_rc=True)
        assert res.n_slots == max(0, length - 30)
        assert np.array_equal(res.fw, ref["fw"]) and np.array_equal(res.canon, ref["canon"]) and np.array_equal(res.hash, ref["hash"])
        if length >= 31:
            mm, pos = batch.minimizers(31, 15)
            rm, rp = ko.minimizers_batch(sub, 31, 15, 15, n_reads=1, fixed_len=length)
            assert np.array_equal(mm, rm) and np.array_equal(pos, rp)
        if length >= 9:
            p = np.arange(0, length - 8, 7, dtype=np.uint64)
            got = batch.get_kmers(9, p)
            want = [ko.sv_get_kmer_u64(ko.sv_from_bytes(sub.tobytes()), length, int(q), 9) for q in p]
            assert got.tolist() == want
    with pytest.raises(kb.KmbPanic):  # assert!(end <= self.len())
        batch.slice(1, 10, int(lens[1]))
    batch.unslice()
    assert np.array_equal(batch.extract_canonical(31, to="host").canon, whole)
