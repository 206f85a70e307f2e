"""Parity tests proper: the CUDA path through the C ABI vs the CPU oracle on
identical seeded inputs.  Everything here is integer work, so the bar is
bit-exact equality."""
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


@pytest.fixture(scope="module")
def ko():
    import oracle
    oracle.lib()
    return oracle


def _check_extract(res, ref, fwrc=False):
    assert res.n_slots == ref["n_slots"]
    assert np.array_equal(res.host("canon"), ref["canon"])
    assert np.array_equal(res.host("hash"), ref["hash"])
    if fwrc:
        assert np.array_equal(res.host("fw"), ref["fw"])
        assert np.array_equal(res.host("rc"), ref["rc"])
    if res.digest is not None:
        assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])


# ------------------------------------------------------------------ fixed-length batches
@pytest.mark.parametrize("k", [1, 2, 3, 8, 15, 16, 17, 24, 31, 32])
def test_fixed_len_all_k(ctx, ko, k):
    """BASELINE config-2 shape (150 bp) at oracle-friendly size, ~1 % invalid bytes, all K classes."""
    rng = np.random.default_rng(100 + k)
    n, L = 3000, 150
    bases, _ = random_reads(rng, n, L, L, p_bad=0.01)
    res = ctx.upload(bases, fixed_len=L).extract_canonical(k, want_fw_rc=True, digest=True)
    ref = ko.extract_canonical(bases, k, n_reads=n, fixed_len=L, want_fw_rc=True, n_threads=4)
    _check_extract(res, ref, fwrc=True)


@pytest.mark.parametrize("L,k", [(31, 31), (32, 31), (38, 31), (39, 31), (40, 31), (45, 31), (149, 31), (151, 31),
                                 (33, 32), (7, 3), (1, 1), (9, 2), (100, 17), (250, 21), (1000, 31), (10007, 31)])
def test_fixed_len_geometry(ctx, ko, L, k):
    """Read lengths around the run size: partial last runs, W odd (unaligned slot starts -> scalar store
    path), one run per read, reads longer than a CTA tile."""
    rng = np.random.default_rng(L * 131 + k)
    n = max(3, 60000 // L)
    bases, _ = random_reads(rng, n, L, L, p_bad=0.005)
    res = ctx.upload(bases, fixed_len=L).extract_canonical(k, digest=True)
    ref = ko.extract_canonical(bases, k, n_reads=n, fixed_len=L, n_threads=4)
    _check_extract(res, ref)


def test_reads_shorter_than_k(ctx, ko):
    bases, _ = random_reads(np.random.default_rng(1), 50, 20, 20)
    res = ctx.upload(bases, fixed_len=20).extract_canonical(31, digest=True)
    assert res.n_slots == 0 and res.digest == (0, 0, 0)


def test_empty_batch(ctx):
    res = ctx.upload(np.zeros(0, dtype=np.uint8), fixed_len=150, n_reads=0).extract_canonical(31, digest=True)
    assert res.n_slots == 0 and res.digest == (0, 0, 0)
    res = ctx.upload(np.zeros(0, dtype=np.uint8), offsets=np.zeros(1, dtype=np.uint64)).extract_canonical(31, digest=True)
    assert res.n_slots == 0 and res.digest == (0, 0, 0)


def test_all_invalid_and_all_valid(ctx, ko):
    import kmers_b200 as kb
    n, L, k = 500, 150, 31
    bad = np.full(n * L, ord("N"), dtype=np.uint8)
    res = ctx.upload(bad, fixed_len=L).extract_canonical(k, digest=True)
    assert (res.host("canon") == kb.SENTINEL).all() and (res.host("hash") == kb.SENTINEL).all()
    assert res.digest == (0, 0, 0)
    good = ko.generate_bases(9, 0, n * L)
    res = ctx.upload(good, fixed_len=L).extract_canonical(k, digest=True)
    assert res.digest[0] == n * (L - k + 1)
    _check_extract(res, ko.extract_canonical(good, k, n_reads=n, fixed_len=L))


def test_every_byte_value(ctx, ko):
    """Only the 8 bytes ACGTacgt are valid (naive_impl/mod.rs:40-50): put each of the 256 byte values in
    the middle of an otherwise valid read."""
    k, L = 5, 21
    base = ko.generate_bases(3, 0, L)
    reads = np.tile(base, (256, 1))
    reads[:, 10] = np.arange(256, dtype=np.uint8)
    flat = reads.reshape(-1)
    res = ctx.upload(flat, fixed_len=L).extract_canonical(k, digest=True)
    _check_extract(res, ko.extract_canonical(flat, k, n_reads=256, fixed_len=L))
    assert res.digest[0] == 8 * (L - k + 1) + 248 * (L - k + 1 - k)


def test_lower_case_is_valid(ctx, ko):
    """Soft-masked (lower-case) spans encode like upper case (naive_impl/mod.rs:44-47)."""
    up = ko.generate_bases(5, 0, 150 * 64)
    lo = up.copy()
    lo[1000:6000] |= 0x20
    a = ctx.upload(up, fixed_len=150).extract_canonical(31, to="host")
    b = ctx.upload(lo, fixed_len=150).extract_canonical(31, to="host")
    assert np.array_equal(a.canon, b.canon) and np.array_equal(a.hash, b.hash)


def test_no_validate_flag_matches_path_e(ctx, ko):
    """KMB_F_NO_VALIDATE = Path-E semantics: every window kept, byte -> (c>>1)&3."""
    rng = np.random.default_rng(17)
    bases, _ = random_reads(rng, 200, 80, 80, p_bad=0.05)
    res = ctx.upload(bases, fixed_len=80).extract_canonical(31, validate=False, digest=True)
    ref = ko.extract_canonical_wide(bases, 31, enc=ko.NAIVE["ACGT"], validate=False, n_reads=200, fixed_len=80)
    assert np.array_equal(res.host("canon"), ref["canon"][:, 0])
    assert np.array_equal(res.host("hash"), ref["hash"][:, 0])
    assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])


def test_host_and_device_outputs_agree(ctx, ko):
    bases = ko.generate_bases(11, 0, 150 * 1000, n_thresh20=2000)
    b = ctx.upload(bases, fixed_len=150)
    dev = b.extract_canonical(31, want_fw_rc=True, to="device")
    host = b.extract_canonical(31, want_fw_rc=True, to="host")
    for name in ("canon", "hash", "fw", "rc"):
        assert np.array_equal(dev.host(name), host.host(name))


def test_attach_device_memory_unaligned(ctx, ko):
    """Borrowed torch memory at every 16-byte misalignment (the tile loader aligns down and guards)."""
    import torch
    L, n, k = 150, 300, 31
    bases = ko.generate_bases(21, 0, n * L, n_thresh20=1500)
    ref = ko.extract_canonical(bases, k, n_reads=n, fixed_len=L)
    buf = torch.zeros(n * L + 64, dtype=torch.uint8, device="cuda")
    for shift in (0, 1, 7, 15, 16, 33):
        buf.zero_()
        view = buf[shift:shift + n * L]
        view.copy_(torch.from_numpy(bases).cuda())
        res = ctx.attach(view, fixed_len=L).extract_canonical(k, digest=True)
        _check_extract(res, ref)


def test_device_generator_matches_oracle(ctx, ko):
    for seed, n, L, thr, first in [(42, 1000, 150, 0, 0), (43, 77, 1000, 1049, 12345), (7, 1, 1, 0, 0), (8, 3, 17, 5000, 99)]:
        b = ctx.generate(seed, n, L, thr, first)
        assert np.array_equal(b.download(), ko.generate_bases(seed, first, n * L, thr))


# ------------------------------------------------------------------ ragged (CSR) batches
@pytest.mark.parametrize("k", [1, 5, 16, 31, 32])
def test_ragged_reads(ctx, ko, k):
    """Empty reads, reads shorter than K, long reads spanning several tiles, invalid bytes."""
    rng = np.random.default_rng(500 + k)
    bases, offs = random_reads(rng, 400, 0, 120, p_bad=0.02)
    b2, o2 = random_reads(rng, 3, 5000, 9000, p_bad=0.001)
    bases = np.concatenate([bases, b2])
    offs = np.concatenate([offs, o2[1:] + offs[-1]])
    batch = ctx.upload(bases, offsets=offs)
    res = batch.extract_canonical(k, want_fw_rc=True, digest=True)
    ref = ko.extract_canonical(bases, k, offsets=offs, want_fw_rc=True)
    _check_extract(res, ref, fwrc=True)
    lens = np.diff(offs.astype(np.int64))
    want_off = np.concatenate([[0], np.cumsum(np.maximum(0, lens - k + 1))]).astype(np.uint64)
    assert np.array_equal(batch.window_offsets(k), want_off)


def test_ragged_equals_fixed(ctx, ko):
    bases = ko.generate_bases(31, 0, 150 * 700, n_thresh20=3000)
    a = ctx.upload(bases, fixed_len=150).extract_canonical(31, digest=True, to="host")
    b = ctx.upload(bases, offsets=np.arange(701, dtype=np.uint64) * 150).extract_canonical(31, digest=True, to="host")
    assert np.array_equal(a.canon, b.canon) and np.array_equal(a.hash, b.hash) and a.digest == b.digest


def test_long_reads_with_invalid_runs(ctx, ko):
    """BASELINE config 4 shape: 10 kbp reads, ~0.1 % N, one N-run per read, a soft-masked span, IUPAC, newline."""
    n, L, k = 40, 10000, 31
    bases = ko.generate_bases(43, 0, n * L, n_thresh20=1049).reshape(n, L).copy()
    rng = np.random.default_rng(43)
    for r in range(n):
        s = int(rng.integers(0, L - 200))
        bases[r, s:s + 1 + int(rng.integers(0, 200))] = ord("N")
        if r % 5 == 0:
            t = int(rng.integers(0, L - 500))
            bases[r, t:t + 500] |= 0x20
        for ch in b"RYKM\n":
            bases[r, int(rng.integers(0, L))] = ch
    flat = bases.reshape(-1)
    res = ctx.upload(flat, fixed_len=L).extract_canonical(k, digest=True)
    _check_extract(res, ko.extract_canonical(flat, k, n_reads=n, fixed_len=L, n_threads=4))


# ------------------------------------------------------------------ fused histogram (config 5 shape)
@pytest.mark.parametrize("ragged", [False, True])
def test_histogram(ctx, ko, ragged):
    n, L, k, bits = 2000, 150, 31, 12
    bases = ko.generate_bases(44, 0, n * L, n_thresh20=500)
    offs = np.arange(n + 1, dtype=np.uint64) * L if ragged else None
    batch = ctx.upload(bases, offsets=offs) if ragged else ctx.upload(bases, fixed_len=L)
    hist, dig = batch.histogram(k, bits, to="host")
    ref = ko.extract_canonical(bases, k, n_reads=n, fixed_len=L, hist_bits=bits, n_threads=4)
    assert np.array_equal(hist, ref["hist"])
    assert dig == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])
    assert int(hist.sum()) == dig[0]
    hist2, _ = batch.histogram(k, bits, hist=hist.copy(), accumulate=True, to="host")
    assert np.array_equal(hist2, 2 * ref["hist"])


def test_single_sequence_sharded_with_halo(ctx, ko):
    """Config 5: one long sequence cut into contiguous chunks with a K-1 halo; a window belongs to the
    chunk holding its first base -> the chunk histograms / digests sum to the whole."""
    G_, k, bits, parts = 200_000, 31, 10, 4
    seq = ko.generate_bases(44, 0, G_, n_thresh20=105)
    whole = ko.extract_canonical(seq, k, n_reads=1, fixed_len=G_, hist_bits=bits)
    total = np.zeros(1 << bits, dtype=np.uint64)
    dig = [0, 0, 0]
    for p in range(parts):
        lo, hi = G_ * p // parts, G_ * (p + 1) // parts
        chunk = seq[lo:min(G_, hi + k - 1)]
        h, d = ctx.upload(chunk, fixed_len=chunk.size).histogram(k, bits, to="host")
        total += h
        dig = [(a + b) % 2**64 for a, b in zip(dig, d)]
    assert np.array_equal(total, whole["hist"])
    assert tuple(dig) == (whole["n_valid"], whole["checksum_canon"], whole["checksum_hash"])


# ------------------------------------------------------------------ pipelined host path
def _host_case(ko, n, L, k, thresh=300):
    bases = ko.generate_bases(42, 0, n * L, n_thresh20=thresh)
    ref = ko.extract_canonical(bases, k, n_reads=n, fixed_len=L, n_threads=8)
    return bases, ref, (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])


def _ref_any(ko, bases, k, n, L, validate):
    """Oracle result with or without validation (Path-E bytes go through the restated Encoding::encode, single-threaded)."""
    if validate:
        return ko.extract_canonical(bases, k, n_reads=n, fixed_len=L, n_threads=8)
    r = ko.extract_canonical_wide(bases, k, enc=ko.NAIVE["ACGT"], validate=False, n_reads=n, fixed_len=L)
    return dict(r, canon=r["canon"][:, 0].copy(), hash=r["hash"][:, 0].copy())


def test_extract_canonical_host_pipelined(ctx, ko):
    n, L, k = 150_000, 150, 31  # several chunks
    bases, ref, rdig = _host_case(ko, n, L, k)
    canon = np.empty(n * (L - k + 1), dtype=np.uint64)
    hsh = np.empty_like(canon)
    dig = ctx.extract_canonical_host(bases, n, L, k, host_canon=canon, host_hash=hsh)
    assert np.array_equal(canon, ref["canon"]) and np.array_equal(hsh, ref["hash"])
    assert dig == rdig
    assert ctx.extract_canonical_host(bases, n, L, k) == dig  # digest-only form
    st = ctx.host_stats()
    assert st["chunks"] >= 1 and st["raw_chunks"] == 0 and st["d2h_bytes"] == 0  # pageable input: every chunk is packed
    assert st["h2d_bytes"] < 0.45 * n * L  # 3 bits per base (+ padding) crossed the link, not 8


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("chunk_mb", [1, 16])
def test_host_pipeline_device_outputs_stay_resident(ctx, ko, monkeypatch, pinned, chunk_mb):
    """Device-output mode: the kernels write the caller's full-batch device arrays in place (nothing is overwritten by a
    later chunk), for pageable input (all chunks packed by the host workers) and pinned input (packed from the front,
    raw ASCII from the back), with one and with many chunks."""
    import torch
    monkeypatch.setenv("KMB_PIPE_CHUNK_MB", str(chunk_mb))
    n, L, k = 120_001, 150, 31
    bases, ref, rdig = _host_case(ko, n, L, k)
    src = bases
    if pinned:
        t = torch.empty(n * L, dtype=torch.uint8, pin_memory=True)
        src = t.numpy()
        src[:] = bases
    W = L - k + 1
    canon = torch.full((n * W,), -7, dtype=torch.int64, device="cuda")
    hsh = torch.full((n * W,), -7, dtype=torch.int64, device="cuda")
    dig = ctx.extract_canonical_host(src, n, L, k, out_canon=canon, out_hash=hsh)
    assert dig == rdig
    assert np.array_equal(canon.cpu().numpy().view(np.uint64), ref["canon"])
    assert np.array_equal(hsh.cpu().numpy().view(np.uint64), ref["hash"])
    st = ctx.host_stats()
    assert st["chunks"] == -(-n // max(16, ((chunk_mb << 20) // L) // 16 * 16))
    if not pinned:
        assert st["raw_chunks"] == 0
    # mixed: canonical words stay on the device, hashes come back to the host; and canonical words only
    canon.fill_(-7)
    h_host = np.empty(n * W, dtype=np.uint64)
    assert ctx.extract_canonical_host(src, n, L, k, out_canon=canon, out_hash=h_host) == rdig
    assert np.array_equal(canon.cpu().numpy().view(np.uint64), ref["canon"]) and np.array_equal(h_host, ref["hash"])
    canon.fill_(-7)
    assert ctx.extract_canonical_host(src, n, L, k, out_canon=canon) == rdig
    assert np.array_equal(canon.cpu().numpy().view(np.uint64), ref["canon"])


@pytest.mark.parametrize("mode", ["pack_only", "raw_only"])
def test_host_pipeline_single_paths(ctx, ko, monkeypatch, mode):
    """The two feeds on their own (the hybrid of the default run is any mixture of them): packed chunks only, raw ASCII only."""
    import torch
    monkeypatch.setenv("KMB_PIPE_CHUNK_MB", "1")
    monkeypatch.setenv("KMB_PIPE_RAW" if mode == "pack_only" else "KMB_PIPE_PACK", "0")
    if mode == "pack_only":
        monkeypatch.setenv("KMB_PIPE_PACK", "1")  # (the default turns packing of pinned input off on hosts with few cores)
    n, L, k = 40_000, 150, 31
    bases, ref, rdig = _host_case(ko, n, L, k, thresh=2000)
    t = torch.empty(n * L, dtype=torch.uint8, pin_memory=True)
    t.numpy()[:] = bases
    canon = np.empty(n * (L - k + 1), dtype=np.uint64)
    hsh = np.empty_like(canon)
    assert ctx.extract_canonical_host(t.numpy(), n, L, k, out_canon=canon, out_hash=hsh) == rdig
    assert np.array_equal(canon, ref["canon"]) and np.array_equal(hsh, ref["hash"])
    st = ctx.host_stats()
    assert st["raw_chunks"] == (0 if mode == "pack_only" else st["chunks"])
    assert st["d2h_bytes"] == 2 * 8 * n * (L - k + 1)


@pytest.mark.parametrize("L,k,n", [(150, 31, 1), (150, 31, 17), (31, 31, 1000), (37, 5, 3333), (1000, 32, 257), (10_000, 31, 50), (20, 31, 10)])
@pytest.mark.parametrize("threads", [1, 3])
def test_host_pipeline_geometries(ctx, ko, monkeypatch, L, k, n, threads):
    """Odd shapes through the packed staging format: chunks that end inside a 16-base word, reads shorter than k (no
    windows at all), one read, long reads; soft-masked and IUPAC bytes; NO_VALIDATE (Path-E bytes through the host packer)."""
    monkeypatch.setenv("KMB_PIPE_CHUNK_MB", "1")
    ctx.set_host_threads(threads)
    try:
        rng = np.random.default_rng(L * 7 + k)
        bases, _ = random_reads(rng, n, L, L, p_bad=0.01)
        for validate in (True, False):
            ref = _ref_any(ko, bases, k, n, L, validate)
            W = max(0, L - k + 1)
            canon = np.empty(n * W, dtype=np.uint64)
            hsh = np.empty_like(canon)
            dig = ctx.extract_canonical_host(bases, n, L, k, out_canon=canon, out_hash=hsh, validate=validate)
            assert dig == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])
            assert np.array_equal(canon, ref["canon"]) and np.array_equal(hsh, ref["hash"])
    finally:
        ctx.set_host_threads(0)


def test_host_pipeline_prepacked_input(ctx, ko, monkeypatch):
    """Reads the caller packed once with kmb_host_pack: bits + masks, and bits alone (SeqVector semantics: no invalid base)."""
    import kmers_b200 as kb
    monkeypatch.setenv("KMB_PIPE_CHUNK_MB", "1")
    n, L, k = 20_000, 150, 31
    bases, ref, rdig = _host_case(ko, n, L, k, thresh=1000)
    bits, inv = kb.host_pack(bases)
    canon = np.empty(n * (L - k + 1), dtype=np.uint64)
    hsh = np.empty_like(canon)
    assert ctx.extract_canonical_host_packed(bits, inv, n, L, k, out_canon=canon, out_hash=hsh) == rdig
    assert np.array_equal(canon, ref["canon"]) and np.array_equal(hsh, ref["hash"])
    st = ctx.host_stats()
    assert st["h2d_bytes"] == 6 * ((n * L + 15) // 16) or st["h2d_bytes"] <= 6 * ((n * L + 15) // 16) + 6 * st["chunks"]
    # without masks every window is emitted, bytes mapping by (c >> 1) & 3: the NO_VALIDATE result
    ref_nv = _ref_any(ko, bases, k, n, L, False)
    dig = ctx.extract_canonical_host_packed(bits, None, n, L, k, out_canon=canon, out_hash=hsh)
    assert dig == (ref_nv["n_valid"], ref_nv["checksum_canon"], ref_nv["checksum_hash"])
    assert np.array_equal(canon, ref_nv["canon"]) and np.array_equal(hsh, ref_nv["hash"])


def test_host_pipeline_repeated_calls_and_errors(ctx, ko):
    import kmers_b200 as kb
    n, L, k = 20_000, 150, 31
    bases, ref, rdig = _host_case(ko, n, L, k)
    for _ in range(3):  # ring buffers, events and the pool are reused across calls
        assert ctx.extract_canonical_host(bases, n, L, k) == rdig
    with pytest.raises(kb.KmbPanic):
        ctx.extract_canonical_host(bases, n, L, 33)
    with pytest.raises(kb.KmbError):
        ctx.extract_canonical_host(bases, n, 0, k)
    assert ctx.extract_canonical_host(bases, n, L, k) == rdig  # still usable after an error


# ------------------------------------------------------------------ wide extension (K <= 64)
@pytest.mark.parametrize("k", [1, 16, 31, 32, 33, 47, 48, 49, 63, 64])
def test_wide_fixed(ctx, ko, k):
    """BASELINE config 3 (K=63, two u64 words) and the other K classes; parity unpinned above 32
    (extension) except through the pinned encode + rev_comp primitives the oracle composes."""
    rng = np.random.default_rng(900 + k)
    n, L = 120, 150
    bases, _ = random_reads(rng, n, L, L, p_bad=0.004)
    res = ctx.upload(bases, fixed_len=L).extract_canonical_wide(k, digest=True)
    ref = ko.extract_canonical_wide(bases, k, n_reads=n, fixed_len=L)
    assert np.array_equal(res.host("canon"), ref["canon"])
    assert np.array_equal(res.host("hash"), ref["hash"])
    assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])


@pytest.mark.parametrize("enc_name", ["ACTG", "ACGT", "TAGC", "GTCA", "CGAT", "XOR10"])
def test_wide_encodings_and_ragged(ctx, ko, enc_name):
    import kmers_b200 as kb
    rng = np.random.default_rng(77)
    bases, offs = random_reads(rng, 60, 0, 200, p_bad=0.01)
    enc_gpu = kb.ENC_XOR10 if enc_name == "XOR10" else int(kb.Naive[enc_name])
    enc_cpu = ko.XOR10 if enc_name == "XOR10" else ko.NAIVE[enc_name]
    for k, validate in ((63, True), (40, False), (21, True)):
        res = ctx.upload(bases, offsets=offs).extract_canonical_wide(k, enc_gpu, digest=True, validate=validate)
        ref = ko.extract_canonical_wide(bases, k, enc=enc_cpu, validate=validate, offsets=offs)
        assert np.array_equal(res.host("canon"), ref["canon"])
        assert np.array_equal(res.host("hash"), ref["hash"])
        assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])


def test_wide_agrees_with_narrow(ctx, ko):
    bases = ko.generate_bases(3, 0, 150 * 400, n_thresh20=2000)
    b = ctx.upload(bases, fixed_len=150)
    n = b.extract_canonical(31, to="host")
    w = b.extract_canonical_wide(31, to="host")
    assert np.array_equal(w.host("canon")[:, 0], n.canon) and np.array_equal(w.host("hash")[:, 0], n.hash)


# ------------------------------------------------------------------ batched Encoding<P,B>
def _encs(ko):
    return list(ko.NAIVE.items()) + [("XOR10", ko.XOR10)]


@pytest.mark.parametrize("word_bits", [8, 16, 32, 64, 128])
def test_pack_unpack_revcomp_all_encodings(ctx, ko, word_bits):
    """Encoding::encode / decode / rev_comp::<K> for all 24 Naive variants + Xor10 (encoding/naive.rs:116-154)."""
    import kmers_b200 as kb
    rng = np.random.default_rng(word_bits)
    for name, enc_cpu in _encs(ko):
        enc_gpu = kb.ENC_XOR10 if name == "XOR10" else enc_cpu
        for k in (1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 128):
            nw = kb.word_for_k(word_bits, k)
            if nw * word_bits > 256:
                continue
            seqs = np.frombuffer(b"ACGTacgtNnUu", dtype=np.uint8)[rng.integers(0, 12, size=(5, k))]
            img = kb.encode(ctx, enc_gpu, seqs, word_bits)
            want = np.stack([ko.encode(enc_cpu, s.tobytes(), word_bits, nw) for s in seqs])
            assert np.array_equal(img, want), (name, k)
            dec = kb.decode(ctx, enc_gpu, img, word_bits)
            assert [d.tobytes() for d in dec] == [ko.decode(enc_cpu, w, word_bits) for w in want], (name, k)
            assert np.array_equal(kb.decode(ctx, enc_gpu, img, word_bits, length=k), dec[:, :k])
            if k >= 2:
                rc = kb.rev_comp(ctx, enc_gpu, k, img, word_bits)
                want_rc = np.stack([ko.rev_comp(enc_cpu, k, w, word_bits) for w in want])
                assert np.array_equal(rc, want_rc), (name, k)


@pytest.mark.parametrize("k,word_bits", [(1, 8), (3, 8), (31, 64), (33, 64), (100, 32), (20000, 64)])
def test_unpack_bulk_any_item_length(ctx, ko, k, word_bits):
    """Bulk decode of many items whose text length is not a multiple of 4 or 16 (staged through shared memory), into
    host memory and into a deliberately misaligned device buffer."""
    import torch
    import kmers_b200 as kb
    from kmers_b200.context import _ptr
    rng = np.random.default_rng(k)
    n = 3 if k > 1000 else 7001
    nw = kb.word_for_k(word_bits, k)
    seqs = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, k))]
    img = kb.encode(ctx, kb.ENC_ACGT, seqs, word_bits)
    assert np.array_equal(kb.decode(ctx, kb.ENC_ACGT, img, word_bits, length=k), seqs)
    dev_in = torch.from_numpy(img.reshape(-1).copy()).cuda()
    for shift in (1, 5, 16):
        buf = torch.zeros(n * k + 64, dtype=torch.uint8, device="cuda")
        ctx._ck(ctx._lib.kmb_unpack(ctx._h, kb.ENC_ACGT, word_bits, _ptr(dev_in), n, nw, k, buf.data_ptr() + shift))
        ctx.sync()
        got = buf.cpu().numpy()
        assert np.array_equal(got[shift:shift + n * k].reshape(n, k), seqs), shift
        assert not got[:shift].any() and not got[shift + n * k:].any()


def test_revcomp_preserves_bits_above_2k(ctx, ko):
    """encoding/naive.rs:138-154 touches fields 0..K-1 only."""
    import kmers_b200 as kb
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(64, 16), dtype=np.uint8)
    for k in (3, 20, 33, 50, 64):
        got = kb.rev_comp(ctx, kb.Naive.ACGT, k, img, 64)
        want = np.stack([ko.rev_comp(ko.NAIVE["ACGT"], k, w, 64) for w in img])
        assert np.array_equal(got, want)


def test_pack_ragged_reads(ctx, ko):
    import kmers_b200 as kb
    rng = np.random.default_rng(12)
    bases, offs = random_reads(rng, 100, 0, 100, p_bad=0.05)
    batch = ctx.upload(bases, offsets=offs)
    for wb in (8, 32, 64, 128):
        img, woff = batch.pack(int(kb.Naive.ACGT), wb)
        lens = np.diff(offs.astype(np.int64))
        nw = (lens + wb // 2 - 1) // (wb // 2)
        assert np.array_equal(woff, np.concatenate([[0], np.cumsum(nw)]).astype(np.uint64))
        for r in range(100):
            seq = bases[int(offs[r]):int(offs[r + 1])].tobytes()
            want = ko.encode(ko.NAIVE["ACGT"], seq, wb, int(nw[r]))
            got = img[int(woff[r]) * wb // 8:int(woff[r + 1]) * wb // 8]
            assert np.array_equal(got, want), (wb, r)


# ------------------------------------------------------------------ naive_impl::Kmer word ops
@pytest.mark.parametrize("k", [1, 3, 16, 31, 32])
def test_word_ops(ctx, ko, k):
    import ctypes as C
    L = ko.lib()
    rng = np.random.default_rng(k)
    mask = (1 << (2 * k)) - 1
    words = (rng.integers(0, 2**63, size=2000, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=2000, dtype=np.uint64)) & np.uint64(mask)
    rc = ctx.reverse_complement_words(words, k)
    canon, flag = ctx.canonical_words(words, k)
    lex = ctx.lexhash_words(words, k)
    others = words.copy()
    others[::3] = rc[::3]
    others[1::3] ^= np.uint64(1)
    match = ctx.match_words(words, others, k)
    for i in range(0, 2000, 7):
        w = int(words[i])
        km = ko.Kmer(k, w)
        assert int(rc[i]) == L.ko_reverse_complement_word(w, k)
        assert int(canon[i]) == L.ko_kmer_to_canonical(km).data and bool(flag[i]) == bool(L.ko_kmer_is_canonical(km))
        assert int(lex[i]) == L.ko_lexhash_word(w, k)
        ck = L.ko_ck_from_u64(w, k, 0)
        assert int(match[i]) == L.ko_ck_get_word_equivalency(C.byref(ck), int(others[i]))
    # quickcheck properties of naive_impl/kmer.rs:280-290
    assert np.array_equal(ctx.reverse_complement_words(rc, k), words)
    c2, f2 = ctx.canonical_words(canon, k)
    assert np.array_equal(c2, canon) and f2.all()


def test_words_to_strings(ctx, ko):
    """String::from(Kmer) in batch (naive_impl/kmer.rs:196-207): lower case, base 0 first; goldens of kmer.rs:434-448."""
    for k in (1, 3, 16, 31, 32):
        rng = np.random.default_rng(k)
        words = rng.integers(0, 2**63, size=500, dtype=np.uint64) & np.uint64((1 << (2 * k)) - 1 if k < 32 else 2**64 - 1)
        got = ctx.words_to_strings(words, k)
        for w, row in zip(words[:100].tolist(), got[:100]):
            assert row.tobytes().decode() == ko.kmer_str(ko.Kmer(k, int(w)))
    assert ctx.words_to_strings(np.array([0b010000, 0b100100], dtype=np.uint64), 3).tobytes() == b"aacacg"
    with pytest.raises(Exception):
        ctx.words_to_strings(np.zeros(1, dtype=np.uint64), 33)


def test_word_ops_on_device_tensors(ctx, ko):
    import torch
    words = torch.randint(0, 2**62, (5000,), dtype=torch.int64, device="cuda")
    rc = ctx.reverse_complement_words(words, 31)
    assert rc.is_cuda
    back = ctx.reverse_complement_words(rc, 31)
    assert torch.equal(back, words)


# ------------------------------------------------------------------ ragged-path corner cases (slot-space CSR geometry)
def test_ragged_window_less_stretch_forces_multiple_passes(ctx, ko):
    """A CTA's slots separated by > one tile (36.8 K bases) of reads too short to hold a window."""
    rng = np.random.default_rng(2024)
    k = 31
    lens = np.concatenate([np.full(40, 200), np.full(4000, 20), np.full(40, 200), np.full(3000, 30), np.full(10, 77)])
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(offs[-1]))].copy()
    bases[rng.integers(0, bases.size, size=200)] = ord("N")
    res = ctx.upload(bases, offsets=offs).extract_canonical(k, want_fw_rc=True, digest=True)
    _check_extract(res, ko.extract_canonical(bases, k, offsets=offs, want_fw_rc=True), fwrc=True)


def test_ragged_many_tiny_reads_uncached_tables(ctx, ko):
    """More reads per CTA than the shared-memory offset cache holds (1024): W_r in {0,1,2,3}."""
    rng = np.random.default_rng(77)
    k = 31
    lens = rng.integers(29, 34, size=30000)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    bases = np.frombuffer(b"ACGTacgt", dtype=np.uint8)[rng.integers(0, 8, size=int(offs[-1]))].copy()
    bases[rng.integers(0, bases.size, size=500)] = ord("N")
    res = ctx.upload(bases, offsets=offs).extract_canonical(k, digest=True)
    _check_extract(res, ko.extract_canonical(bases, k, offsets=offs))
    w = ctx.upload(bases, offsets=offs).extract_canonical_wide(k, digest=True)
    assert np.array_equal(w.host("canon")[:, 0], res.host("canon"))


@pytest.mark.parametrize("L,k", [(32, 31), (33, 31), (37, 31), (9, 5), (70, 63), (66, 63)])
def test_fixed_reads_with_fewer_than_eight_windows(ctx, ko, L, k):
    """W < 8: one work item covers several reads (window-by-window path), narrow and wide engines."""
    rng = np.random.default_rng(L * 7 + k)
    n = 5000
    bases, _ = random_reads(rng, n, L, L, p_bad=0.01)
    if k <= 32:
        res = ctx.upload(bases, fixed_len=L).extract_canonical(k, digest=True)
        _check_extract(res, ko.extract_canonical(bases, k, n_reads=n, fixed_len=L, n_threads=4))
    w = ctx.upload(bases, fixed_len=L).extract_canonical_wide(k, digest=True)
    ref = ko.extract_canonical_wide(bases[:300 * L], k, n_reads=300, fixed_len=L)
    assert np.array_equal(w.host("canon")[:ref["canon"].shape[0]], ref["canon"])
    assert np.array_equal(w.host("hash")[:ref["hash"].shape[0]], ref["hash"])


@pytest.mark.parametrize("W,k", [(9, 63), (13, 63), (14, 63), (15, 63), (16, 63), (17, 40), (23, 33), (88, 63), (89, 64), (14, 17)])
def test_wide_pair_shape_boundaries(ctx, ko, W, k):
    """Two-word engine, paired item shape (an item spans 14 slots): reads with fewer windows than that go window by
    window, W around 14..17 and odd W make nearly every item straddle a read boundary; invalid bases at read edges."""
    rng = np.random.default_rng(W * 100 + k)
    L, n = k + W - 1, 3000
    bases, _ = random_reads(rng, n, L, L, p_bad=0.002)
    b2 = bases.reshape(n, L)
    b2[::5, 0] = ord("N")
    b2[3::7, -1] = ord("n")
    res = ctx.upload(bases, fixed_len=L).extract_canonical_wide(k, digest=True)
    ref = ko.extract_canonical_wide(bases, k, n_reads=n, fixed_len=L)
    assert np.array_equal(res.host("canon"), ref["canon"])
    assert np.array_equal(res.host("hash"), ref["hash"])
    assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])
    # the same reads as a ragged batch, with a few window-less reads mixed in
    lens = np.full(n, L, dtype=np.int64)
    lens[::11] = k - 1
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    rb = bases[: int(offs[-1])]
    res = ctx.upload(rb, offsets=offs).extract_canonical_wide(k, digest=True)
    ref = ko.extract_canonical_wide(rb, k, offsets=offs)
    assert np.array_equal(res.host("canon"), ref["canon"]) and np.array_equal(res.host("hash"), ref["hash"])
    assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])


def test_straddling_items_with_invalid_bases(ctx, ko):
    """W = 130 (not a multiple of 8): every 16th item straddles a read boundary; N's at read edges."""
    rng = np.random.default_rng(5)
    n, L, k = 4000, 150, 21
    bases, _ = random_reads(rng, n, L, L, p_bad=0.0)
    b2 = bases.reshape(n, L)
    b2[::3, :2] = ord("N")
    b2[1::3, -3:] = ord("n")
    b2[2::7, 70] = ord("-")
    res = ctx.upload(bases, fixed_len=L).extract_canonical(k, want_fw_rc=True, digest=True)
    _check_extract(res, ko.extract_canonical(bases, k, n_reads=n, fixed_len=L, want_fw_rc=True, n_threads=4), fwrc=True)


def test_pack_tiled_kernel_lengths(ctx, ko):
    """Tiled pack kernel across read lengths incl. L < 16, padding-only groups, and (u8 / u16 words) regions that end
    inside a 32-bit word of the image, so that one stored word holds bytes of up to four reads."""
    import kmers_b200 as kb
    rng = np.random.default_rng(31)
    for L, wb in [(150, 64), (31, 64), (1, 32), (16, 32), (17, 32), (100, 128), (5, 128), (1000, 64), (40000, 64),
                  (150, 8), (150, 16), (1, 8), (2, 8), (3, 8), (4, 8), (5, 8), (7, 16), (9, 16), (33, 8), (101, 16), (40001, 8)]:
        n = max(2, 20000 // L)
        bases, _ = random_reads(rng, n, L, L, p_bad=0.05)
        img, _ = ctx.upload(bases, fixed_len=L).pack(int(kb.Naive.TGCA), wb)
        nw = kb.word_for_k(wb, L)
        want = np.concatenate([ko.encode(ko.NAIVE["TGCA"], bases[r * L:(r + 1) * L].tobytes(), wb, nw) for r in range(n)])
        assert np.array_equal(img, want), (L, wb)


@pytest.mark.parametrize("bits", [4, 12, 16, 20])
def test_histogram_without_digest_and_bin_counts(ctx, ko, bits):
    """Shared-memory bins (<= 16 bits) and global bins (> 16), with and without the digest; a counter that
    wraps its 16-bit shared half (one bin takes > 65535 windows) still ends exact."""
    n, L, k = 3000, 150, 31
    bases = np.full(n * L, ord("A"), dtype=np.uint8)  # every window is poly-A: one bin takes all 360 000 counts
    bases[: 500 * L] = ko.generate_bases(5, 0, 500 * L)
    ref = ko.extract_canonical(bases, k, n_reads=n, fixed_len=L, hist_bits=bits, n_threads=4, materialize=False)
    b = ctx.upload(bases, fixed_len=L)
    for dg in (False, True):
        hist, d = b.histogram(k, bits, digest=dg, to="host")
        assert np.array_equal(hist, ref["hist"]), (bits, dg)
        assert int(hist.max()) > 65535
