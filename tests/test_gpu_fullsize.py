"""BASELINE.json configs at their FULL sizes, checked through size-independent properties (the oracle cannot
materialise 19 GB in seconds, so: digests from the multi-threaded oracle, sums of the materialised arrays,
idempotence, hash-of-canonical, strand symmetry, and bit-exact comparison of random chunks)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N_READS, L, K, SEED = 10_000_000, 150, 31, 42
W = L - K + 1


@pytest.fixture(scope="module")
def full():
    import torch
    import kmers_b200 as kb
    ctx = kb.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    batch = ctx.generate(SEED, N_READS, L)
    res = batch.extract_canonical(K, digest=True, to="device")
    torch.cuda.synchronize()
    yield ctx, batch, res
    ctx.close()


def _u64sum(t):
    return int(t.sum().item()) % 2**64


def test_config2_digest_matches_oracle_and_arrays(full):
    import oracle as ko
    ctx, batch, res = full
    assert res.n_slots == N_READS * W
    # checksum of checksums: device digest == sum of the materialised arrays == multi-threaded oracle digest
    assert res.digest == (N_READS * W, _u64sum(res.canon), _u64sum(res.hash))
    bases = ko.generate_bases(SEED, 0, N_READS * L)
    ref = ko.extract_canonical(bases, K, n_reads=N_READS, fixed_len=L, n_threads=os.cpu_count() or 1, materialize=False)
    assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])


def test_config2_random_chunks_bit_exact(full):
    import oracle as ko
    ctx, batch, res = full
    rng = np.random.default_rng(0)
    for r0 in [0, N_READS - 5000] + [int(x) for x in rng.integers(0, N_READS - 5000, size=10)]:
        bases = ko.generate_bases(SEED, r0 * L, 5000 * L)
        ref = ko.extract_canonical(bases, K, n_reads=5000, fixed_len=L)
        sl = slice(r0 * W, (r0 + 5000) * W)
        assert np.array_equal(res.canon[sl].cpu().numpy().view(np.uint64), ref["canon"])
        assert np.array_equal(res.hash[sl].cpu().numpy().view(np.uint64), ref["hash"])


def test_config2_idempotence_and_hash_of_canonical(full):
    """to_canonical(canonical) == canonical, is_canonical; LexHasher(canonical) == the hash array."""
    import torch
    ctx, batch, res = full
    c2, flag = ctx.canonical_words(res.canon, K)
    assert torch.equal(c2, res.canon) and bool(flag.all())
    assert torch.equal(ctx.lexhash_words(res.canon, K), res.hash)
    rc = ctx.reverse_complement_words(res.canon, K)
    # canonical <= its reverse complement as unsigned numbers (both < 2^62, so signed compare is fine)
    assert bool((res.canon <= rc).all())


def test_config2_strand_symmetry(full):
    """Reverse-complementing every read reverses each read's canonical / hash arrays."""
    import torch
    import kmers_b200 as kb
    ctx, batch, res = full
    want_c = res.canon.view(N_READS, W).flip(1).contiguous()
    want_h = res.hash.view(N_READS, W).flip(1).contiguous()
    bases = torch.from_numpy(batch.download()).cuda()
    lut = torch.zeros(256, dtype=torch.uint8, device="cuda")
    for a, b in zip(b"ACGT", b"TGCA"):
        lut[a] = b
    rc_reads = lut[bases.view(N_READS, L).flip(1).long()].contiguous().view(-1)
    del bases
    with kb.Context(0, stream=torch.cuda.current_stream().cuda_stream) as ctx2:
        r2 = ctx2.attach(rc_reads, fixed_len=L).extract_canonical(K, to="device")
        torch.cuda.synchronize()
        assert torch.equal(r2.canon.view(N_READS, W), want_c)
        assert torch.equal(r2.hash.view(N_READS, W), want_h)


def test_config3_k63_two_words_full_size():
    """K=63 over the same 10^7 reads: digest == sums of the arrays; random chunks vs the oracle extension."""
    import torch
    import kmers_b200 as kb
    import oracle as ko
    with kb.Context(0, stream=torch.cuda.current_stream().cuda_stream) as ctx:
        batch = ctx.generate(SEED, N_READS, L)
        res = batch.extract_canonical_wide(63, want_hash=False, digest=True, to="device")
        torch.cuda.synchronize()
        w63 = L - 63 + 1
        assert res.n_slots == N_READS * w63 and res.digest[0] == N_READS * w63
        assert res.digest[1] == _u64sum(res.canon)
        for r0 in (0, 4_321_000, N_READS - 300):
            bases = ko.generate_bases(SEED, r0 * L, 300 * L)
            ref = ko.extract_canonical_wide(bases, 63, n_reads=300, fixed_len=L, want_hash=False)
            got = res.canon[2 * r0 * w63:2 * (r0 + 300) * w63].cpu().numpy().view(np.uint64).reshape(-1, 2)
            assert np.array_equal(got, ref["canon"])


def test_config4_long_reads_full_size():
    """10^5 x 10 kbp with the full SURVEY 8(d) C4 mixture -- per read one run of N of 1..200 bases, ~0.1 % isolated N,
    soft-masked (lower-case, still valid) 500-base spans on 5 % of the reads, IUPAC letters and newlines: the WHOLE input's
    digest vs the multi-threaded oracle on the same bytes, sums of arrays, chunk compares."""
    import torch
    import kmers_b200 as kb
    import oracle as ko
    import bench  # the mixture recipe the driver-clocked config-4 row uses
    from kmers_b200.context import _ptr
    n, Lr = 100_000, 10_000
    with kb.Context(0, stream=torch.cuda.current_stream().cuda_stream) as ctx:
        flat = torch.empty(n * Lr, dtype=torch.uint8, device="cuda")
        ctx.generate(43, n, Lr, n_thresh20=1049)
        ctx._ck(ctx._lib.kmb_batch_download(ctx._h, _ptr(flat), n * Lr))
        bench.c4_mixture(torch, np, flat, n, Lr)
        bases = flat.cpu().numpy()
        assert (bases == ord("N")).sum() > 1.0e7 and ((bases >= 97) & (bases <= 122)).sum() > 2.0e6  # runs of N, lower case
        assert (bases == ord("\n")).sum() > 2000 and np.isin(bases, list(b"RYKM")).sum() > 2000
        batch = ctx.attach(flat, fixed_len=Lr)
        res = batch.extract_canonical(K, digest=True, to="device")
        torch.cuda.synchronize()
        valid = res.canon != -1
        assert res.digest[0] == int(valid.sum().item())
        assert res.digest[1] == _u64sum(torch.where(valid, res.canon, torch.zeros_like(res.canon)))
        ref = ko.extract_canonical(bases, K, n_reads=n, fixed_len=Lr, n_threads=os.cpu_count() or 1, materialize=False)
        assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])
        assert ref["n_valid"] < 0.97 * n * (Lr - K + 1)  # the N runs and isolated N invalidate > 3 % of the windows
        wr = Lr - K + 1
        for r0 in (0, 37 * 1000, 77_777, n - 20):
            ref = ko.extract_canonical(bases[r0 * Lr:(r0 + 20) * Lr], K, n_reads=20, fixed_len=Lr)
            assert np.array_equal(res.canon[r0 * wr:(r0 + 20) * wr].cpu().numpy().view(np.uint64), ref["canon"])
            assert np.array_equal(res.hash[r0 * wr:(r0 + 20) * wr].cpu().numpy().view(np.uint64), ref["hash"])
        # the same bytes through the iterator-identical stream: its length is the digest's count
        m = C.c_uint64()
        ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, K, 0, None, None, None, None, 0, C.byref(m)))
        assert int(m.value) == res.digest[0]


def test_more_than_2_pow_32_slots():
    """Maximum sizes: 3.6e7 reads x 150 bp = 4.32e9 output slots (> 2^32): 64-bit slot arithmetic end to end.
    Digest vs the multi-threaded oracle; the last reads bit-exact; sums of the arrays."""
    import torch
    import kmers_b200 as kb
    import oracle as ko
    n = 36_000_000
    assert n * W > 2**32
    with kb.Context(0, stream=torch.cuda.current_stream().cuda_stream) as ctx:
        batch = ctx.generate(SEED, n, L, n_thresh20=200)
        res = batch.extract_canonical(K, digest=True, to="device")
        torch.cuda.synchronize()
        assert res.n_slots == n * W
        tail = 2000
        bases_tail = ko.generate_bases(SEED, (n - tail) * L, tail * L, 200)
        ref = ko.extract_canonical(bases_tail, K, n_reads=tail, fixed_len=L)
        assert np.array_equal(res.canon[(n - tail) * W:].cpu().numpy().view(np.uint64), ref["canon"])
        assert np.array_equal(res.hash[(n - tail) * W:].cpu().numpy().view(np.uint64), ref["hash"])
        valid = res.canon != -1
        assert res.digest[0] == int(valid.sum().item())
        assert res.digest[1] == _u64sum(torch.where(valid, res.canon, torch.zeros_like(res.canon)))
        del valid
        # oracle digest over the whole input, in 4 chunks to bound host memory
        tot = [0, 0, 0]
        for c in range(4):
            r0, r1 = n * c // 4, n * (c + 1) // 4
            hb = ko.generate_bases(SEED, r0 * L, (r1 - r0) * L, 200)
            d = ko.extract_canonical(hb, K, n_reads=r1 - r0, fixed_len=L, n_threads=os.cpu_count() or 1, materialize=False)
            tot = [(a + b) % 2**64 for a, b in zip(tot, (d["n_valid"], d["checksum_canon"], d["checksum_hash"]))]
            del hb
        assert res.digest == tuple(tot)


def test_ragged_batch_with_more_than_2_pow_32_slots():
    """The ragged (CSR) geometry at maximum size: 4.6e7 reads of 100..160 bases cut from one generated stream, 4.6e9
    output slots (> 2^32).  Digest vs the multi-threaded oracle; the last reads bit-exact."""
    import torch
    import kmers_b200 as kb
    import oracle as ko
    n = 46_000_000
    rng = np.random.default_rng(9)
    lens = rng.integers(100, 161, size=n).astype(np.int64)
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    total = int(offs[-1])
    n_slots = int((lens - (K - 1)).sum())
    assert n_slots > 2**32
    with kb.Context(0, stream=torch.cuda.current_stream().cuda_stream) as ctx:
        flat = ctx.generate(SEED, 1, total, n_thresh20=200)          # one stream of `total` bases on the device ...
        bases_dev = torch.empty(total, dtype=torch.uint8, device="cuda")
        ctx._ck(ctx._lib.kmb_batch_download(ctx._h, bases_dev.data_ptr(), total))
        ctx.sync()
        batch = ctx.attach(bases_dev, dev_offsets=torch.from_numpy(offs).cuda())   # ... read as ragged reads
        out = kb.CanonicalKmers(k=K, n_slots=n_slots, canon=torch.empty(n_slots, dtype=torch.int64, device="cuda"), hash=None)
        res = batch.extract_canonical(K, digest=True, out=out)
        torch.cuda.synchronize()
        tail = 3000
        b0 = int(offs[n - tail])
        bases_tail = ko.generate_bases(SEED, b0, total - b0, 200)
        ref = ko.extract_canonical(bases_tail, K, offsets=(offs[n - tail:] - b0).astype(np.uint64))
        assert np.array_equal(out.canon[n_slots - ref["canon"].size:].cpu().numpy().view(np.uint64), ref["canon"])
        tot = [0, 0, 0]
        parts = 6
        for c in range(parts):
            r0, r1 = n * c // parts, n * (c + 1) // parts
            hb = ko.generate_bases(SEED, int(offs[r0]), int(offs[r1] - offs[r0]), 200)
            d = ko.extract_canonical(hb, K, offsets=(offs[r0:r1 + 1] - offs[r0]).astype(np.uint64), n_threads=os.cpu_count() or 1,
                                     materialize=False)
            tot = [(a + b) % 2**64 for a, b in zip(tot, (d["n_valid"], d["checksum_canon"], d["checksum_hash"]))]
            del hb
        assert res.digest == tuple(tot)


def test_histogram_exact_under_maximum_contention():
    """Shared-memory bins: every thread of every CTA hammers the two 16-bit counters of ONE 32-bit word (hist_bits = 2:
    poly-A / poly-T reads fall into bin 0, poly-G reads -- canonical poly-C -- into bin 1).  240 M increments; the
    totals must still be exact: no carry may leak from one counter into its neighbour."""
    import torch
    import kmers_b200 as kb
    n, k = 2_000_000, 31
    reads = torch.full((n, L), ord("A"), dtype=torch.uint8, device="cuda")
    reads[1::2] = ord("T")      # canonical poly-A again (reverse complement): same bin
    reads[2::5] = ord("G")      # canonical poly-C
    n_c = len(range(2, n, 5))
    poly_c = int("01" * k, 2)   # LexHash of CCC...C
    with kb.Context(0) as ctx:
        batch = ctx.attach(reads.reshape(-1), fixed_len=L, n_reads=n)
        for bits in (1, 2, 16, 20):
            hist, dig = batch.histogram(k, bits, to="host")
            assert dig[0] == n * (L - k + 1)
            want = np.zeros(1 << bits, dtype=np.uint64)
            want[0] = (n - n_c) * (L - k + 1)
            want[poly_c >> (2 * k - bits)] += n_c * (L - k + 1)
            assert np.array_equal(hist, want), bits
