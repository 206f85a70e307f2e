""""Next" row N1: minimizers on the GPU vs the restated reference algorithms (monotone deque of
naive_impl/seq_vector/minimizers.rs:38-142 and Kmer::minimizer_word, naive_impl/kmer.rs:170-191)."""
import numpy as np
import pytest

from golden_util import kmer_word, random_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import kmers_b200 as kb
    c = kb.Context(0)
    yield c
    c.close()


def test_reference_minimizer_goldens(ctx):
    """minimizers.rs:221-290: leftmost_mmer, mmers0, mmers1, mmers2 (k, w, LexHasherState(hash_k))."""
    aac, acc, aaa, aca = 0b010000, 0b010100, 0, 0b000100
    cases = [(b"AAAAAAA", 5, 3, 3, [(0, 0), (0, 1), (0, 2)]),
             (b"AAACAAA", 6, 3, 6, [(0, 0), (0, 4)]),
             (b"AACCAAA", 5, 3, 5, [(aac, 0), (acc, 1), (aaa, 4)]),
             (b"CACACACCAC", 7, 3, 3, [(aca, 1), (aca, 1), (aca, 3), (aca, 3)])]
    for seq, k, w, hk, want in cases:
        mm, pos = ctx.upload(seq, fixed_len=len(seq)).minimizers(k, w, hk)
        assert list(zip(mm.tolist(), pos.tolist())) == want, seq


@pytest.mark.parametrize("k,w,hk", [(31, 15, 15), (31, 21, 21), (31, 24, 32), (31, 31, 31), (31, 1, 1), (32, 17, 17), (21, 11, 11),
                                     (15, 12, 5), (9, 3, 3), (7, 7, 7), (31, 19, 10), (5, 2, 2), (32, 32, 32), (20, 13, 13),
                                     # the three width classes of the compare (w <= 13, <= 15, wider), short / long windows
                                     (31, 14, 14), (31, 15, 7), (16, 15, 15), (32, 13, 13), (32, 8, 8), (32, 1, 1), (20, 14, 3),
                                     (31, 16, 16), (22, 15, 15), (21, 14, 14), (32, 13, 2)])
def test_minimizers_fixed_vs_oracle(ctx, k, w, hk):
    import oracle as ko
    rng = np.random.default_rng(k * 100 + w)
    n, L = 400, 150
    bases, _ = random_reads(rng, n, L, L, p_bad=0.004)
    mm, pos = ctx.upload(bases, fixed_len=L).minimizers(k, w, hk)
    rmm, rpos = ko.minimizers_batch(bases, k, w, hk, n_reads=n, fixed_len=L)
    assert np.array_equal(mm, rmm) and np.array_equal(pos, rpos)


def test_minimizers_low_complexity_ties(ctx):
    """Homopolymers / short repeats: every tie must resolve to the leftmost lmer (minimizers.rs:72-78)."""
    import oracle as ko
    reads = [b"A" * 100, b"AC" * 50, b"ACG" * 34, b"T" * 40 + b"A" * 60, (b"ACGT" * 30)[:100], b"G" * 99 + b"N"]
    bases = np.frombuffer(b"".join(r[:100].ljust(100, b"A") for r in reads), dtype=np.uint8)
    for k, w in [(31, 15), (17, 5), (9, 9), (32, 8), (31, 13), (20, 14), (31, 16), (14, 13), (32, 2)]:
        mm, pos = ctx.upload(bases, fixed_len=100).minimizers(k, w)
        rmm, rpos = ko.minimizers_batch(bases, k, w, w, n_reads=len(reads), fixed_len=100)
        assert np.array_equal(mm, rmm) and np.array_equal(pos, rpos), (k, w)


def test_minimizers_ragged_and_geometry(ctx):
    import oracle as ko
    rng = np.random.default_rng(8)
    bases, offs = random_reads(rng, 500, 0, 140, p_bad=0.01)
    b2, o2 = random_reads(rng, 2, 20000, 30000, p_bad=0.001)
    bases = np.concatenate([bases, b2])
    offs = np.concatenate([offs, o2[1:] + offs[-1]])
    for k, w in [(31, 15), (25, 9), (13, 13)]:
        mm, pos = ctx.upload(bases, offsets=offs).minimizers(k, w)
        rmm, rpos = ko.minimizers_batch(bases, k, w, w, offsets=offs)
        assert np.array_equal(mm, rmm) and np.array_equal(pos, rpos), (k, w)
    for L, k, w in [(33, 31, 15), (150, 21, 7), (37, 31, 31)]:   # W < 8, straddling items
        b, _ = random_reads(rng, 600, L, L, p_bad=0.01)
        mm, pos = ctx.upload(b, fixed_len=L).minimizers(k, w)
        rmm, rpos = ko.minimizers_batch(b, k, w, w, n_reads=600, fixed_len=L)
        assert np.array_equal(mm, rmm) and np.array_equal(pos, rpos), (L, k, w)


def test_minimizer_words_vs_oracle(ctx):
    """Kmer::minimizer_word with the pinned hasher; naive_impl/kmer.rs:560-579 test_minimizer on 'ACTTGAT'."""
    import oracle as ko
    s = "ACTTGAT"
    for w in range(1, len(s)):
        mm, off = ctx.minimizer_words(np.array([kmer_word(s)], dtype=np.uint64), len(s), w)
        assert int(mm[0]) == kmer_word(s[int(off[0]):int(off[0]) + w])
        assert (int(mm[0]), int(off[0])) == ko.minimizer_word(kmer_word(s), len(s), w, w)
    rng = np.random.default_rng(3)
    for k, w, hk in [(31, 15, 15), (32, 32, 32), (31, 1, 7), (20, 13, 9)]:
        words = rng.integers(0, 2**63, size=3000, dtype=np.uint64) & np.uint64((1 << (2 * k)) - 1 if k < 32 else 2**64 - 1)
        mm, off = ctx.minimizer_words(words, k, w, hk)
        for i in range(0, 3000, 11):
            assert (int(mm[i]), int(off[i])) == ko.minimizer_word(int(words[i]), k, w, hk)


def test_minimizer_panics(ctx):
    import kmers_b200 as kb
    b = ctx.upload(b"ACGT" * 20, fixed_len=80)
    for k, w, hk in [(31, 32, 5), (33, 5, 5), (31, 0, 5), (31, 5, 33), (31, 5, 0)]:
        with pytest.raises(kb.KmbPanic):
            b.minimizers(k, w, hk)
