"""The CUDA path (through the C ABI) against the reference's own known-answer
vectors (tests/golden/reference_goldens.json, transcribed from the reference's
unit tests with file:line).  Mirrors the reference's test modules."""
import numpy as np
import pytest

from golden_util import kmer_word, load_goldens, words_from_image

pytestmark = pytest.mark.gpu
G = load_goldens()


@pytest.fixture(scope="module")
def ctx():
    import kmers_b200 as kb
    c = kb.Context(0)
    yield c
    c.close()


def _u64(xs):
    return np.array(xs, dtype=np.uint64)


def test_iterator_goldens(ctx):
    """canonical_kmer_iterator.rs:123-206: the n-th emitted k-mer and its `pos`."""
    import kmers_b200 as kb
    k = G["iterator"]["k"]
    for case in G["iterator"]["cases"]:
        read = case["read"].encode()
        res = ctx.upload(read, fixed_len=len(read)).extract_canonical(k, want_fw_rc=True, digest=True, to="host")
        emitted = np.flatnonzero(res.canon != kb.SENTINEL)  # iterator order == increasing pos
        pos = int(emitted[case["steps"]])
        assert pos == case["pos"], case["cite"]
        win = case["read"][pos:pos + k]
        fw = kmer_word(win)
        assert int(res.fw[pos]) == fw, case["cite"]
        # CanonicalKmer::from(&r[pos..pos+31]) == iterator's km : rc word = reverse complement of fw
        rc = int(ctx.reverse_complement_words(_u64([fw]), k)[0])
        assert int(res.rc[pos]) == rc and int(res.canon[pos]) == min(fw, rc), case["cite"]
        assert res.digest[0] == emitted.size
    ex = G["iterator"]["exhausted"]
    res = ctx.upload(ex["read"].encode(), fixed_len=len(ex["read"])).extract_canonical(k, digest=True, to="host")
    assert res.digest[0] == ex["n_kmers"] == res.n_slots


def test_iterator_n_positions(ctx):
    """N at index 35 -> positions [0..4, 36..]; 60 k-mers."""
    import kmers_b200 as kb
    read = G["iterator"]["cases"][4]["read"].encode()
    res = ctx.upload(read, fixed_len=len(read)).extract_canonical(31, to="host")
    pos = np.flatnonzero(res.canon != kb.SENTINEL).tolist()
    assert pos == [0, 1, 2, 3, 4] + list(range(36, len(read) - 31 + 1)) and len(pos) == 60
    assert (res.hash[res.canon == kb.SENTINEL] == kb.SENTINEL).all()


def test_rc_goldens(ctx):  # naive_impl/kmer.rs:387-424
    for s, rc in G["rc_pairs"]:
        assert int(ctx.reverse_complement_words(_u64([kmer_word(s)]), len(s))[0]) == kmer_word(rc), (s, rc)


def test_canon_goldens(ctx):  # naive_impl/kmer.rs:293-317
    for s, canon in G["canon_pairs"]:
        c, flag = ctx.canonical_words(_u64([kmer_word(s)]), len(s))
        assert int(c[0]) == kmer_word(canon) and bool(flag[0]) == (s == canon), (s, canon)
        # and through the extraction path: one read that is exactly this k-mer
        res = ctx.upload(s.encode(), fixed_len=len(s)).extract_canonical(len(s), to="host")
        assert res.n_slots == 1 and int(res.canon[0]) == kmer_word(canon)


def test_bin_repr_goldens(ctx):  # naive_impl/kmer.rs:434-448 via Kmer::from == window fw word
    for s, v in G["bin_repr"].items():
        res = ctx.upload(s.encode(), fixed_len=3).extract_canonical(3, want_fw_rc=True, to="host")
        assert int(res.fw[0]) == v


def test_lexhash_goldens(ctx):  # naive_impl/hash.rs:84-104
    for s, v in G["lexhash_k3"].items():
        assert int(ctx.lexhash_words(_u64([kmer_word(s)]), 3)[0]) == v


def test_equivalency_goldens(ctx):  # naive_impl/canonical_kmer.rs:244-259, 283-297
    import kmers_b200 as kb
    ck = G["canonical_kmer"]
    assert int(ctx.reverse_complement_words(_u64([kmer_word(ck["fw"])]), 5)[0]) == kmer_word(ck["rc"])
    name = {"none": kb.NO_MATCH, "identity": kb.IDENTITY_MATCH, "twin": kb.TWIN_MATCH}
    for a, b, m in ck["equivalency"]:
        assert int(ctx.match_words(_u64([kmer_word(a)]), _u64([kmer_word(b)]), 5)[0]) == name[m]


@pytest.mark.parametrize("key", ["naive_acgt", "xor10"])
def test_encoding_goldens(ctx, key):
    """encoding/naive.rs:297-445 and xor10.rs:159-302: packed words at five word widths, decode with
    padding, multi-word rev_comp."""
    import kmers_b200 as kb
    enc = kb.Naive.ACGT if key == "naive_acgt" else kb.Xor10
    for g in G[key]:
        seq, wb = g["seq"].encode(), g["word_bits"]
        k = len(seq)
        img = kb.encode(ctx, enc, np.frombuffer(seq, dtype=np.uint8).reshape(1, k), wb)
        assert img.shape == (1, kb.num_bytes(wb, k))
        assert words_from_image(img[0], wb) == [int(w) for w in g["words"]], g["cite"]
        cap = img.shape[1] * 4
        dec = kb.decode(ctx, enc, img, wb)  # padding positions decode as 'A' (code 00), SURVEY Q12
        assert dec[0].tobytes() == seq + b"A" * (cap - k), g["cite"]
        rc = kb.rev_comp(ctx, enc, k, img, wb)
        assert kb.decode(ctx, enc, rc, wb)[0].tobytes() == g["rc_decoded"].encode(), g["cite"]


def test_kmer_get_goldens(ctx):  # kmer.rs:167-184
    import kmers_b200 as kb
    for g in G["kmer_get"]:
        img = kb.encode(ctx, kb.Naive[g["enc"]], np.frombuffer(g["seq"].encode(), dtype=np.uint8).reshape(1, 4), 8)
        assert [(int(img[0, 0]) >> (2 * i)) & 3 for i in range(4)] == g["codes"]


def test_panics_become_error_codes(ctx):
    """naive_impl/kmer.rs:476-480 too_long: k = 33 panics in the reference -> KMB_ERR_PANIC here."""
    import kmers_b200 as kb
    b = ctx.upload(b"A" * 40, fixed_len=40)
    with pytest.raises(kb.KmbPanic):
        b.extract_canonical(33)
    with pytest.raises(kb.KmbPanic):
        b.extract_canonical(0)
    with pytest.raises(kb.KmbPanic):
        ctx.reverse_complement_words(_u64([0]), 33)
    b.extract_canonical(32)  # not_too_long
