"""The small accessors of naive_impl::Kmer / CanonicalKmer and kmer::Kmer<P,K,B> in batched form, against the reference's
own unit-test vectors (naive_impl/kmer.rs:325-384 append / prepend, :530-542 sub_kmer; canonical_kmer.rs:272-297 shift /
swap; kmer.rs:156-203 get / get_prefix / bitmer_to_bytes) and against the restated scalar functions of the oracle."""
import ctypes as C

import numpy as np
import pytest

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


def w(ko, s):
    return ko.kmer_from(s).data


def test_append_prepend_goldens(ctx, ko):  # naive_impl/kmer.rs:325-384
    A, C_, G, T = 0, 1, 2, 3
    for src, base, want, off in (("att", "c", "ttc", A), ("ttcga", "g", "tcgag", T)):
        for ascii_ in (True, False):
            b = base.encode() if ascii_ else bytes([int(ko.lib().ko_encode_binary_u8(ord(base)))])
            out, dropped = ctx.append_base_words([w(ko, src)], b, len(src), ascii=ascii_)
            assert int(out[0]) == w(ko, want) and int(dropped[0]) == off
    for src, base, want, off in (("att", "c", "cat", T), ("ttcga", "g", "gttcg", A)):
        for ascii_ in (True, False):
            b = base.encode() if ascii_ else bytes([int(ko.lib().ko_encode_binary_u8(ord(base)))])
            out, dropped = ctx.prepend_base_words([w(ko, src)], b, len(src), ascii=ascii_)
            assert int(out[0]) == w(ko, want) and int(dropped[0]) == off


def test_sub_kmer_goldens(ctx, ko):  # naive_impl/kmer.rs:530-542
    import kmers_b200 as kb
    s = "ACTTGAT"
    km = w(ko, s)
    for i in range(len(s)):
        for j in range(i, len(s)):
            got = ctx.sub_kmer_words([km], len(s), i, j - i)
            assert int(got[0]) == (w(ko, s[i:j]) if j > i else 0)
    with pytest.raises(kb.KmbPanic):  # assert!(pos + width <= k)
        ctx.sub_kmer_words([km], 7, 3, 5)
    with pytest.raises(kb.KmbPanic):  # assert!(pos < k)
        ctx.sub_kmer_words([km], 7, 7, 0)


def test_canonical_shift_and_swap_goldens(ctx, ko):  # canonical_kmer.rs:272-297
    ck = ko.ck_from("acttg")
    fw, rc, _ = ctx.canonical_append_base_words([ck.fw.data], [ck.rc.data], b"a", 5, ascii=True)
    assert (int(fw[0]), int(rc[0])) == (w(ko, "cttga"), w(ko, "tcaag"))
    fw, rc, _ = ctx.canonical_prepend_base_words(fw, rc, b"c", 5, ascii=True)
    assert (int(fw[0]), int(rc[0])) == (w(ko, "ccttg"), w(ko, "caagg"))
    # test_equivalency: swap = exchanging the two arrays; then the fw word of caagt's twin is acttg itself
    ck2 = ko.ck_from("caagt")
    fw2, rc2 = np.array([ck2.rc.data], dtype=np.uint64), np.array([ck2.fw.data], dtype=np.uint64)  # swapped
    assert int(ctx.match_words([ck.fw.data], fw2, 5)[0]) == 1  # IdentityMatch
    fw3, _, _ = ctx.canonical_append_base_words(fw2, rc2, b"c", 5, ascii=True)
    assert int(ctx.match_words([ck.fw.data], fw3, 5)[0]) == 0  # NoMatch
    assert ctx.is_fw_canonical_words([ck.fw.data, ck.rc.data], [ck.rc.data, ck.fw.data]).tolist() == [int(ck.fw.data < ck.rc.data), int(ck.rc.data < ck.fw.data)]


@pytest.mark.parametrize("k", [1, 2, 15, 16, 17, 31, 32])
def test_shifts_match_the_restated_scalar_functions(ctx, ko, k):
    rng = np.random.default_rng(k)
    n = 5000
    mask = (1 << (2 * k)) - 1
    words = (rng.integers(0, 2**63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)) & np.uint64(mask)
    letters = np.frombuffer(bytes(rng.choice(list(b"ACGTacgtNn-\x00\xff"), size=n).astype(np.uint8)), dtype=np.uint8)
    codes = rng.integers(0, 4, size=n, dtype=np.uint8)
    L = ko.lib()
    for ascii_, bases in ((True, letters), (False, codes)):
        ga, gda = ctx.append_base_words(words, bases, k, ascii=ascii_)
        gp, gdp = ctx.prepend_base_words(words, bases, k, ascii=ascii_)
        for i in range(0, n, 37):
            for got, gd, fn_u8, fn in ((ga, gda, L.ko_kmer_append_base_u8, L.ko_kmer_append_base),
                                       (gp, gdp, L.ko_kmer_prepend_base_u8, L.ko_kmer_prepend_base)):
                km = ko.Kmer(k, int(words[i]))
                prepend = fn is L.ko_kmer_prepend_base
                if ascii_:
                    r = fn_u8(C.byref(km), int(bases[i]), 0) if prepend else fn_u8(C.byref(km), int(bases[i]))
                else:
                    r = fn(C.byref(km), int(bases[i]), 0) if prepend else fn(C.byref(km), int(bases[i]))
                assert int(got[i]) == km.data and int(gd[i]) == (r & 0xFF), (k, i, ascii_, prepend)
    # CanonicalKmer shifts keep fw / rc reverse complements of each other for valid bases
    rcw = ctx.reverse_complement_words(words, k)
    fw, rc, _ = ctx.canonical_append_base_words(words, rcw, codes, k)
    assert np.array_equal(ctx.reverse_complement_words(fw, k), rc)
    fw, rc, _ = ctx.canonical_prepend_base_words(words, rcw, codes, k)
    assert np.array_equal(ctx.reverse_complement_words(fw, k), rc)
    # sub_kmer at every (pos, width) of a few words
    for pos in range(0, k, max(1, k // 5)):
        for width in range(0, k - pos + 1, max(1, k // 4)):
            got = ctx.sub_kmer_words(words[:64], k, pos, width)
            for i in range(64):
                out = C.c_uint64()
                assert L.ko_sub_kmer_word(int(words[i]), k, pos, width, 0, C.byref(out)) == 0
                assert int(got[i]) == out.value


@pytest.mark.parametrize("word_bits,words_per_item", [(8, 4), (16, 2), (32, 1), (64, 1), (64, 2), (128, 1), (128, 2)])
def test_kmer_get_and_prefix_every_word_width(ctx, ko, word_bits, words_per_item):
    import kmers_b200 as kb
    rng = np.random.default_rng(word_bits + words_per_item)
    n = 300
    item_bytes = word_bits // 8 * words_per_item
    img = rng.integers(0, 256, size=n * item_bytes, dtype=np.uint8)
    L = ko.lib()
    for index in range(0, item_bytes * 4, max(1, item_bytes // 3)):
        got = ctx.kmer_get(word_bits, words_per_item, img, n, index)
        want = [L.ko_kmer_get(img[i * item_bytes:].ctypes.data, index) for i in range(n)]
        assert got.tolist() == want
    for length in range(0, min(word_bits, item_bytes * 8) // 2):
        if 2 * length + 1 > word_bits:
            break
        got = ctx.kmer_get_prefix(word_bits, words_per_item, img, n, length)
        nb = 2 * length + 1
        for i in range(0, n, 29):
            v = int.from_bytes(img[i * item_bytes:(i + 1) * item_bytes].tobytes(), "little") & ((1 << nb) - 1)
            assert int.from_bytes(got[i].tobytes(), "little") == v
            if nb <= 64:
                assert L.ko_kmer_get_prefix(img[i * item_bytes:].ctypes.data, length) == v
    with pytest.raises(kb.KmbPanic):
        ctx.kmer_get(word_bits, words_per_item, img, n, item_bytes * 4)  # one field beyond the array
    with pytest.raises(kb.KmbPanic):
        ctx.kmer_get_prefix(word_bits, words_per_item, img, n, word_bits // 2)  # 2 len + 1 bits > bits of P


def test_prefix_and_bitmer_goldens(ctx, ko):  # kmer.rs:186-203
    import kmers_b200 as kb
    img = ko.encode(ko.NAIVE["ACGT"], b"GTAC", 64, 1)
    pref = ctx.kmer_get_prefix(64, 1, img, 1, 4)
    assert int.from_bytes(pref[0].tobytes(), "little") == 0b01001110
    assert ctx.bitmer_to_bytes([0b01001110], 4).tobytes() == b"GTAC"
    rng = np.random.default_rng(9)
    mers = rng.integers(0, 2**62, size=1000, dtype=np.uint64)
    for length in (1, 7, 31, 32):
        got = ctx.bitmer_to_bytes(mers, length)
        buf = (C.c_uint8 * length)()
        for i in range(0, 1000, 41):
            ko.lib().ko_bitmer_to_bytes(int(mers[i]), length, buf)
            assert got[i].tobytes() == bytes(buf)


def test_encoding_module_accessors(ctx, ko):  # kmer.rs:186-203 through kmers_b200.encoding, the mirror of encoding::*
    from kmers_b200 import encoding as enc
    s = b"TAAGGATTCTAATCATAAGGATTCTAATCATAAGGATTCTAATCA"
    img = enc.encode(ctx, enc.Naive.ACGT, np.frombuffer(s, dtype=np.uint8).reshape(1, -1), 64)
    assert img.shape == (1, 16)
    letters = b"ACGT"
    assert bytes(letters[int(enc.get(ctx, img, 64, i)[0])] for i in range(len(s))) == s
    word0 = int.from_bytes(img[0, :8].tobytes(), "little")
    assert int.from_bytes(enc.get_prefix(ctx, img, 64, 5)[0].tobytes(), "little") == word0 & 0x7FF
    assert enc.bitmer_to_bytes(ctx, [word0], 32).tobytes() == s[:32]
