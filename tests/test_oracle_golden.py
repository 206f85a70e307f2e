"""Pins the CPU oracle against EVERY golden vector / known-answer test the
reference's own unit tests hold for the hot path.  Each test
names the reference test it transcribes (file:line under /root/reference/src).
The reference (Rust) cannot be built in this image; these vectors are the pin.
"""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

import oracle as ko

L = ko.lib()
ACGT = ko.NAIVE["ACGT"]
ACTG = ko.NAIVE["ACTG"]
TAGC = ko.NAIVE["TAGC"]

u64s = st.integers(min_value=0, max_value=2**64 - 1)

READ = (b"TTTTGGCCATTTTTCCTGTTCTTCAAGAAAACAGGAGATAACTAGAAGGACTAGAGAATGGGGCTGCCAGAACTAGTGGGAAGCTCCC"
        b"TAGAAATGGTGACATCGCCCACCAAACAGACC")
READ_N4 = (b"TTTTNGGCCATTTTTCCTGTTCTTCAAGAAAACAGGAGATAACTAGAAGGACTAGAGAATGGGGCTGCCAGAACTAGTGGGAAGCTCCC"
           b"TAGAAATGGTGACATCGCCCACCAAACAGACC")
READ_N35 = (b"TTTTGGCCATTTTTCCTGTTCTTCAAGAAAACAGGNAGATAACTAGAAGGACTAGAGAATGGGGCTGCCAGAACTAGTGGGAAGCTCCC"
            b"TAGAAATGGTGACATCGCCCACCAAACAGACC")


def eq(a: ko.Kmer, b: ko.Kmer) -> bool:  # PartialEq, naive_impl/kmer.rs:12-16
    return a.data == b.data and a.k == b.k


def ck_eq(a: ko.CanonicalKmer, b: ko.CanonicalKmer) -> bool:
    return eq(a.fw, b.fw) and eq(a.rc, b.rc)


# ---------------------------------------------------------------- naive_impl/mod.rs (prelude)
def test_encode_binary():  # naive_impl/kmer.rs:488-498 test_encode_binary
    for ch, code in zip("AaCcGgTt", [0, 0, 1, 1, 2, 2, 3, 3]):
        out = C.c_uint64()
        assert L.ko_encode_binary(ord(ch), C.byref(out)) == ko.OK
        assert out.value == code
        assert L.ko_encode_binary_u8(ord(ch)) == code


def test_encode_panics():  # naive_impl/kmer.rs:500-504 encode_panics
    out = C.c_uint64()
    assert L.ko_encode_binary(ord("N"), C.byref(out)) == ko.PANIC
    # naive_impl/mod.rs:48 : the u8 variant returns u64::MAX for anything else
    valid = set(b"ACGTacgt")
    for c in range(256):
        got = L.ko_encode_binary_u8(c)
        assert (got == ko.INVALID_BASE) == (c not in valid)


def test_complement_base():  # naive_impl/kmer.rs:506-512
    A, Cc, G, T = 0, 1, 2, 3
    assert L.ko_complement_base(A) == T
    assert L.ko_complement_base(T) == A
    assert L.ko_complement_base(Cc) == G
    assert L.ko_complement_base(G) == Cc


def test_is_valid_nuc():  # naive_impl/kmer.rs:514-528
    for b in (0, 1, 2, 3):
        assert L.ko_is_valid_nuc(b)
    assert not L.ko_is_valid_nuc(5)
    assert not L.ko_is_valid_nuc(3112)


def test_mask_table():  # naive_impl/kmer.rs:584-618 (incl. the k == 32 literal 0)
    for k in range(32):
        assert L.ko_mask_table(k, 1) == (1 << (2 * k)) - 1
    assert L.ko_mask_table(32, 1) == 0
    assert L.ko_mask_table(32, 0) == 2**64 - 1


# ---------------------------------------------------------------- naive_impl/kmer.rs
@settings(max_examples=300, deadline=None)
@given(u64s)
def test_rc_identity(word):  # naive_impl/kmer.rs:280-284 quickcheck rc_identity
    km = L.ko_kmer_from_u64(word, 31, 1)
    assert eq(km, L.ko_kmer_to_reverse_complement(L.ko_kmer_to_reverse_complement(km)))


@settings(max_examples=300, deadline=None)
@given(u64s)
def test_to_canonical_is_canonical(word):  # naive_impl/kmer.rs:286-290
    km = L.ko_kmer_from_u64(word, 31, 1)
    assert L.ko_kmer_is_canonical(L.ko_kmer_to_canonical(km))


def test_into_canon():  # naive_impl/kmer.rs:293-311
    s1, s2 = ko.kmer_from("taa"), ko.kmer_from("tta")
    assert eq(L.ko_kmer_to_canonical(s1), s1)
    assert eq(L.ko_kmer_to_canonical(s2), s1)
    s1, s2 = ko.kmer_from("atc"), ko.kmer_from("gat")
    assert eq(L.ko_kmer_to_canonical(s1), s1)
    assert eq(L.ko_kmer_to_canonical(s2), s1)
    not_canon = ko.kmer_from("gatacataggatgg")
    rc = L.ko_kmer_to_reverse_complement(ko.kmer_from("gatacataggatgg"))
    assert eq(rc, L.ko_kmer_to_canonical(not_canon))
    canon = ko.kmer_from("agatacataggatgg")
    assert eq(canon, L.ko_kmer_to_canonical(canon))


def test_is_canon():  # naive_impl/kmer.rs:313-317
    assert L.ko_kmer_is_canonical(ko.kmer_from("agatacataggatgg"))
    assert not L.ko_kmer_is_canonical(ko.kmer_from("gatacataggatgg"))


def test_ord():  # naive_impl/kmer.rs:319-322 : numeric, not lexicographic
    assert L.ko_kmer_cmp(ko.kmer_from("tcc"), ko.kmer_from("cct")) < 0


def test_append():  # naive_impl/kmer.rs:325-353
    A, T = 0, 3
    for use_u8 in (True, False):
        k1, k2 = ko.kmer_from("att"), ko.kmer_from("ttc")
        off = (L.ko_kmer_append_base_u8(C.byref(k1), ord("c")) if use_u8 else
               L.ko_kmer_append_base(C.byref(k1), L.ko_encode_binary_u8(ord("c"))))
        assert eq(k1, k2) and off == A
        k1, k2 = ko.kmer_from("ttcga"), ko.kmer_from("tcgag")
        off = (L.ko_kmer_append_base_u8(C.byref(k1), ord("g")) if use_u8 else
               L.ko_kmer_append_base(C.byref(k1), L.ko_encode_binary_u8(ord("g"))))
        assert eq(k1, k2) and off == T


def test_prepend():  # naive_impl/kmer.rs:355-384
    A, T = 0, 3
    for use_u8 in (True, False):
        k1, k2 = ko.kmer_from("att"), ko.kmer_from("cat")
        off = (L.ko_kmer_prepend_base_u8(C.byref(k1), ord("c"), 1) if use_u8 else
               L.ko_kmer_prepend_base(C.byref(k1), L.ko_encode_binary_u8(ord("c")), 1))
        assert eq(k1, k2) and off == T
        k1, k2 = ko.kmer_from("ttcga"), ko.kmer_from("gttcg")
        off = (L.ko_kmer_prepend_base_u8(C.byref(k1), ord("g"), 1) if use_u8 else
               L.ko_kmer_prepend_base(C.byref(k1), L.ko_encode_binary_u8(ord("g")), 1))
        assert eq(k1, k2) and off == A


@pytest.mark.parametrize("src,want", [  # naive_impl/kmer.rs:387-424 test_rc
    ("a", "t"), ("aaa", "ttt"), ("ttt", "aaa"), ("ta", "ta"), ("ccg", "cgg"), ("aat", "att"),
    ("gatacataggatgg", "ccatcctatgtatc"),
])
def test_rc(src, want):
    assert eq(L.ko_kmer_to_reverse_complement(ko.kmer_from(src)), ko.kmer_from(want))
    assert L.ko_reverse_complement_word(ko.kmer_from(src).data, len(src)) == ko.kmer_from(want).data


def test_rc_raw_struct():  # naive_impl/kmer.rs:388-391 : Kmer{k:1,data:0}.rc == "t"
    assert eq(L.ko_kmer_to_reverse_complement(ko.Kmer(1, 0)), ko.kmer_from("t"))


def test_str_repr():  # naive_impl/kmer.rs:426-431
    assert ko.kmer_str(ko.kmer_from("catagatacat")) == "catagatacat"


def test_bin_repr():  # naive_impl/kmer.rs:433-448
    assert ko.kmer_from("aaa").data == 0b000000
    assert ko.kmer_from("aac").data == 0b010000
    assert ko.kmer_from("acc").data == 0b010100
    assert ko.kmer_from("ccc").data == 0b010101


def test_aaa():  # naive_impl/kmer.rs:450-466
    x = ko.kmer_from("aaa")
    assert eq(x, L.ko_kmer_from_u64(0, 3, 1))
    assert x.data == 0 and x.k == 3
    for k in range(1, 33):
        x = ko.kmer_from("A" * k)
        assert x.data == 0 and x.k == k


def test_eq():  # naive_impl/kmer.rs:468-474
    assert eq(ko.kmer_from("aaa"), ko.kmer_from("AAA"))
    assert eq(ko.kmer_from("aCa"), ko.kmer_from("AcA"))
    assert not eq(ko.kmer_from("a"), ko.kmer_from("aa"))


def test_too_long():  # naive_impl/kmer.rs:476-485 too_long / not_too_long
    with pytest.raises(RuntimeError):
        ko.kmer_from("a" * 33)
    ko.kmer_from("a" * 32)


def test_sub_kmer():  # naive_impl/kmer.rs:530-542
    s = "ACTTGAT"
    km = ko.kmer_from(s)
    for i in range(len(s)):
        for j in range(i, len(s)):
            w = j - i
            out = C.c_uint64()
            assert L.ko_sub_kmer_word(km.data, km.k, i, w, 1, C.byref(out)) == ko.OK
            sub = L.ko_kmer_from_u64(out.value, w, 1)
            assert eq(ko.kmer_from(s[i:j]), sub)
    out = C.c_uint64()
    assert L.ko_sub_kmer_word(km.data, 7, 7, 0, 1, C.byref(out)) == ko.PANIC  # assert!(pos < k)
    assert L.ko_sub_kmer_word(km.data, 7, 3, 5, 1, C.byref(out)) == ko.PANIC  # assert!(pos+width <= k)


# ---------------------------------------------------------------- naive_impl/hash.rs
def test_lex_order():  # naive_impl/hash.rs:84-104
    h = lambda s: L.ko_lexhash_word(ko.kmer_from(s).data, 3)
    assert h("aaa") == 0
    assert h("aaa") < h("aac")
    assert h("aac") == 0b00001
    assert h("caa") < h("cac")
    assert h("caa") == 0b010000
    assert h("cac") == 0b010001


@settings(max_examples=200, deadline=None)
@given(st.text(alphabet="acgt", min_size=1, max_size=32), st.data())
def test_lexhash_is_lexicographic_rank(a, data):
    # what hash.rs:60-71 computes: string order == hash order for equal k
    b = data.draw(st.text(alphabet="acgt", min_size=len(a), max_size=len(a)))
    ha = L.ko_lexhash_word(ko.kmer_from(a).data, len(a))
    hb = L.ko_lexhash_word(ko.kmer_from(b).data, len(b))
    assert (a < b) == (ha < hb) and (a == b) == (ha == hb)


# ---------------------------------------------------------------- naive_impl/canonical_kmer.rs
@settings(max_examples=200, deadline=None)
@given(u64s)
def test_swap_identity(word):  # canonical_kmer.rs:216-223
    a = L.ko_ck_from_u64(word, 31, 1)
    b = L.ko_ck_from_u64(word, 31, 1)
    L.ko_ck_swap(C.byref(a))
    L.ko_ck_swap(C.byref(a))
    assert ck_eq(a, b)


@settings(max_examples=200, deadline=None)
@given(u64s)
def test_equivalency_property(word):  # canonical_kmer.rs:225-241
    canon_km = L.ko_ck_from_u64(word, 31, 1)
    canon_km2 = ko.CanonicalKmer(canon_km.rc, L.ko_kmer_to_reverse_complement(canon_km.rc))
    assert L.ko_ck_get_word_equivalency(C.byref(canon_km), canon_km2.fw.data) == ko.TWIN_MATCH
    L.ko_ck_swap(C.byref(canon_km2))
    assert L.ko_ck_get_word_equivalency(C.byref(canon_km), canon_km2.fw.data) == ko.IDENTITY_MATCH
    L.ko_ck_append_base_u8(C.byref(canon_km2), ord("c"), 1)
    e = L.ko_ck_get_word_equivalency(C.byref(canon_km), canon_km2.fw.data)
    # the reference asserts NoMatch; it holds unless the shifted word happens to
    # equal fw or rc again (homopolymer words), which quickcheck never drew
    fw2 = canon_km2.fw.data
    if fw2 != canon_km.fw.data and fw2 != canon_km.rc.data:
        assert e == ko.NO_MATCH


def test_from_u64_and_kmer():  # canonical_kmer.rs:244-259
    km = ko.kmer_from("acttg")
    ck = L.ko_ck_from_u64(km.data, km.k, 1)
    assert ko.kmer_str(ck.fw) == "acttg" and ko.kmer_str(ck.rc) == "caagt"
    ck = ko.ck_from("acttg")
    assert ko.kmer_str(ck.fw) == "acttg" and ko.kmer_str(ck.rc) == "caagt"


def test_swap():  # canonical_kmer.rs:261-269
    ck = ko.ck_from("acttg")
    L.ko_ck_swap(C.byref(ck))
    assert ko.kmer_str(ck.rc) == "acttg" and ko.kmer_str(ck.fw) == "caagt"


def test_shift():  # canonical_kmer.rs:271-280
    ck = ko.ck_from("acttg")
    L.ko_ck_append_base_u8(C.byref(ck), ord("a"), 1)
    assert ko.kmer_str(ck.fw) == "cttga" and ko.kmer_str(ck.rc) == "tcaag"
    L.ko_ck_prepend_base_u8(C.byref(ck), ord("c"), 1)
    assert ko.kmer_str(ck.rc) == "caagg" and ko.kmer_str(ck.fw) == "ccttg"


def test_equivalency():  # canonical_kmer.rs:282-297
    ck = ko.ck_from("acttg")
    ck2 = ko.ck_from("caagt")
    assert L.ko_ck_get_word_equivalency(C.byref(ck), ck2.fw.data) == ko.TWIN_MATCH
    L.ko_ck_swap(C.byref(ck2))
    assert L.ko_ck_get_word_equivalency(C.byref(ck), ck2.fw.data) == ko.IDENTITY_MATCH
    L.ko_ck_append_base_u8(C.byref(ck2), ord("c"), 1)
    assert L.ko_ck_get_word_equivalency(C.byref(ck), ck2.fw.data) == ko.NO_MATCH


def test_blank_of_size():  # canonical_kmer.rs:22-29
    ck = L.ko_ck_blank_of_size(31)
    assert ck.fw.data == 0 and ck.rc.data == 2**64 - 1 and ck.fw.k == 31 and ck.rc.k == 31


# ---------------------------------------------------------------- canonical_kmer_iterator.rs
def test_iter_init():  # canonical_kmer_iterator.rs:123-134
    it = ko.Iter(READ, 31)
    assert ck_eq(ko.ck_from(READ[0:31]), it.km) and it.pos == 0


def test_iter_inc():  # canonical_kmer_iterator.rs:137-148
    it = ko.Iter(READ, 31)
    it.inc()
    assert ck_eq(ko.ck_from(READ[1:32]), it.km) and it.pos == 1


def test_iter_inc_by():  # canonical_kmer_iterator.rs:151-162
    it = ko.Iter(READ, 31)
    it.inc_by(10)
    assert ck_eq(ko.ck_from(READ[10:41]), it.km) and it.pos == 10


def test_iter_init_invalid():  # canonical_kmer_iterator.rs:165-175 : N at index 4 -> first pos 5
    it = ko.Iter(READ_N4, 31)
    assert ck_eq(ko.ck_from(READ_N4[5:36]), it.km) and it.pos == 5


def test_iter_inc_by_invalid():  # canonical_kmer_iterator.rs:178-189 : N at 35, inc_by(5) -> 36
    it = ko.Iter(READ_N35, 31)
    it.inc_by(5)
    assert ck_eq(ko.ck_from(READ_N35[36:67]), it.km) and it.pos == 36


def test_exhausted_works():  # canonical_kmer_iterator.rs:192-206
    it = ko.Iter(READ, 31)
    it.inc_by(20)
    assert not it.exhausted()
    it.inc_by(len(READ) - 20)
    assert it.exhausted()
    it.inc()
    assert it.exhausted()


def test_iter_positions_with_n():  # positions [0..4, 36..], 60 k-mers, rolling == direct
    it = ko.Iter(READ_N35, 31)
    pos = []
    while not it.exhausted():
        pos.append(it.pos)
        assert ck_eq(ko.ck_from(READ_N35[it.pos:it.pos + 31]), it.km)
        it.inc()
    assert pos == [0, 1, 2, 3, 4] + list(range(36, len(READ_N35) - 31 + 1))
    assert len(pos) == 60


def test_iter_short_and_empty():  # canonical_kmer_iterator.rs:42-70 : no window -> exhausted at once
    assert ko.Iter(b"", 31).exhausted()
    assert ko.Iter(b"ACGT" * 7, 31).exhausted()  # 28 < 31
    it = ko.Iter(b"ACGT" * 8, 32 - 1)
    assert not it.exhausted() and it.pos == 0


# ---------------------------------------------------------------- encoding/naive.rs
def test_one_base_all_encoding():  # naive.rs:167-205
    for enc in ko.NAIVE.values():
        assert L.ko_nuc2bits(enc, ord("A")) == (enc >> 6) & 3
        assert L.ko_nuc2bits(enc, ord("C")) == (enc >> 4) & 3
        assert L.ko_nuc2bits(enc, ord("T")) == (enc >> 2) & 3
        assert L.ko_nuc2bits(enc, ord("G")) == enc & 3


def test_one_base_all_decoding():  # naive.rs:207-250
    for enc in ko.NAIVE.values():
        assert L.ko_bits2nuc(enc, (enc >> 6) & 3) == ord("A")
        assert L.ko_bits2nuc(enc, (enc >> 4) & 3) == ord("C")
        assert L.ko_bits2nuc(enc, (enc >> 2) & 3) == ord("T")
        assert L.ko_bits2nuc(enc, enc & 3) == ord("G")


def test_comp_one_base_all_encoding():  # naive.rs:252-294
    for enc in ko.NAIVE.values():
        n2b = lambda ch: L.ko_nuc2bits(enc, ord(ch))
        assert L.ko_complement_bits(enc, n2b("A")) == n2b("T")
        assert L.ko_complement_bits(enc, n2b("C")) == n2b("G")
        assert L.ko_complement_bits(enc, n2b("T")) == n2b("A")
        assert L.ko_complement_bits(enc, n2b("G")) == n2b("C")


def test_complement_is_constant_xor():  # complement == XOR with one 2-bit constant
    for enc in ko.NAIVE.values():
        xs = {b ^ L.ko_complement_bits(enc, b) for b in range(4)}
        assert len(xs) == 1 and xs.pop() in (1, 2, 3)


S15 = b"TAAGGATTCTAATCA"
T15 = [3, 0, 0, 2, 2, 0, 3, 3, 1, 3, 0, 0, 3, 1, 0]
X15 = [2, 0, 0, 3, 3, 0, 2, 2, 1, 2, 0, 0, 2, 1, 0]

NAIVE_GOLD = [  # (test name, file:line, seq, word_bits, n_words, packed words, rc decoded)
    ("k15pu8", "naive.rs:297-313", S15, 8, 4, [131, 242, 13, 7], b"TGATTAGAATCCTTAA"),
    ("k15pu16", "naive.rs:316-334", S15, 16, 2, [62083, 1805], b"TGATTAGAATCCTTAA"),
    ("k15pu32", "naive.rs:337-355", S15, 32, 1, [118354563], b"TGATTAGAATCCTTAA"),
    ("k30pu32", "naive.rs:358-385", S15 * 2, 32, 2, [3339580035, 29588640],
     b"TGATTAGAATCCTTATGATTAGAATCCTTAAA"),
    ("k45pu64", "naive.rs:388-416", S15 * 3, 64, 2, [3585846758293238403, 7397160],
     b"TGATTAGAATCCTTATGATTAGAATCCTTATGATTAGAATCCTTAAAAAAAAAAAAAAAAAAAA"),
    ("k65pu128", "naive.rs:419-445", S15 * 4 + b"GGGGG", 128, 2,
     [226115275135941975929349834069397860995, 2],
     b"CCCCCTGATTAGAATCCTTATGATTAGAATCCTTATGATTAGAATCCTTATGATTAGAATCCTTA" + b"A" * 63),
]


@pytest.mark.parametrize("name,cite,seq,wb,nw,packed,rc_dec", NAIVE_GOLD, ids=[g[0] for g in NAIVE_GOLD])
def test_naive_goldens(name, cite, seq, wb, nw, packed, rc_dec):
    k = len(seq)
    arr = ko.encode(ACGT, seq, wb, nw)
    assert ko.words(arr, wb) == packed
    table = (T15 * 5)[:k] if k != 65 else T15 * 4 + [2] * 5
    assert [L.ko_kmer_get(arr.ctypes.data, i) for i in range(k)] == table
    cap = nw * wb // 2
    assert ko.decode(ACGT, arr, wb) == seq + b"A" * (cap - k)  # padding decoded too (Q12)
    assert ko.decode(ACGT, ko.rev_comp(ACGT, k, arr, wb, strict=True), wb) == rc_dec


XOR10_GOLD = [  # commented-out "intended" tests, encoding/xor10.rs:159-302
    ("k15pu8", S15, 8, 4, [194, 163, 9, 6], b"TGATTAGAATCCTTAA"),
    ("k15pu16", S15, 16, 2, [41922, 1545], b"TGATTAGAATCCTTAA"),
    ("k15pu32", S15, 32, 1, [101295042], b"TGATTAGAATCCTTAA"),
    ("k30pu32", S15 * 2, 32, 2, [2248778690, 25323760], b"TGATTAGAATCCTTATGATTAGAATCCTTAAA"),
    ("k45pu64", S15 * 3, 64, 2, [2414607732474225602, 6330940],
     b"TGATTAGAATCCTTATGATTAGAATCCTTATGATTAGAATCCTTAAAAAAAAAAAAAAAAAAAA"),
    ("k65pu128", S15 * 4 + b"GGGGG", 128, 2, [339078536113543227067743297186703188930, 3],
     b"CCCCCTGATTAGAATCCTTATGATTAGAATCCTTATGATTAGAATCCTTATGATTAGAATCCTTA" + b"A" * 63),
]


@pytest.mark.parametrize("name,seq,wb,nw,packed,rc_dec", XOR10_GOLD, ids=[g[0] for g in XOR10_GOLD])
def test_xor10_goldens(name, seq, wb, nw, packed, rc_dec):
    k = len(seq)
    arr = ko.encode(ko.XOR10, seq, wb, nw)
    assert ko.words(arr, wb) == packed
    table = (X15 * 5)[:k] if k != 65 else X15 * 4 + [3] * 5
    assert [L.ko_kmer_get(arr.ctypes.data, i) for i in range(k)] == table
    cap = nw * wb // 2
    assert ko.decode(ko.XOR10, arr, wb) == seq + b"A" * (cap - k)
    # intended behaviour (strict=False): the swap loop for every B
    assert ko.decode(ko.XOR10, ko.rev_comp(ko.XOR10, k, arr, wb, strict=False), wb) == rc_dec
    # Xor10 and Naive::ACTG are the same encoding
    assert np.array_equal(arr, ko.encode(ACTG, seq, wb, nw))
    assert np.array_equal(ko.rev_comp(ko.XOR10, k, arr, wb), ko.rev_comp(ACTG, k, arr, wb))


def test_xor10_one_base():  # xor10.rs:116-157 (commented out)
    for ch, code in zip("ACTG", [0, 1, 2, 3]):
        assert L.ko_nuc2bits(ko.XOR10, ord(ch)) == code
        assert L.ko_bits2nuc(ko.XOR10, code) == ord(ch)
    n2b = lambda ch: L.ko_nuc2bits(ko.XOR10, ord(ch))
    for a, b in ("AT", "CG", "TA", "GC"):
        assert L.ko_complement_bits(ko.XOR10, n2b(a)) == n2b(b)


def test_xor10_single_word_quirk():  # xor10.rs:75-85, SURVEY Q2 -- unpinned, restated literally
    arr = ko.encode(ko.XOR10, b"ACGT", 64, 1)
    strict = ko.words(ko.rev_comp(ko.XOR10, 4, arr, 64, strict=True), 64)[0]
    w = ko.words(arr, 64)[0]
    r = 0
    for i in range(32):
        r |= ((w >> (2 * i)) & 3) << (2 * (31 - i))
    assert strict == (64 - 2 * r) % 2**64
    with pytest.raises(RuntimeError):  # impl only for u64/u128 (xor10.rs:50)
        ko.rev_comp(ko.XOR10, 4, ko.encode(ko.XOR10, b"ACGT", 32, 1), 32, strict=True)
    with pytest.raises(RuntimeError):  # u128 with high bits: to_u64().unwrap() panics (xor10.rs:76)
        ko.rev_comp(ko.XOR10, 64, ko.encode(ko.XOR10, b"G" * 64, 128, 1), 128, strict=True)


def test_rev_comp_k1_panics():  # naive.rs:140 `j -= 2` underflow, SURVEY Q9
    arr = ko.encode(ACGT, b"A", 8, 1)
    with pytest.raises(RuntimeError):
        ko.rev_comp(ACGT, 1, arr, 8, strict=True)


def test_encode_overflow_panics():  # naive.rs:120 set_bits range assert
    with pytest.raises(RuntimeError):
        ko.encode(ACGT, b"ACGTA", 8, 1)


def test_encode_ignores_validity():  # only bits 1-2 of the byte are looked at
    arr = ko.encode(ACGT, b"NnUuacgt", 16, 1)
    codes = [L.ko_kmer_get(arr.ctypes.data, i) for i in range(8)]
    g, t = L.ko_nuc2bits(ACGT, ord("G")), L.ko_nuc2bits(ACGT, ord("T"))
    assert codes == [g, g, t, t, 0, 1, 2, 3]


# ---------------------------------------------------------------- kmer.rs (generic Kmer<P,K,B>)
def test_choose_number_of_word():  # kmer.rs:97-118
    for wb, cases in {8: [(1, 1), (4, 1), (5, 2)], 16: [(1, 1), (8, 1), (9, 2)], 32: [(1, 1), (16, 1), (17, 2)],
                      64: [(1, 1), (32, 1), (64, 2)], 128: [(1, 1), (64, 1), (65, 2)]}.items():
        for k, want in cases:
            assert L.ko_word_for_k(wb, k) == want


def test_num_bytes():  # kmer.rs:120-153
    assert [L.ko_num_bytes(wb, 15) for wb in (8, 16, 32, 64, 128)] == [4, 4, 4, 8, 16]


def test_kmer_with_data():  # kmer.rs:155-165
    data = np.array([0b11100100], dtype=np.uint8)
    assert [L.ko_kmer_get(data.ctypes.data, i) for i in range(4)] == [0, 1, 2, 3]


def test_kmer_naive_encoder():  # kmer.rs:167-184
    a = ko.encode(ACTG, b"ACTG", 8, 1)
    assert [L.ko_kmer_get(a.ctypes.data, i) for i in range(4)] == [0, 1, 2, 3]
    a = ko.encode(TAGC, b"ACTG", 8, 1)
    assert [L.ko_kmer_get(a.ctypes.data, i) for i in range(4)] == [1, 3, 0, 2]


def test_kmer_prefix_and_to_bytes():  # kmer.rs:186-203
    a = ko.encode(ACGT, b"GTAC", 64, 1)
    pref = L.ko_kmer_get_prefix(a.ctypes.data, 4)
    assert pref == 0b01001110
    out = np.zeros(4, dtype=np.uint8)
    L.ko_bitmer_to_bytes(pref, 4, out.ctypes.data)
    assert out.tobytes() == b"GTAC"
    L.ko_bitmer_to_bytes(0b01001110, 4, out.ctypes.data)
    assert out.tobytes() == b"GTAC"


# ---------------------------------------------------------------- batch drivers vs the scalar restatement
def _rand_reads(rng, n, lo, hi, p_bad=0.02):
    alphabet = np.frombuffer(b"ACGTacgt", dtype=np.uint8)
    bad = np.frombuffer(b"NnRYKM\n\x00\xff-", dtype=np.uint8)
    lens = rng.integers(lo, hi + 1, size=n)
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    bases = alphabet[rng.integers(0, 8, size=int(offs[-1]))]
    m = rng.random(bases.size) < p_bad
    bases[m] = bad[rng.integers(0, bad.size, size=int(m.sum()))]
    return bases, offs


@pytest.mark.parametrize("k", [1, 2, 5, 15, 31])
def test_batch_matches_direct_definition(k):
    """Dense-slot driver == per-window direct construction (a window
    the iterator emits depends only on its own K bytes)."""
    rng = np.random.default_rng(7 + k)
    bases, offs = _rand_reads(rng, 60, 0, 90)
    res = ko.extract_canonical(bases, k, offsets=offs, want_fw_rc=True, hist_bits=min(2 * k, 8))
    valid = set(b"ACGTacgt")
    slot = 0
    nvalid = 0
    hist = np.zeros(1 << min(2 * k, 8), dtype=np.uint64)
    for r in range(60):
        seq = bases[int(offs[r]):int(offs[r + 1])].tobytes()
        for p in range(max(0, len(seq) - k + 1)):
            win = seq[p:p + k]
            if all(c in valid for c in win):
                ck = ko.ck_from(win)
                canon = min(ck.fw.data, ck.rc.data)
                h = L.ko_lexhash_word(canon, k)
                assert res["fw"][slot] == ck.fw.data and res["rc"][slot] == ck.rc.data
                assert res["canon"][slot] == canon and res["hash"][slot] == h
                hist[h >> (2 * k - min(2 * k, 8))] += 1
                nvalid += 1
            else:
                assert res["canon"][slot] == ko.SENTINEL and res["hash"][slot] == ko.SENTINEL
            slot += 1
    assert slot == res["n_slots"] and nvalid == res["n_valid"]
    assert np.array_equal(hist, res["hist"])
    ok = res["canon"] != ko.SENTINEL
    assert int(res["canon"][ok].sum(dtype=np.uint64)) == res["checksum_canon"]
    assert int(res["hash"][ok].sum(dtype=np.uint64)) == res["checksum_hash"]


def test_batch_threads_and_fixed_len_agree():
    bases = ko.generate_bases(42, 0, 150 * 500, n_thresh20=3000)
    a = ko.extract_canonical(bases, 31, n_reads=500, fixed_len=150, hist_bits=10)
    b = ko.extract_canonical(bases, 31, n_reads=500, fixed_len=150, hist_bits=10, n_threads=4)
    offs = np.arange(501, dtype=np.uint64) * 150
    c = ko.extract_canonical(bases, 31, offsets=offs, n_threads=3, hist_bits=10)
    for other in (b, c):
        assert np.array_equal(a["canon"], other["canon"]) and np.array_equal(a["hash"], other["hash"])
        assert np.array_equal(a["hist"], other["hist"])
        assert (a["n_valid"], a["checksum_canon"], a["checksum_hash"]) == (
            other["n_valid"], other["checksum_canon"], other["checksum_hash"])
    assert a["n_slots"] == 500 * 120 and 0 < a["n_valid"] < a["n_slots"]


def test_bench_faithful_equals_iterator_on_acgt():
    """benches/simple_benchmark.rs:14-44 per-window path == rolling iterator path on pure ACGT."""
    bases = ko.generate_bases(1, 0, 1 << 12)
    a = ko.bench_windows(bases, 31, n_reads=1, fixed_len=1 << 12)
    b = ko.extract_canonical(bases, 31, n_reads=1, fixed_len=1 << 12)
    assert np.array_equal(a["canon"], b["canon"]) and np.array_equal(a["hash"], b["hash"])
    assert a["checksum_canon"] == b["checksum_canon"] and a["n_valid"] == (1 << 12) - 30
    with pytest.raises(RuntimeError):  # encode_binary panics on N (naive_impl/mod.rs:35)
        ko.bench_windows(np.frombuffer(b"ACGTN" * 10, dtype=np.uint8), 3, n_reads=1, fixed_len=50)


def test_k32_mask_quirk():  # strict k=32 -> rc always 0 -> canonical word 0
    bases = ko.generate_bases(5, 0, 100)
    s = ko.extract_canonical(bases, 32, n_reads=1, fixed_len=100, strict=True)
    assert (s["canon"] == 0).all()
    i = ko.extract_canonical(bases, 32, n_reads=1, fixed_len=100, strict=False, want_fw_rc=True)
    for p in (0, 17, 68):
        ck = ko.ck_from(bases[p:p + 32].tobytes())  # From<&[u8]> + to_reverse_complement are fine at 32
        assert i["fw"][p] == ck.fw.data and i["rc"][p] == ck.rc.data


def test_wide_extension_agrees_with_path_n_for_k_le_32():
    rng = np.random.default_rng(3)
    bases, offs = _rand_reads(rng, 20, 0, 80)
    for k in (1, 7, 31, 32):
        n = ko.extract_canonical(bases, k, offsets=offs)
        w = ko.extract_canonical_wide(bases, k, offsets=offs, enc=ACGT, validate=True)
        ok = n["canon"] != ko.SENTINEL
        assert np.array_equal(w["canon"][:, 0], n["canon"])
        assert (w["canon"][ok, 1] == 0).all() and (w["canon"][~ok, 1] == ko.SENTINEL).all()
        assert np.array_equal(w["hash"][:, 0], n["hash"])
        assert w["n_valid"] == n["n_valid"] and w["checksum_canon"] == n["checksum_canon"]


def test_wide_k63_words_follow_pinned_encode_revcomp():
    seq = S15 * 4 + b"GGG"  # 63 bases
    w = ko.extract_canonical_wide(np.frombuffer(seq, dtype=np.uint8), 63, n_reads=1, fixed_len=63)
    fw = ko.encode(ACGT, seq, 64, 2)
    rc = ko.rev_comp(ACGT, 63, fw, 64)
    f, c = ko.words(fw, 64), ko.words(rc, 64)
    want = f if (f[1], f[0]) < (c[1], c[0]) else c
    assert [int(x) for x in w["canon"][0]] == want


def test_generator_is_counter_based():
    a = ko.generate_bases(42, 0, 1000)
    b = ko.generate_bases(42, 300, 200)
    assert np.array_equal(a[300:500], b)
    assert set(a.tobytes()) <= set(b"ACGT")
    n = ko.generate_bases(43, 0, 200000, n_thresh20=1049)
    frac = (n == ord("N")).mean()
    assert 0.0005 < frac < 0.0015
    # python restatement of splitmix64
    def sm(x):
        z = (x + 0x9E3779B97F4A7C15) % 2**64
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) % 2**64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) % 2**64
        return z ^ (z >> 31)
    assert all(L.ko_splitmix64(x) == sm(x) for x in (0, 1, 42, 2**63, 2**64 - 1))
    assert a[:8].tobytes() == bytes(b"ACGT"[sm(42 + i) >> 62] for i in range(8))


# ---------------------------------------------------------------- "next" rows: minimizers + SeqVector
def test_seq_vector_slice():  # naive_impl/seq_vector.rs:304-321 seq_slice_test
    words = np.array([1, 2, 3], dtype=np.uint64)
    assert ko.sv_get_kmer_u64(words, 96, 0, 32) == 1
    # slice(1, 96).get_kmer_u64(0, 32) == sv.get_kmer_u64(1, 32): bits 2..66 -> (1 >> 2) | (2 << 62)
    assert ko.sv_get_kmer_u64(words, 96, 1, 32) == ((1 >> 2) | (2 << 62)) % 2**64
    assert ko.sv_get_kmer_u64(words, 96, 75, 7) == (3 >> 22) & ((1 << 14) - 1)
    with pytest.raises(RuntimeError):
        ko.sv_get_kmer_u64(words, 96, 96, 1)  # assert!(pos < self.len())


def test_seq_vector_iter_kmers():  # seq_vector.rs:342-358 iter_kmers
    s = b"ACTTGAT"
    words = ko.sv_from_bytes(s)
    mers = ["act", "ctt", "ttg", "tga", "gat"]
    got = [ko.kmer_str(ko.Kmer(3, ko.sv_get_kmer_u64(words, 7, i, 3))) for i in range(5)]
    assert got == mers
    assert words.tolist() == [ko.kmer_from(s).data]
    long = b"A" * 30 + b"C" * 40  # seq_vector.rs:328-339 push_chars: 30 A then 40 C
    w2 = ko.sv_from_bytes(long)
    assert "".join("acgt"[ko.sv_get_kmer_u64(w2, 70, i, 1)] for i in range(70)).upper().encode() == long


def test_minimizer_leftmost():  # minimizers.rs:221-236 leftmost_mmer (any hasher: all lmers equal)
    assert ko.sv_minimizers(b"AAAAAAA", 5, 3, 3) == [(0, 0), (0, 1), (0, 2)]


def test_minimizer_mmers0():  # minimizers.rs:238-249 : LexHasherState::new(6), k=6, w=3
    assert ko.sv_minimizers(b"AAACAAA", 6, 3, 6) == [(0, 0), (0, 4)]


def test_minimizer_mmers1():  # minimizers.rs:251-270 : LexHasherState::new(5), k=5, w=3
    aac, acc, aaa = 0b010000, 0b010100, 0b000000
    assert ko.sv_minimizers(b"AACCAAA", 5, 3, 5) == [(aac, 0), (acc, 1), (aaa, 4)]


def test_minimizer_mmers2():  # minimizers.rs:272-290 : LexHasherState::new(3), k=7, w=3
    aca = 0b000100
    assert ko.sv_minimizers(b"CACACACCAC", 7, 3, 3) == [(aca, 1), (aca, 1), (aca, 3), (aca, 3)]


def test_minimizer_word_matches_definition():  # naive_impl/kmer.rs:560-579 test_minimizer, with the pinned hasher
    s = "ACTTGAT"
    km = ko.kmer_from(s)
    for w in range(1, len(s)):
        mm, o = ko.minimizer_word(km.data, km.k, w, w)
        h_min = L.ko_lexhash_word(mm, w)
        for i in range(len(s) - w + 1):
            out = C.c_uint64()
            assert L.ko_sub_kmer_word(km.data, km.k, i, w, 0, C.byref(out)) == ko.OK
            assert h_min <= L.ko_lexhash_word(out.value, w)
        assert ko.kmer_from(s[o:o + w]).data == mm
    with pytest.raises(RuntimeError):
        ko.minimizer_word(km.data, 7, 8, 3)


@settings(max_examples=60, deadline=None)
@given(st.text(alphabet="ACGT", min_size=12, max_size=80), st.integers(5, 12), st.data())
def test_minimizer_deque_equals_bruteforce(seq, k, data):
    """The restated deque iterator == leftmost argmin of the lmer hashes over every k-mer (what it computes)."""
    w = data.draw(st.integers(1, k))
    hk = data.draw(st.integers(w, 32))
    got = ko.sv_minimizers(seq.encode(), k, w, hk)
    words = ko.sv_from_bytes(seq.encode())
    for i, (word, pos) in enumerate(got):
        cands = [(L.ko_lexhash_word(ko.sv_get_kmer_u64(words, len(seq), p, w), hk), p) for p in range(i, i + k - w + 1)]
        best = min(cands)  # ties -> smallest p
        assert pos == best[1] and word == ko.sv_get_kmer_u64(words, len(seq), pos, w)


def test_minimizers_batch_with_invalid_bases():
    seq = np.frombuffer(b"AACCAAANAACCAAAACGTNNACGTTTT", dtype=np.uint8)
    mm, pos = ko.minimizers_batch(seq, 5, 3, 5, n_reads=1, fixed_len=seq.size)
    assert mm[:3].tolist() == [0b010000, 0b010100, 0] and pos[:3].tolist() == [0, 1, 4]
    assert (mm[3:8] == ko.SENTINEL).all() and (pos[3:8] == 2**32 - 1).all()
    assert pos[8] == 8 and mm[8] == 0b010000
