"""The committed fixture tests/golden/reference_goldens.json (transcribed from
the reference's unit tests) against the CPU oracle."""
import ctypes as C

import numpy as np

import oracle as ko
from golden_util import kmer_word, load_goldens, words_from_image

G = load_goldens()
L = ko.lib()


def test_iterator_cases():
    k = G["iterator"]["k"]
    for case in G["iterator"]["cases"]:
        read = case["read"].encode()
        it = ko.Iter(read, k)
        it.inc_by(case["steps"])
        assert it.pos == case["pos"], case["cite"]
        want = ko.ck_from(read[case["pos"]:case["pos"] + k])
        assert (it.km.fw.data, it.km.rc.data) == (want.fw.data, want.rc.data), case["cite"]
    ex = G["iterator"]["exhausted"]
    it = ko.Iter(ex["read"].encode(), k)
    n = 0
    while not it.exhausted():
        n += 1
        it.inc()
    assert n == ex["n_kmers"]


def test_word_level_goldens():
    for s, rc in G["rc_pairs"]:
        assert L.ko_reverse_complement_word(kmer_word(s), len(s)) == kmer_word(rc)
    for s, canon in G["canon_pairs"]:
        assert L.ko_kmer_to_canonical(ko.kmer_from(s)).data == kmer_word(canon)
    for s, v in G["bin_repr"].items():
        assert ko.kmer_from(s).data == v == kmer_word(s)
    for s, v in G["lexhash_k3"].items():
        assert L.ko_lexhash_word(kmer_word(s), 3) == v
    ck = ko.ck_from(G["canonical_kmer"]["fw"])
    assert ck.rc.data == kmer_word(G["canonical_kmer"]["rc"])
    name = {"none": ko.NO_MATCH, "identity": ko.IDENTITY_MATCH, "twin": ko.TWIN_MATCH}
    for a, b, m in G["canonical_kmer"]["equivalency"]:
        assert L.ko_ck_get_word_equivalency(C.byref(ko.ck_from(a)), kmer_word(b)) == name[m]


def test_encoding_goldens():
    for key, enc in (("naive_acgt", ko.NAIVE["ACGT"]), ("xor10", ko.XOR10)):
        for g in G[key]:
            seq, wb = g["seq"].encode(), g["word_bits"]
            nw = L.ko_word_for_k(wb, len(seq))
            img = ko.encode(enc, seq, wb, nw)
            assert words_from_image(img, wb) == [int(w) for w in g["words"]], g["cite"]
            assert ko.decode(enc, ko.rev_comp(enc, len(seq), img, wb), wb) == g["rc_decoded"].encode(), g["cite"]
    for g in G["kmer_get"]:
        img = ko.encode(ko.NAIVE[g["enc"]], g["seq"].encode(), 8, 1)
        assert [L.ko_kmer_get(img.ctypes.data, i) for i in range(4)] == g["codes"]
