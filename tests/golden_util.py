"""Shared helpers for the golden-fixture tests (oracle side and CUDA side)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_goldens():
    with open(os.path.join(HERE, "golden", "reference_goldens.json")) as f:
        return json.load(f)


def kmer_word(s: str) -> int:
    """naive_impl::Kmer::from(&str) (naive_impl/kmer.rs:209-232): base i at bits 2i+1:2i, A0 C1 G2 T3."""
    code = {"a": 0, "c": 1, "g": 2, "t": 3}
    w = 0
    for i, ch in enumerate(s.lower()):
        w |= code[ch] << (2 * i)
    return w


def words_from_image(img: np.ndarray, word_bits: int):
    b = np.ascontiguousarray(img, dtype=np.uint8).tobytes()
    n = word_bits // 8
    return [int.from_bytes(b[i:i + n], "little") for i in range(0, len(b), n)]


def random_reads(rng, n, lo, hi, p_bad=0.02, lower=True):
    """Ragged reads over ACGT(acgt) with a sprinkling of bytes the reference treats as invalid."""
    alphabet = np.frombuffer(b"ACGTacgt" if lower else b"ACGT", dtype=np.uint8)
    bad = np.frombuffer(b"NnRYKM\n\x00\xff-Uu@[`{", dtype=np.uint8)
    lens = rng.integers(lo, hi + 1, size=n)
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    bases = alphabet[rng.integers(0, alphabet.size, size=int(offs[-1]))]
    m = rng.random(bases.size) < p_bad
    bases[m] = bad[rng.integers(0, bad.size, size=int(m.sum()))]
    return bases, offs
