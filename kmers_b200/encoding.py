"""Mirror of `src/encoding` of the reference: the 24-way `Naive` enum
(encoding/naive.rs:49-74; the discriminant byte IS the code table: bits 7-6,
5-4, 3-2, 1-0 hold the 2-bit codes of A, C, T, G) and `Xor10`
(encoding/xor10.rs:12, identical to Naive::ACTG).  The trait methods
`encode / decode / rev_comp::<K>` (encoding/mod.rs:14-23) are exposed in
batched form over a `Context` / `ReadBatch`."""
from __future__ import annotations

import enum

import numpy as np

from . import _native as nv


class Naive(enum.IntEnum):
    ACTG = 0b00_01_10_11
    ACGT = 0b00_01_11_10
    ATCG = 0b00_10_01_11
    ATGC = 0b00_11_01_10
    AGCT = 0b00_10_11_01
    AGTC = 0b00_11_10_01
    CATG = 0b01_00_10_11
    CAGT = 0b01_00_11_10
    CTAG = 0b10_00_01_11
    CTGA = 0b11_00_01_10
    CGAT = 0b10_00_11_01
    CGTA = 0b11_00_10_01
    TACG = 0b01_10_00_11
    TAGC = 0b01_11_00_10
    TCAG = 0b10_01_00_11
    TCGA = 0b11_01_00_10
    TGAC = 0b10_11_00_01
    TGCA = 0b11_10_00_01
    GACT = 0b01_10_11_00
    GATC = 0b01_11_10_00
    GCAT = 0b10_01_11_00
    GCTA = 0b11_01_10_00
    GTAC = 0b10_11_01_00
    GTCA = 0b11_10_01_00


class _Xor10(int):
    def __repr__(self):
        return "Xor10"


Xor10 = _Xor10(nv.ENC_XOR10)


def word_for_k(word_bits: int, k: int) -> int:
    """kmer::word_for_k::<P, K>() (kmer.rs:67-69)."""
    per = word_bits // 2
    return (per + k - 1) // per


def num_bytes(word_bits: int, k: int) -> int:
    """Kmer::<P,K,B>::num_bytes (kmer.rs:41-43)."""
    return (word_bits // 8) * word_for_k(word_bits, k)


def encode(ctx, enc: int, seqs: np.ndarray, word_bits: int) -> np.ndarray:
    """Batched `Encoding::encode` / `Kmer::<P,K,B>::new` (encoding/naive.rs:116-124, kmer.rs:21-28).

    `seqs` is an (n, K) uint8 array of ASCII k-mers; returns the (n, B*word_bits/8)
    little-endian byte image of the n arrays [P; B], B = word_for_k."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    n, k = seqs.shape
    batch = ctx.upload(seqs.reshape(-1), fixed_len=k, n_reads=n)
    img, _ = batch.pack(int(enc), word_bits, to="host")
    return img.reshape(n, num_bytes(word_bits, k)) if n else img


def decode(ctx, enc: int, images: np.ndarray, word_bits: int, length=None) -> np.ndarray:
    """Batched `Encoding::decode` (encoding/naive.rs:126-136): (n, bytes) images -> (n, length) ASCII.
    length=None decodes every position of the array, padding included, like the reference."""
    images = np.ascontiguousarray(images, dtype=np.uint8)
    n, nb = images.shape
    return ctx.unpack(int(enc), word_bits, images, n, nb * 8 // word_bits, length)


def rev_comp(ctx, enc: int, k: int, images: np.ndarray, word_bits: int) -> np.ndarray:
    """Batched `Encoding::rev_comp::<K>` (encoding/naive.rs:138-154) on (n, bytes) images."""
    images = np.ascontiguousarray(images, dtype=np.uint8)
    n, nb = images.shape
    return ctx.revcomp_words(int(enc), k, word_bits, images, n, nb * 8 // word_bits).reshape(n, nb)


def get(ctx, images: np.ndarray, word_bits: int, index: int) -> np.ndarray:
    """Batched `Kmer::<P,K,B>::get(index)` (kmer.rs:46-48) on (n, bytes) images -> n 2-bit codes."""
    images = np.ascontiguousarray(images, dtype=np.uint8)
    n, nb = images.shape
    return ctx.kmer_get(word_bits, nb * 8 // word_bits, images, n, index)


def get_prefix(ctx, images: np.ndarray, word_bits: int, length: int) -> np.ndarray:
    """Batched `Kmer::<P,K,B>::get_prefix(len)` (kmer.rs:50-52; the reference's inclusive 0..=2*len bit range)
    on (n, bytes) images -> (n, word_bits/8) little-endian words."""
    images = np.ascontiguousarray(images, dtype=np.uint8)
    n, nb = images.shape
    return ctx.kmer_get_prefix(word_bits, nb * 8 // word_bits, images, n, length)


def bitmer_to_bytes(ctx, mers, length: int) -> np.ndarray:
    """Batched `bitmer_to_bytes` (kmer.rs:71-91): u64 words -> (n, length) upper-case ASCII."""
    return ctx.bitmer_to_bytes(mers, length)
