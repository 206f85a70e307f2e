"""kmers_b200 -- B200-native (sm_100a) hot path of COMBINE-lab/kmers.

Bulk ASCII -> 2-bit packing, every-window k-mer extraction, reverse complement,
canonical-min and LexHash, as hand-written CUDA behind the C ABI of
include/kmers_b200.h.  This package is the Python mirror of the reference's
interface for that path; there is no CPU fallback.
"""
from ._native import (ENC_ACGT, ENC_ACTG, ENC_XOR10, IDENTITY_MATCH, NO_MATCH, SENTINEL, TWIN_MATCH, Digest, KmbError,
                      KmbPanic, SO_PATH)
from .context import CanonicalKmers, Context, ReadBatch, host_pack, host_pack_isa, host_read_probe, parse_fastx
from .encoding import Naive, Xor10, decode, encode, num_bytes, rev_comp, word_for_k

__all__ = [
    "Context", "ReadBatch", "CanonicalKmers", "parse_fastx", "host_pack", "host_pack_isa", "host_read_probe", "Naive", "Xor10", "encode", "decode", "rev_comp", "word_for_k",
    "num_bytes", "Digest", "KmbError", "KmbPanic", "SENTINEL", "ENC_ACGT", "ENC_ACTG", "ENC_XOR10", "NO_MATCH",
    "IDENTITY_MATCH", "TWIN_MATCH", "SO_PATH",
]
