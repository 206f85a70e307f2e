"""CPU ORACLE bindings (test infrastructure, NOT the product).

ctypes view of ``oracle/libkmers_oracle.so`` (built from ``kmers_oracle.c`` by
``make -C oracle``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this
package.  Nothing under ``kmers_b200/`` does.

Parity status: PINNED against the reference's own golden vectors (see
``kmers_oracle.h``); the reference is Rust and cannot be built in this image.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkmers_oracle.so")  # portable build (-march=x86-64-v2): travels to any box


def use_native_build() -> str:
    """Switch this process to a `-O3 -march=native` build of the oracle, compiled ON THE MACHINE THAT RUNS IT (the
    shipped library is built portable because it travels to another host).  Used by bench.py's CPU legs so that the
    reported CPU baseline is not handicapped.  Must be called before the first `lib()`.  Returns the
    flags in effect ("-march=native", or the portable ones if the native build is not possible here)."""
    global _SO
    import hashlib
    assert _lib is None, "use_native_build() must come before the first oracle call"
    try:
        with open("/proc/cpuinfo") as f:
            ident = "".join(l for l in f if l.startswith(("model name", "flags")))[:20000]
        tag = hashlib.sha1(ident.encode()).hexdigest()[:10]
        so = os.path.join(_HERE, f"libkmers_oracle_native_{tag}.so")
        src, hdr = os.path.join(_HERE, "kmers_oracle.c"), os.path.join(_HERE, "kmers_oracle.h")
        if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr)):
            subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-fno-semantic-interposition", "-std=c11", "-pthread",
                                   "-shared", "-o", so + ".tmp", src], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            os.replace(so + ".tmp", so)
        _SO = so
        return "-O3 -march=native (built on this host)"
    except Exception:
        return "-O3 -march=x86-64-v2 (portable build; native build failed here)"

OK = 0
PANIC = -1
SENTINEL = 0xFFFFFFFFFFFFFFFF
INVALID_BASE = 0xFFFFFFFFFFFFFFFF
XOR10 = 0x100
NO_MATCH, IDENTITY_MATCH, TWIN_MATCH = 0, 1, 2

# Naive enum discriminants, encoding/naive.rs:49-74
NAIVE = {
    "ACTG": 0b00011011, "ACGT": 0b00011110, "ATCG": 0b00100111, "ATGC": 0b00110110,
    "AGCT": 0b00101101, "AGTC": 0b00111001, "CATG": 0b01001011, "CAGT": 0b01001110,
    "CTAG": 0b10000111, "CTGA": 0b11000110, "CGAT": 0b10001101, "CGTA": 0b11001001,
    "TACG": 0b01100011, "TAGC": 0b01110010, "TCAG": 0b10010011, "TCGA": 0b11010010,
    "TGAC": 0b10110001, "TGCA": 0b11100001, "GACT": 0b01101100, "GATC": 0b01111000,
    "GCAT": 0b10011100, "GCTA": 0b11011000, "GTAC": 0b10110100, "GTCA": 0b11100100,
}


class Kmer(C.Structure):
    _fields_ = [("k", C.c_uint8), ("data", C.c_uint64)]


class CanonicalKmer(C.Structure):
    _fields_ = [("fw", Kmer), ("rc", Kmer)]


class CkIter(C.Structure):
    _fields_ = [
        ("seq", C.c_void_p), ("seq_len", C.c_size_t), ("km", CanonicalKmer), ("pos", C.c_int32),
        ("invalid", C.c_int), ("last_invalid", C.c_int32), ("k", C.c_int32), ("strict", C.c_int),
    ]


class Digest(C.Structure):
    _fields_ = [("n_valid", C.c_uint64), ("checksum_canon", C.c_uint64), ("checksum_hash", C.c_uint64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kmers_oracle.c")
    hdr = os.path.join(_HERE, "kmers_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if _SO.endswith("libkmers_oracle.so"):
        build()
    L = C.CDLL(_SO)
    u8, u64, sz, i32, u32 = C.c_uint8, C.c_uint64, C.c_size_t, C.c_int, C.c_uint
    P = C.POINTER
    vp = C.c_void_p

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("ko_encode_binary_u8", u64, u8)
    sig("ko_encode_binary", i32, u8, P(u64))
    sig("ko_complement_base", u64, u64)
    sig("ko_is_valid_nuc", i32, u64)
    sig("ko_mask_table", u64, u32, i32)
    sig("ko_kmer_from_bytes", i32, vp, sz, P(Kmer))
    sig("ko_kmer_from_u64", Kmer, u64, u8, i32)
    sig("ko_kmer_to_string", None, Kmer, C.c_char_p)
    sig("ko_kmer_append_base", u64, P(Kmer), u64)
    sig("ko_kmer_prepend_base", u64, P(Kmer), u64, i32)
    sig("ko_kmer_append_base_u8", u64, P(Kmer), u8)
    sig("ko_kmer_prepend_base_u8", u64, P(Kmer), u8, i32)
    sig("ko_reverse_complement_word", u64, u64, u32)
    sig("ko_kmer_to_reverse_complement", Kmer, Kmer)
    sig("ko_kmer_is_canonical", i32, Kmer)
    sig("ko_kmer_to_canonical", Kmer, Kmer)
    sig("ko_kmer_cmp", i32, Kmer, Kmer)
    sig("ko_sub_kmer_word", i32, u64, sz, sz, sz, i32, P(u64))
    sig("ko_lexhash_word", u64, u64, u32)
    sig("ko_ck_blank_of_size", CanonicalKmer, u8)
    sig("ko_ck_from_u64", CanonicalKmer, u64, u8, i32)
    sig("ko_ck_from_bytes", i32, vp, sz, P(CanonicalKmer))
    sig("ko_ck_swap", None, P(CanonicalKmer))
    sig("ko_ck_is_fw_canonical", i32, P(CanonicalKmer))
    sig("ko_ck_append_base", u64, P(CanonicalKmer), u64, i32)
    sig("ko_ck_prepend_base", u64, P(CanonicalKmer), u64, i32)
    sig("ko_ck_append_base_u8", u64, P(CanonicalKmer), u8, i32)
    sig("ko_ck_prepend_base_u8", u64, P(CanonicalKmer), u8, i32)
    sig("ko_ck_get_canonical_word", u64, P(CanonicalKmer))
    sig("ko_ck_get_word_equivalency", i32, P(CanonicalKmer), u64)
    sig("ko_iter_from_u8_slice", None, P(CkIter), vp, sz, u8, i32)
    sig("ko_iter_exhausted", i32, P(CkIter))
    sig("ko_iter_inc", i32, P(CkIter))
    sig("ko_iter_inc_by", i32, P(CkIter), sz)
    sig("ko_nuc2bits", u8, i32, u8)
    sig("ko_bits2nuc", u8, i32, u8)
    sig("ko_complement_bits", u8, i32, u8)
    sig("ko_rev_encoding", u8, u8)
    sig("ko_encode", i32, i32, vp, sz, u32, sz, vp)
    sig("ko_decode", None, i32, vp, u32, sz, vp)
    sig("ko_rev_comp", i32, i32, u32, u32, sz, vp, i32)
    sig("ko_word_for_k", sz, u32, sz)
    sig("ko_num_bytes", sz, u32, sz)
    sig("ko_kmer_get", u8, vp, sz)
    sig("ko_kmer_get_prefix", u64, vp, sz)
    sig("ko_bitmer_to_bytes", None, u64, sz, vp)
    sig("ko_splitmix64", u64, u64)
    sig("ko_generate_bases", None, u64, u64, sz, C.c_uint32, vp)
    sig("ko_count_slots", u64, vp, sz, u64, u32)
    sig("ko_extract_canonical", i32, vp, vp, sz, u64, u32, i32, vp, vp, vp, vp, vp, u32, P(Digest), i32)
    sig("ko_bench_windows", i32, vp, vp, sz, u64, u32, vp, vp, P(Digest), i32)
    sig("ko_extract_canonical_wide", i32, vp, vp, sz, u64, u32, i32, i32, vp, vp, P(Digest))
    sig("ko_extract_canonical_wide_mt", i32, vp, vp, sz, u64, u32, i32, i32, vp, vp, P(Digest), i32)
    sig("ko_minimizer_word", i32, u64, sz, sz, u32, i32, P(u64), P(sz))
    sig("ko_sv_from_bytes", i32, vp, sz, vp)
    sig("ko_sv_get_kmer_u64", i32, vp, sz, sz, sz, P(u64))
    sig("ko_sv_minimizers", i32, vp, sz, sz, sz, u32, vp, vp)
    sig("ko_minimizers_batch", i32, vp, vp, sz, u64, u32, u32, u32, vp, vp)
    _lib = L
    return L


# --------------------------------------------------------------------------
# small pythonic helpers used by the tests
# --------------------------------------------------------------------------

def _buf(b: bytes):
    return C.cast(C.c_char_p(b), C.c_void_p)


def kmer_from(s) -> Kmer:
    """naive_impl::Kmer::from(&str / &[u8]); raises on the reference's panics."""
    if isinstance(s, str):
        s = s.encode()
    km = Kmer()
    if lib().ko_kmer_from_bytes(_buf(s), len(s), C.byref(km)) != OK:
        raise RuntimeError("panic: Kmer::from")
    return km


def kmer_str(km: Kmer) -> str:
    out = C.create_string_buffer(km.k + 1)
    lib().ko_kmer_to_string(km, out)
    return out.value.decode()


def ck_from(s) -> CanonicalKmer:
    if isinstance(s, str):
        s = s.encode()
    ck = CanonicalKmer()
    if lib().ko_ck_from_bytes(_buf(s), len(s), C.byref(ck)) != OK:
        raise RuntimeError("panic: CanonicalKmer::from")
    return ck


class Iter:
    """CanonicalKmerIterator over a bytes object (kept alive here)."""

    def __init__(self, seq: bytes, k: int, strict: bool = False):
        self._seq = bytes(seq)
        self._keep = C.create_string_buffer(self._seq, len(self._seq) + 1)
        self.it = CkIter()
        lib().ko_iter_from_u8_slice(C.byref(self.it), C.cast(self._keep, C.c_void_p), len(self._seq), k,
                                    int(strict))

    def exhausted(self) -> bool:
        return bool(lib().ko_iter_exhausted(C.byref(self.it)))

    def inc(self) -> bool:
        return bool(lib().ko_iter_inc(C.byref(self.it)))

    def inc_by(self, n: int) -> bool:
        return bool(lib().ko_iter_inc_by(C.byref(self.it), n))

    @property
    def pos(self) -> int:
        return self.it.pos

    @property
    def km(self) -> CanonicalKmer:
        return self.it.km


def encode(enc: int, seq: bytes, word_bits: int, n_words: int) -> np.ndarray:
    """Encoding::encode -> little-endian byte image of [P; B]."""
    out = np.zeros(n_words * word_bits // 8, dtype=np.uint8)
    if lib().ko_encode(enc, _buf(seq), len(seq), word_bits, n_words, out.ctypes.data) != OK:
        raise RuntimeError("panic: encode")
    return out


def words(img: np.ndarray, word_bits: int) -> list:
    """Interpret a byte image as python ints of word_bits each (LE)."""
    b = img.tobytes()
    n = word_bits // 8
    return [int.from_bytes(b[i:i + n], "little") for i in range(0, len(b), n)]


def image(ws, word_bits: int) -> np.ndarray:
    n = word_bits // 8
    return np.frombuffer(b"".join(int(w).to_bytes(n, "little") for w in ws), dtype=np.uint8).copy()


def decode(enc: int, img: np.ndarray, word_bits: int) -> bytes:
    n_words = img.size * 8 // word_bits
    out = np.zeros(n_words * word_bits // 2, dtype=np.uint8)
    lib().ko_decode(enc, img.ctypes.data, word_bits, n_words, out.ctypes.data)
    return out.tobytes()


def rev_comp(enc: int, k: int, img: np.ndarray, word_bits: int, strict: bool = False) -> np.ndarray:
    out = np.ascontiguousarray(img).copy()
    n_words = out.size * 8 // word_bits
    if lib().ko_rev_comp(enc, k, word_bits, n_words, out.ctypes.data, int(strict)) != OK:
        raise RuntimeError("panic: rev_comp")
    return out


def generate_bases(seed: int, first_index: int, n: int, n_thresh20: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    lib().ko_generate_bases(seed, first_index, n, n_thresh20, out.ctypes.data)
    return out


def _offs_ptr(offsets):
    if offsets is None:
        return None, None
    o = np.ascontiguousarray(offsets, dtype=np.uint64)
    return o, o.ctypes.data


def count_slots(offsets, n_reads: int, fixed_len: int, k: int) -> int:
    o, p = _offs_ptr(offsets)
    return int(lib().ko_count_slots(p, n_reads, fixed_len, k))


def extract_canonical(bases: np.ndarray, k: int, *, offsets=None, n_reads=None, fixed_len=0,
                      strict=False, want_fw_rc=False, hist_bits=0, n_threads=1, materialize=True,
                      canon_out=None, hash_out=None):
    """Dense-slot canonical extraction through the restated iterator.

    Returns dict(canon, hash, [fw, rc], [hist], n_valid, checksum_canon, checksum_hash)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    o, op = _offs_ptr(offsets)
    if o is not None:
        n_reads = o.size - 1
    n_slots = count_slots(o, n_reads, fixed_len, k)
    canon = canon_out if canon_out is not None else (np.empty(n_slots, dtype=np.uint64) if materialize else None)
    hsh = hash_out if hash_out is not None else (np.empty(n_slots, dtype=np.uint64) if materialize else None)
    fw = np.empty(n_slots, dtype=np.uint64) if want_fw_rc else None
    rc = np.empty(n_slots, dtype=np.uint64) if want_fw_rc else None
    hist = np.zeros(1 << hist_bits, dtype=np.uint64) if hist_bits else None
    d = Digest()
    ptr = lambda a: a.ctypes.data if a is not None else None
    st = lib().ko_extract_canonical(bases.ctypes.data, op, n_reads, fixed_len, k, int(strict), ptr(canon),
                                    ptr(hsh), ptr(fw), ptr(rc), ptr(hist), hist_bits, C.byref(d), n_threads)
    if st != OK:
        raise RuntimeError("panic: extract_canonical")
    return dict(canon=canon, hash=hsh, fw=fw, rc=rc, hist=hist, n_valid=d.n_valid,
                checksum_canon=d.checksum_canon, checksum_hash=d.checksum_hash, n_slots=n_slots)


def bench_windows(bases: np.ndarray, k: int, *, offsets=None, n_reads=None, fixed_len=0, n_threads=1,
                  materialize=True):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    o, op = _offs_ptr(offsets)
    if o is not None:
        n_reads = o.size - 1
    n_slots = count_slots(o, n_reads, fixed_len, k)
    canon = np.empty(n_slots, dtype=np.uint64) if materialize else None
    hsh = np.empty(n_slots, dtype=np.uint64) if materialize else None
    d = Digest()
    ptr = lambda a: a.ctypes.data if a is not None else None
    st = lib().ko_bench_windows(bases.ctypes.data, op, n_reads, fixed_len, k, ptr(canon), ptr(hsh),
                                C.byref(d), n_threads)
    if st != OK:
        raise RuntimeError("panic: bench_windows (non-ACGT input)")
    return dict(canon=canon, hash=hsh, n_valid=d.n_valid, checksum_canon=d.checksum_canon,
                checksum_hash=d.checksum_hash, n_slots=n_slots)


def extract_canonical_wide(bases: np.ndarray, k: int, *, enc=NAIVE["ACGT"], validate=True, offsets=None,
                           n_reads=None, fixed_len=0, want_hash=True, n_threads=1):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    o, op = _offs_ptr(offsets)
    if o is not None:
        n_reads = o.size - 1
    n_slots = count_slots(o, n_reads, fixed_len, k)
    canon = np.empty(2 * n_slots, dtype=np.uint64)
    hsh = np.empty(2 * n_slots, dtype=np.uint64) if want_hash else None
    d = Digest()
    st = lib().ko_extract_canonical_wide_mt(bases.ctypes.data, op, n_reads, fixed_len, k, enc, int(validate),
                                            canon.ctypes.data, hsh.ctypes.data if want_hash else None,
                                            C.byref(d), n_threads)
    if st != OK:
        raise RuntimeError("panic: extract_canonical_wide")
    return dict(canon=canon.reshape(-1, 2), hash=hsh.reshape(-1, 2) if want_hash else None,
                n_valid=d.n_valid, checksum_canon=d.checksum_canon, checksum_hash=d.checksum_hash,
                n_slots=n_slots)


# ---- "next" rows: minimizers + SeqVector -------------------------------------------------------------
def minimizer_word(word: int, k: int, width: int, hash_k: int, strict: bool = False):
    """Kmer::minimizer_word with LexHasherState(hash_k) -> (mmer word, offset)."""
    m, o = C.c_uint64(), C.c_size_t()
    if lib().ko_minimizer_word(word, k, width, hash_k, int(strict), C.byref(m), C.byref(o)) != OK:
        raise RuntimeError("panic: minimizer_word")
    return int(m.value), int(o.value)


def sv_from_bytes(seq: bytes) -> np.ndarray:
    """SeqVector::from(&[u8]) -> its u64 words."""
    out = np.zeros((len(seq) + 31) // 32, dtype=np.uint64)
    if lib().ko_sv_from_bytes(_buf(seq), len(seq), out.ctypes.data) != OK:
        raise RuntimeError("panic: SeqVector::from")
    return out


def sv_get_kmer_u64(words: np.ndarray, length: int, pos: int, k: int) -> int:
    v = C.c_uint64()
    if lib().ko_sv_get_kmer_u64(words.ctypes.data, length, pos, k, C.byref(v)) != OK:
        raise RuntimeError("panic: get_kmer_u64")
    return int(v.value)


def sv_minimizers(seq: bytes, k: int, w: int, hash_k: int):
    """SeqVecMinimizerIter over SeqVector::from(seq) -> list of (word, pos)."""
    words = sv_from_bytes(seq)
    n = len(seq) - k + 1
    if n <= 0:
        raise RuntimeError("panic: assert!(sv.len() >= k)")
    mw, mp = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)
    if lib().ko_sv_minimizers(words.ctypes.data, len(seq), k, w, hash_k, mw.ctypes.data, mp.ctypes.data) != OK:
        raise RuntimeError("panic: SeqVecMinimizerIter")
    return list(zip(mw.tolist(), mp.tolist()))


def minimizers_batch(bases: np.ndarray, k: int, w: int, hash_k: int, *, offsets=None, n_reads=None, fixed_len=0):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    o, op = _offs_ptr(offsets)
    if o is not None:
        n_reads = o.size - 1
    n_slots = count_slots(o, n_reads, fixed_len, k)
    mm, pos = np.empty(n_slots, dtype=np.uint64), np.empty(n_slots, dtype=np.uint32)
    if lib().ko_minimizers_batch(bases.ctypes.data, op, n_reads, fixed_len, k, w, hash_k, mm.ctypes.data, pos.ctypes.data) != OK:
        raise RuntimeError("panic: minimizers_batch")
    return mm, pos
