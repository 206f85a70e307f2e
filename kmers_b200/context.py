"""Host-side mirror of the reference interface, in batched form.

`Context` wraps a `kmb_ctx` (one CUDA stream + the device-resident read batch);
`ReadBatch` is the batched stand-in for the `&[u8]` a caller hands to
`CanonicalKmerIterator::from_u8_slice` (naive_impl/canonical_kmer_iterator.rs:72-83)
or `Encoding::encode` (encoding/mod.rs:16).  PyTorch is used only as the owner
of device memory and streams; all compute goes through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _native as nv
from ._native import Digest, KmbError, KmbPanic, check


def _torch():
    import torch
    return torch


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x) -> Optional[int]:
    """Raw address of a numpy array (host) or torch tensor (host or device)."""
    if x is None:
        return None
    if _is_tensor(x):
        assert x.is_contiguous()
        return x.data_ptr()
    assert x.flags["C_CONTIGUOUS"]
    return x.ctypes.data


def _as_u64_host(t) -> np.ndarray:
    """int64 torch tensor (device or host) -> numpy uint64 view."""
    return t.detach().cpu().numpy().view(np.uint64)


@dataclass
class CanonicalKmers:
    """Dense-slot result of `ReadBatch.extract_canonical` (SURVEY.md 8d layout).

    Slot `win_offsets[r] + pos` holds the window at `pos` of read `r`; windows
    the reference iterator skips hold SENTINEL.  Arrays are torch int64 CUDA
    tensors (bit pattern of the u64 words) or numpy uint64 arrays."""
    k: int
    n_slots: int
    canon: object = None
    hash: object = None
    fw: object = None
    rc: object = None
    digest: Optional[tuple] = None  # (n_valid, checksum_canon, checksum_hash)
    words_per_kmer: int = 1

    def host(self, name: str) -> np.ndarray:
        a = getattr(self, name)
        if a is None:
            raise ValueError(f"{name} was not requested")
        out = _as_u64_host(a) if _is_tensor(a) else a
        return out.reshape(-1, self.words_per_kmer) if self.words_per_kmer > 1 else out


class Context:
    """One per (host thread, GPU).  Not thread-safe; contexts are independent."""

    def __init__(self, device: int = -1, stream: Optional[int] = None):
        """stream=None: the context owns a private non-blocking stream.  stream=<cudaStream_t handle>
        (e.g. torch.cuda.current_stream().cuda_stream): borrow it; handle 0 means the legacy default stream."""
        self._lib = nv.lib()
        if stream is not None and int(stream) == 0:
            stream = 1  # cudaStreamLegacy
        h = C.c_void_p()
        check(None, self._lib.kmb_ctx_create(device, stream, C.byref(h)))
        self._h = h
        self._keep = []  # objects a borrowed batch must outlive

    # ---- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.kmb_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, code: int):
        check(self._h, code)

    def sync(self):
        self._ck(self._lib.kmb_ctx_sync(self._h))

    # ---- stream ordering against torch (the owner of the device tensors handed in / out)
    # Async contract: calls that write DEVICE outputs return once the work is enqueued on the context's stream.  The
    # tensors are allocated by torch on ITS current stream, so when the two streams differ (the default: a context owns a
    # private non-blocking stream) `_order_in` makes the context's stream wait for torch's pending work before the call and
    # `_order_out` makes torch's current stream wait for the kernels after it.  Any torch op on the results -- `.cpu()`,
    # `CanonicalKmers.host()`, or the caching allocator handing the memory of a dropped tensor to later work on that
    # stream -- is then ordered after the kernels.  (No `record_stream`: the allocator would later record events on the
    # context's stream, which may be gone by then -- contexts are closed while result tensors live on.)
    def _streams(self):
        t = _torch()
        if not t.cuda.is_available():
            return None
        cur = t.cuda.current_stream()
        mine = self.stream
        if mine in (0, 1):  # legacy default stream: synchronises with torch's default stream implicitly
            mine = 0
        if int(cur.cuda_stream) == mine:
            return None
        return t, cur, t.cuda.ExternalStream(self.stream if self.stream not in (0, 1) else 0)

    def _order_in(self, *xs):
        if not any(_is_tensor(x) and x.is_cuda for x in xs if x is not None):
            return
        st = self._streams()
        if st is not None:
            _, cur, ext = st
            ext.wait_stream(cur)

    def _order_out(self, *xs):
        ts = [x for x in xs if x is not None and _is_tensor(x) and x.is_cuda]
        if not ts:
            return
        st = self._streams()
        if st is not None:
            _, cur, ext = st
            cur.wait_stream(ext)

    @property
    def stream(self) -> int:
        return int(self._lib.kmb_ctx_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._lib.kmb_ctx_launch_count(self._h))

    # ---- batches
    def upload(self, bases, offsets=None, fixed_len: int = 0, n_reads: Optional[int] = None) -> "ReadBatch":
        """Host reads -> pinned staging -> device (kmb_batch_upload)."""
        bases = np.ascontiguousarray(np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray))
                                     else bases, dtype=np.uint8)
        offs = None
        if offsets is not None:
            offs = np.ascontiguousarray(offsets, dtype=np.uint64)
            n_reads = offs.size - 1
        elif n_reads is None:
            n_reads = bases.size // fixed_len if fixed_len else 0
        self._ck(self._lib.kmb_batch_upload(self._h, _ptr(bases), bases.size, _ptr(offs), n_reads, fixed_len))
        self._keep = []
        return ReadBatch(self, bases.size, n_reads, fixed_len, offs is not None)

    def attach(self, dev_bases, dev_offsets=None, fixed_len: int = 0, n_reads: Optional[int] = None) -> "ReadBatch":
        """Borrow device-resident reads (torch uint8 CUDA tensor [+ int64 offsets]) with no copy."""
        n_bytes = dev_bases.numel()
        if dev_offsets is not None:
            n_reads = dev_offsets.numel() - 1
        elif n_reads is None:
            n_reads = n_bytes // fixed_len if fixed_len else 0
        self._order_in(dev_bases, dev_offsets)
        self._ck(self._lib.kmb_batch_attach(self._h, _ptr(dev_bases), n_bytes, _ptr(dev_offsets), n_reads, fixed_len))
        self._keep = [dev_bases, dev_offsets]
        return ReadBatch(self, n_bytes, n_reads, fixed_len, dev_offsets is not None)

    def generate(self, seed: int, n_reads: int, fixed_len: int, n_thresh20: int = 0, first_index: int = 0) -> "ReadBatch":
        """Synthetic reads generated on the device (counter-based splitmix64)."""
        self._ck(self._lib.kmb_batch_generate(self._h, seed, first_index, n_reads, fixed_len, n_thresh20))
        self._keep = []
        return ReadBatch(self, n_reads * fixed_len, n_reads, fixed_len, False)

    def new_packed(self, capacity_bases: int = 0) -> "ReadBatch":
        """SeqVector::with_capacity (seq_vector.rs:135-139): an empty packed sequence, grown with ReadBatch.push_chars."""
        self._ck(self._lib.kmb_batch_new_packed(self._h, capacity_bases))
        self._keep = []
        b = ReadBatch(self, 0, 1, 0, False)
        b.packed = True
        return b

    def ingest_fastx(self, text: bytes) -> "ReadBatch":
        """FASTA / FASTQ text -> pinned host batch -> device (kmb_batch_ingest_fastx); returns the ragged ReadBatch."""
        nr, nb = C.c_uint64(), C.c_uint64()
        self._ck(self._lib.kmb_batch_ingest_fastx(self._h, text, len(text), C.byref(nr), C.byref(nb)))
        self._keep = []
        return ReadBatch(self, int(nb.value), int(nr.value), 0, True)

    # ---- one-shot host path (e2e): chunked; host packing, H2D, kernel and D2H overlapped
    def extract_canonical_host(self, host_bases: np.ndarray, n_reads: int, fixed_len: int, k: int, *,
                               out_canon=None, out_hash=None, host_canon=None, host_hash=None,
                               validate: bool = True, digest: bool = True):
        """kmb_extract_canonical_host: reads in HOST memory -> canonical words / LexHashes.

        out_canon / out_hash: torch int64 CUDA tensors for the whole batch (results stay resident on the device), numpy
        uint64 / pinned torch arrays (results copied back), or None (not produced).  host_canon / host_hash are the
        round-1 names of the same arguments.  Returns the digest tuple (or None)."""
        out_canon = host_canon if out_canon is None else out_canon
        out_hash = host_hash if out_hash is None else out_hash
        d = Digest()
        flags = 0 if validate else nv.F_NO_VALIDATE
        self._order_in(out_canon, out_hash)
        self._ck(self._lib.kmb_extract_canonical_host(self._h, _ptr(host_bases), n_reads, fixed_len, k, flags,
                                                      _ptr(out_canon), _ptr(out_hash),
                                                      C.byref(d) if digest else None))
        self._order_out(out_canon, out_hash)
        return d.astuple() if digest else None

    def extract_canonical_host_packed(self, host_bits: np.ndarray, host_inv: Optional[np.ndarray], n_reads: int,
                                      fixed_len: int, k: int, *, out_canon=None, out_hash=None, validate: bool = True,
                                      digest: bool = True):
        """kmb_extract_canonical_host_packed: reads already 2-bit packed on the host (`host_pack`)."""
        d = Digest()
        flags = 0 if validate else nv.F_NO_VALIDATE
        self._order_in(out_canon, out_hash)
        self._ck(self._lib.kmb_extract_canonical_host_packed(self._h, _ptr(host_bits), _ptr(host_inv), n_reads, fixed_len,
                                                             k, flags, _ptr(out_canon), _ptr(out_hash),
                                                             C.byref(d) if digest else None))
        self._order_out(out_canon, out_hash)
        return d.astuple() if digest else None

    def set_host_threads(self, n: int):
        """Worker threads of the host pipeline (0 = default)."""
        self._ck(self._lib.kmb_ctx_set_host_threads(self._h, n))

    def host_stats(self) -> dict:
        """What the last extract_canonical_host* call did."""
        st = (C.c_uint64 * 4)()
        self._ck(self._lib.kmb_ctx_host_stats(self._h, st))
        return {"chunks": int(st[0]), "raw_chunks": int(st[1]), "h2d_bytes": int(st[2]), "d2h_bytes": int(st[3])}

    # ---- element-wise word ops (naive_impl::Kmer in batch)
    @staticmethod
    def _words_in(words, to):
        """(count, array, destination) for a u64 word array given as torch int64 (device) or numpy."""
        if _is_tensor(words):
            return words.numel(), words, to or "device"
        w = np.ascontiguousarray(words, dtype=np.uint64)
        return w.size, w, to or "host"

    def _alloc(self, n, to, dtype):
        if to == "device":
            t = _torch()
            return t.empty(n, dtype=t.int64 if dtype == np.uint64 else t.uint8, device="cuda")
        return np.empty(n, dtype=dtype)

    def reverse_complement_words(self, words, k: int, to=None):
        """Kmer::get_reverse_complement_word (naive_impl/kmer.rs:138-147) on every word."""
        n, w, to = self._words_in(words, to)
        out = self._alloc(n, to, np.uint64)
        self._order_in(w, out)
        self._ck(self._lib.kmb_reverse_complement_words(self._h, k, _ptr(w), _ptr(out), n))
        self.sync()
        return out

    def canonical_words(self, words, k: int, to=None):
        """(Kmer::to_canonical, Kmer::is_canonical) (naive_impl/kmer.rs:55-74) on every word."""
        n, w, to = self._words_in(words, to)
        out = self._alloc(n, to, np.uint64)
        flag = self._alloc(n, to, np.uint8)
        self._order_in(w, out, flag)
        self._ck(self._lib.kmb_canonical_words(self._h, k, _ptr(w), _ptr(out), _ptr(flag), n))
        self.sync()
        return out, flag

    def lexhash_words(self, words, k: int, to=None):
        """hash_one(&LexHasherState::new(k), kmer) (naive_impl/hash.rs:10-20, 60-71) on every word."""
        n, w, to = self._words_in(words, to)
        out = self._alloc(n, to, np.uint64)
        self._order_in(w, out)
        self._ck(self._lib.kmb_lexhash_words(self._h, k, _ptr(w), _ptr(out), n))
        self.sync()
        return out

    def match_words(self, words, others, k: int, to=None):
        """CanonicalKmer::from_u64(w, k).get_word_equivalency(other) (canonical_kmer.rs:152-161) -> MatchType u8."""
        n, w, to = self._words_in(words, to)
        o = others if _is_tensor(others) else np.ascontiguousarray(others, dtype=np.uint64)
        out = self._alloc(n, to, np.uint8)
        self._order_in(w, o, out)
        self._ck(self._lib.kmb_match_words(self._h, k, _ptr(w), _ptr(o), _ptr(out), n))
        self.sync()
        return out

    def minimizer_words(self, words, k: int, w: int, hash_k=None, to=None):
        """Kmer::minimizer_word (naive_impl/kmer.rs:170-191) with LexHasherState(hash_k) on every word -> (mmer, offset)."""
        hash_k = w if hash_k is None else hash_k
        n, wd, to = self._words_in(words, to)
        mm = self._alloc(n, to, np.uint64)
        if to == "device":
            t = _torch()
            off = t.empty(n, dtype=t.int32, device="cuda")
        else:
            off = np.empty(n, dtype=np.uint32)
        self._order_in(wd, mm, off)
        self._ck(self._lib.kmb_minimizer_words(self._h, k, w, hash_k, _ptr(wd), n, _ptr(mm), _ptr(off)))
        self.sync()
        return mm, off

    # ---- the small Kmer / CanonicalKmer accessors, batched (host numpy in, host numpy out)
    def sub_kmer_words(self, words, k: int, pos: int, width: int) -> np.ndarray:
        """Kmer::sub_kmer_word (naive_impl/kmer.rs:150-161) on every word."""
        w = np.ascontiguousarray(words, dtype=np.uint64)
        out = np.empty_like(w)
        self._ck(self._lib.kmb_sub_kmer_words(self._h, k, pos, width, _ptr(w), _ptr(out), w.size))
        return out

    def _shift(self, fn, words, bases, k, ascii_):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        b = np.ascontiguousarray(np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else bases, dtype=np.uint8)
        assert b.size == w.size
        out, dropped = np.empty_like(w), np.empty(w.size, dtype=np.uint8)
        self._ck(fn(self._h, k, _ptr(w), _ptr(b), int(ascii_), _ptr(out), _ptr(dropped), w.size))
        return out, dropped

    def append_base_words(self, words, bases, k: int, ascii: bool = False):
        """Kmer::append_base / append_base_u8 (naive_impl/kmer.rs:83-102) -> (shifted words, bases shifted off)."""
        return self._shift(self._lib.kmb_append_base_words, words, bases, k, ascii)

    def prepend_base_words(self, words, bases, k: int, ascii: bool = False):
        """Kmer::prepend_base / prepend_base_u8 (naive_impl/kmer.rs:76-95) -> (shifted words, bases shifted off)."""
        return self._shift(self._lib.kmb_prepend_base_words, words, bases, k, ascii)

    def _cshift(self, fn, fw, rc, bases, k, ascii_):
        f, r = np.ascontiguousarray(fw, dtype=np.uint64), np.ascontiguousarray(rc, dtype=np.uint64)
        b = np.ascontiguousarray(np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else bases, dtype=np.uint8)
        fo, ro, dropped = np.empty_like(f), np.empty_like(r), np.empty(f.size, dtype=np.uint8)
        self._ck(fn(self._h, k, _ptr(f), _ptr(r), _ptr(b), int(ascii_), _ptr(fo), _ptr(ro), _ptr(dropped), f.size))
        return fo, ro, dropped

    def canonical_append_base_words(self, fw, rc, bases, k: int, ascii: bool = False):
        """CanonicalKmer::append_base[_u8] (canonical_kmer.rs:70-92) on (fw, rc) pairs -> (fw, rc, dropped)."""
        return self._cshift(self._lib.kmb_canonical_append_base_words, fw, rc, bases, k, ascii)

    def canonical_prepend_base_words(self, fw, rc, bases, k: int, ascii: bool = False):
        """CanonicalKmer::prepend_base[_u8] (canonical_kmer.rs:79-100) on (fw, rc) pairs -> (fw, rc, dropped)."""
        return self._cshift(self._lib.kmb_canonical_prepend_base_words, fw, rc, bases, k, ascii)

    def is_fw_canonical_words(self, fw, rc) -> np.ndarray:
        """CanonicalKmer::is_fw_canonical (canonical_kmer.rs:67-69)."""
        f, r = np.ascontiguousarray(fw, dtype=np.uint64), np.ascontiguousarray(rc, dtype=np.uint64)
        out = np.empty(f.size, dtype=np.uint8)
        self._ck(self._lib.kmb_is_fw_canonical_words(self._h, _ptr(f), _ptr(r), _ptr(out), f.size))
        return out

    def kmer_get(self, word_bits: int, words_per_item: int, arrays: np.ndarray, n_items: int, index: int) -> np.ndarray:
        """kmer::Kmer<P,K,B>::get(index) (kmer.rs:46-48) on n_items arrays (byte image) -> 2-bit codes."""
        img = np.ascontiguousarray(arrays).view(np.uint8).reshape(-1)
        out = np.empty(n_items, dtype=np.uint8)
        self._ck(self._lib.kmb_kmer_get(self._h, word_bits, words_per_item, _ptr(img), n_items, index, _ptr(out)))
        return out

    def kmer_get_prefix(self, word_bits: int, words_per_item: int, arrays: np.ndarray, n_items: int, length: int) -> np.ndarray:
        """kmer::Kmer<P,K,B>::get_prefix(len) (kmer.rs:50-52; 2 len + 1 bits, as the reference) -> (n_items, word_bytes) byte image."""
        img = np.ascontiguousarray(arrays).view(np.uint8).reshape(-1)
        out = np.empty(n_items * word_bits // 8, dtype=np.uint8)
        self._ck(self._lib.kmb_kmer_get_prefix(self._h, word_bits, words_per_item, _ptr(img), n_items, length, _ptr(out)))
        return out.reshape(n_items, word_bits // 8)

    def bitmer_to_bytes(self, mers, length: int) -> np.ndarray:
        """bitmer_to_bytes (kmer.rs:71-91) on every u64 -> (n, length) upper-case ASCII."""
        w = np.ascontiguousarray(mers, dtype=np.uint64)
        out = np.empty((w.size, length), dtype=np.uint8)
        self._ck(self._lib.kmb_bitmer_to_bytes(self._h, length, _ptr(w), w.size, _ptr(out)))
        self.sync()
        return out

    # ---- batched Encoding::decode / rev_comp on arrays [P; B]
    def words_to_strings(self, words, k: int) -> np.ndarray:
        """Batched `String::from(Kmer)` (naive_impl/kmer.rs:196-207): (n, k) lower-case ASCII (kmb_words_to_strings)."""
        w = np.ascontiguousarray(words, dtype=np.uint64)
        out = np.empty((w.size, k), dtype=np.uint8)
        self._ck(self._lib.kmb_words_to_strings(self._h, k, w.ctypes.data, w.size, out.ctypes.data))
        return out

    def unpack(self, enc: int, word_bits: int, words: np.ndarray, n_items: int, words_per_item: int,
               bases_per_item: Optional[int] = None) -> np.ndarray:
        """Encoding::decode (encoding/naive.rs:126-136) of n_items arrays; default length reproduces the
        reference's padding positions."""
        if bases_per_item is None:
            bases_per_item = words_per_item * word_bits // 2
        img = np.ascontiguousarray(words).view(np.uint8).reshape(-1)
        out = np.empty(n_items * bases_per_item, dtype=np.uint8)
        self._ck(self._lib.kmb_unpack(self._h, enc, word_bits, _ptr(img), n_items, words_per_item, bases_per_item,
                                      _ptr(out)))
        self.sync()
        return out.reshape(n_items, bases_per_item) if n_items else out

    def revcomp_words(self, enc: int, k: int, word_bits: int, words: np.ndarray, n_items: int,
                      words_per_item: int) -> np.ndarray:
        """Encoding::rev_comp::<K> (encoding/naive.rs:138-154) of n_items arrays (byte image in, byte image out)."""
        img = np.ascontiguousarray(words).view(np.uint8).reshape(-1)
        out = np.empty_like(img)
        self._ck(self._lib.kmb_revcomp_words(self._h, enc, k, word_bits, words_per_item, _ptr(img), _ptr(out), n_items))
        self.sync()
        return out


class ReadBatch:
    """The device-resident read batch of a Context (one at a time per context)."""

    def __init__(self, ctx: Context, n_bytes: int, n_reads: int, fixed_len: int, ragged: bool):
        self.ctx = ctx
        self.n_bytes, self.n_reads, self.fixed_len, self.ragged = n_bytes, n_reads, fixed_len, ragged
        self.packed = False

    def num_slots(self, k: int) -> int:
        n = C.c_uint64()
        self.ctx._ck(self.ctx._lib.kmb_batch_num_slots(self.ctx._h, k, C.byref(n)))
        return int(n.value)

    def window_offsets(self, k: int) -> np.ndarray:
        """Exclusive prefix of per-read window counts (n_reads + 1)."""
        if not self.ragged:
            w = max(0, self.fixed_len - k + 1)
            return np.arange(self.n_reads + 1, dtype=np.uint64) * np.uint64(w)
        out = np.empty(self.n_reads + 1, dtype=np.uint64)
        self.ctx._ck(self.ctx._lib.kmb_batch_window_offsets(self.ctx._h, k, _ptr(out)))
        return out

    def download(self) -> np.ndarray:
        out = np.empty(self.n_bytes, dtype=np.uint8)
        self.ctx._ck(self.ctx._lib.kmb_batch_download(self.ctx._h, _ptr(out), self.n_bytes))
        return out

    def _alloc(self, n, to):
        if to == "device":
            t = _torch()
            return t.empty(n, dtype=t.int64, device="cuda")
        return np.empty(n, dtype=np.uint64)

    def extract_canonical(self, k: int, *, want_hash: bool = True, want_fw_rc: bool = False, digest: bool = False,
                          validate: bool = True, to: str = "device", out: Optional[CanonicalKmers] = None) -> CanonicalKmers:
        """Batched CanonicalKmerIterator + get_canonical_word + LexHasher (kmb_extract_canonical)."""
        n = self.num_slots(k)
        if out is None:
            out = CanonicalKmers(k=k, n_slots=n)
            out.canon = self._alloc(n, to)
            out.hash = self._alloc(n, to) if want_hash else None
            if want_fw_rc:
                out.fw, out.rc = self._alloc(n, to), self._alloc(n, to)
        d = Digest()
        flags = 0 if validate else nv.F_NO_VALIDATE
        self.ctx._order_in(out.canon, out.hash, out.fw, out.rc)
        self.ctx._ck(self.ctx._lib.kmb_extract_canonical(self.ctx._h, k, flags, _ptr(out.canon), _ptr(out.hash),
                                                         _ptr(out.fw), _ptr(out.rc), C.byref(d) if digest else None))
        self.ctx._order_out(out.canon, out.hash, out.fw, out.rc)
        out.digest = d.astuple() if digest else None
        return out

    def to_packed(self, strict: bool = True) -> "ReadBatch":
        """Switch the resident batch to a 2-bit packed store (SeqVector layout; kmb_batch_repack).  strict: raise
        KmbPanic on a byte outside ACGTacgt, as SeqVector::from would panic."""
        self.ctx._ck(self.ctx._lib.kmb_batch_repack(self.ctx._h, int(strict)))
        self.packed = True
        return self

    def push_chars(self, bases) -> "ReadBatch":
        """SeqVector::push_chars (seq_vector.rs:141-161): append ASCII bases to a one-sequence packed batch."""
        b = np.ascontiguousarray(np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else bases, dtype=np.uint8)
        self.ctx._ck(self.ctx._lib.kmb_packed_push_chars(self.ctx._h, _ptr(b), b.size))
        self.fixed_len += b.size
        self.n_bytes = (self.fixed_len + 31) // 32 * 8
        return self

    def slice(self, read: int, start: int, length: int) -> "ReadBatch":
        """SeqVector::slice / SeqVectorSlice (seq_vector.rs:24-90): the batch becomes a view of bases [start, start + length)
        of `read` (no copy) until `unslice()`; extraction, minimizers, get_kmers then work on the view."""
        self.ctx._ck(self.ctx._lib.kmb_batch_slice(self.ctx._h, read, start, length))
        if not hasattr(self, "_parent"):
            self._parent = (self.n_reads, self.fixed_len, self.ragged)
        self.n_reads, self.fixed_len, self.ragged = 1, 0, True
        return self

    def unslice(self) -> "ReadBatch":
        self.ctx._ck(self.ctx._lib.kmb_batch_unslice(self.ctx._h))
        if hasattr(self, "_parent"):
            self.n_reads, self.fixed_len, self.ragged = self._parent
            del self._parent
        return self

    def get_kmers(self, k: int, pos, reads=None) -> np.ndarray:
        """SeqVector::get_kmer_u64 at (read, pos) pairs of a packed batch (kmb_packed_get_kmers)."""
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        rd = None if reads is None else np.ascontiguousarray(reads, dtype=np.uint64)
        out = np.empty(pos.size, dtype=np.uint64)
        self.ctx._ck(self.ctx._lib.kmb_packed_get_kmers(self.ctx._h, k, _ptr(rd), _ptr(pos), pos.size, _ptr(out)))
        return out

    def minimizers(self, k: int, w: int, hash_k=None, *, validate: bool = True, to: str = "host"):
        """SeqVecMinimizerIter in batch (kmb_minimizers): (lmer word, position in read) of the leftmost minimum-LexHash
        w-mer of every k-mer window; dense slots.  Returns (mmer u64, pos u32)."""
        hash_k = w if hash_k is None else hash_k
        n = self.num_slots(k)
        if to == "device":
            t = _torch()
            mm, pos = t.empty(n, dtype=t.int64, device="cuda"), t.empty(n, dtype=t.int32, device="cuda")
        else:
            mm, pos = np.empty(n, dtype=np.uint64), np.empty(n, dtype=np.uint32)
        flags = 0 if validate else nv.F_NO_VALIDATE
        self.ctx._order_in(mm, pos)
        self.ctx._ck(self.ctx._lib.kmb_minimizers(self.ctx._h, k, w, hash_k, flags, _ptr(mm), _ptr(pos)))
        self.ctx.sync()
        return mm, pos

    def extract_compact(self, k: int, *, validate: bool = True, to: str = "host"):
        """Iterator-identical output (kmb_extract_compact): only the k-mers CanonicalKmerIterator emits, in order.
        Returns dict(pos int32, canon, hash, emit_offsets (n_reads + 1), n)."""
        flags = 0 if validate else nv.F_NO_VALIDATE
        n = C.c_uint64()
        self.ctx._ck(self.ctx._lib.kmb_extract_compact(self.ctx._h, k, flags, None, None, None, None, 0, C.byref(n)))
        m = int(n.value)
        if to == "device":
            t = _torch()
            canon, hsh = t.empty(m, dtype=t.int64, device="cuda"), t.empty(m, dtype=t.int64, device="cuda")
            pos = t.empty(m, dtype=t.int32, device="cuda")
            offs = t.empty(self.n_reads + 1, dtype=t.int64, device="cuda")
        else:
            canon, hsh = np.empty(m, dtype=np.uint64), np.empty(m, dtype=np.uint64)
            pos = np.empty(m, dtype=np.int32)
            offs = np.empty(self.n_reads + 1, dtype=np.uint64)
        self.ctx._order_in(canon, hsh, pos, offs)
        self.ctx._ck(self.ctx._lib.kmb_extract_compact(self.ctx._h, k, flags, _ptr(canon), _ptr(hsh), _ptr(pos), _ptr(offs), m,
                                                       C.byref(n)))
        return dict(pos=pos, canon=canon, hash=hsh, emit_offsets=offs, n=int(n.value))

    def extract_canonical_wide(self, k: int, enc: int = nv.ENC_ACGT, *, want_hash: bool = True, digest: bool = False,
                               validate: bool = True, to: str = "device") -> CanonicalKmers:
        """EXTENSION: 1 <= k <= 64, two u64 words per slot (kmb_extract_canonical_wide)."""
        n = self.num_slots(k)
        out = CanonicalKmers(k=k, n_slots=n, words_per_kmer=2)
        out.canon = self._alloc(2 * n, to)
        out.hash = self._alloc(2 * n, to) if want_hash else None
        d = Digest()
        flags = 0 if validate else nv.F_NO_VALIDATE
        self.ctx._order_in(out.canon, out.hash)
        self.ctx._ck(self.ctx._lib.kmb_extract_canonical_wide(self.ctx._h, k, enc, flags, _ptr(out.canon), _ptr(out.hash),
                                                              C.byref(d) if digest else None))
        self.ctx._order_out(out.canon, out.hash)
        out.digest = d.astuple() if digest else None
        return out

    def histogram(self, k: int, hist_bits: int, *, hist=None, accumulate: bool = False, digest: bool = True,
                  validate: bool = True, to: str = "device", digest_in_hist: bool = False):
        """Fused extraction -> LexHash-prefix histogram + digest, nothing materialised (kmb_histogram).

        digest_in_hist: `hist` has 2^hist_bits + 3 words and the digest is accumulated into the last three on the
        device (KMB_F_DIGEST_IN_HIST): one buffer to all-reduce in place, no read-back, the call stays asynchronous."""
        n = (1 << hist_bits) + (3 if digest_in_hist else 0)
        if hist is None:
            hist = self._alloc(n, to)
            accumulate = False
        d = Digest()
        flags = (0 if validate else nv.F_NO_VALIDATE) | (nv.F_DIGEST_IN_HIST if digest_in_hist else 0)
        want = digest and not digest_in_hist
        self.ctx._order_in(hist)
        self.ctx._ck(self.ctx._lib.kmb_histogram(self.ctx._h, k, flags, hist_bits, _ptr(hist), int(accumulate),
                                                 C.byref(d) if want else None))
        self.ctx._order_out(hist)
        return hist, (d.astuple() if want else None)

    def pack(self, enc: int = nv.ENC_ACGT, word_bits: int = 64, to: str = "host"):
        """Encoding::encode of every read (kmb_pack).  Returns (byte image, word offsets or None)."""
        n = C.c_uint64()
        self.ctx._ck(self.ctx._lib.kmb_pack_num_words(self.ctx._h, word_bits, C.byref(n)))
        nbytes = int(n.value) * word_bits // 8
        if to == "device":
            t = _torch()
            out = t.empty(nbytes, dtype=t.uint8, device="cuda")
        else:
            out = np.empty(nbytes, dtype=np.uint8)
        woff = np.empty(self.n_reads + 1, dtype=np.uint64) if self.ragged else None
        self.ctx._order_in(out)
        self.ctx._ck(self.ctx._lib.kmb_pack(self.ctx._h, enc, word_bits, _ptr(out), _ptr(woff)))
        return out, woff


def host_pack(bases, bits: Optional[np.ndarray] = None, inv: Optional[np.ndarray] = None):
    """ASCII bases -> (bits uint32, inv uint16), 16 bases per entry: the flat 2-bit + validity staging format
    (kmb_host_pack; host SIMD code, needs no GPU).  `bits` / `inv` may be preallocated (e.g. pinned) arrays."""
    lib = nv.lib()
    b = np.ascontiguousarray(np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray)) else bases,
                             dtype=np.uint8)
    nw = (b.size + 15) // 16
    bits = np.empty(nw, dtype=np.uint32) if bits is None else bits
    inv = np.empty(nw, dtype=np.uint16) if inv is None else inv
    assert bits.size >= nw and inv.size >= nw
    check(None, lib.kmb_host_pack(_ptr(b), b.size, _ptr(bits), _ptr(inv)))
    return bits, inv


def host_read_probe(buf: np.ndarray, n_threads: int) -> float:
    """Seconds n_threads host threads need to read `buf` once (kmb_host_read_probe): this host's memory-read floor."""
    sec = C.c_double()
    check(None, nv.lib().kmb_host_read_probe(_ptr(buf), buf.nbytes, n_threads, C.byref(sec)))
    return float(sec.value)


def host_pack_isa() -> str:
    return nv.lib().kmb_host_pack_isa().decode()


def parse_fastx(text: bytes):
    """FASTA / FASTQ text -> (bases uint8, offsets uint64) on the host (kmb_parse_fastx; needs no GPU)."""
    lib = nv.lib()
    nr, nb = C.c_uint64(), C.c_uint64()
    check(None, lib.kmb_parse_fastx(text, len(text), None, 0, None, 0, C.byref(nr), C.byref(nb)))
    bases = np.empty(int(nb.value), dtype=np.uint8)
    offs = np.empty(int(nr.value) + 1, dtype=np.uint64)
    check(None, lib.kmb_parse_fastx(text, len(text), bases.ctypes.data, bases.size, offs.ctypes.data, offs.size, None, None))
    return bases, offs
