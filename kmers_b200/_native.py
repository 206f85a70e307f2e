"""ctypes binding of the C ABI in include/kmers_b200.h (libkmers_b200.so).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).
There is no CPU fallback anywhere in this package: if the shared library is
missing, or no CUDA device is visible, the first use raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("KMERS_B200_SO") or os.path.join(_HERE, "libkmers_b200.so")  # env override: kernel experiments

OK = 0
ERR_INVALID_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE, ERR_PANIC, ERR_NOMEM = -1, -2, -3, -4, -5, -6
SENTINEL = 0xFFFFFFFFFFFFFFFF
ENC_ACGT, ENC_ACTG, ENC_XOR10 = 0x1E, 0x1B, 0x100
F_NO_VALIDATE = 0x1
F_DIGEST_IN_HIST = 0x2
NO_MATCH, IDENTITY_MATCH, TWIN_MATCH = 0, 1, 2


class KmbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"kmers_b200 error {code}: {msg}")
        self.code = code


class KmbPanic(KmbError):
    """The reference would panic!/assert! on these arguments."""


class Digest(C.Structure):
    _fields_ = [("n_valid", C.c_uint64), ("checksum_canon", C.c_uint64), ("checksum_hash", C.c_uint64)]

    def astuple(self):
        return (int(self.n_valid), int(self.checksum_canon), int(self.checksum_hash))


# name -> (restype, argtypes); mirrors include/kmers_b200.h one to one
_vp, _u64, _u32, _i32, _sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_size_t
_pp = C.POINTER(C.c_void_p)
_pu64 = C.POINTER(C.c_uint64)
_pd = C.POINTER(Digest)
SIGNATURES = {
    "kmb_version": (_i32, []),
    "kmb_device_count": (_i32, []),
    "kmb_ctx_create": (_i32, [_i32, _vp, _pp]),
    "kmb_ctx_destroy": (_i32, [_vp]),
    "kmb_last_error": (C.c_char_p, [_vp]),
    "kmb_ctx_sync": (_i32, [_vp]),
    "kmb_ctx_stream": (_vp, [_vp]),
    "kmb_ctx_launch_count": (_u64, [_vp]),
    "kmb_device_alloc": (_i32, [_vp, _sz, _pp]),
    "kmb_device_free": (_i32, [_vp, _vp]),
    "kmb_host_alloc_pinned": (_i32, [_vp, _sz, _pp]),
    "kmb_host_free_pinned": (_i32, [_vp, _vp]),
    "kmb_memcpy": (_i32, [_vp, _vp, _vp, _sz]),
    "kmb_batch_upload": (_i32, [_vp, _vp, _u64, _vp, _u64, _u64]),
    "kmb_batch_attach": (_i32, [_vp, _vp, _u64, _vp, _u64, _u64]),
    "kmb_batch_generate": (_i32, [_vp, _u64, _u64, _u64, _u64, _u32]),
    "kmb_batch_download": (_i32, [_vp, _vp, _u64]),
    "kmb_batch_info": (_i32, [_vp, _pu64, _pu64, _pu64]),
    "kmb_batch_num_slots": (_i32, [_vp, _u32, _pu64]),
    "kmb_batch_window_offsets": (_i32, [_vp, _u32, _vp]),
    "kmb_extract_canonical": (_i32, [_vp, _u32, _u32, _vp, _vp, _vp, _vp, _pd]),
    "kmb_extract_compact": (_i32, [_vp, _u32, _u32, _vp, _vp, _vp, _vp, _u64, _pu64]),
    "kmb_extract_canonical_wide": (_i32, [_vp, _u32, _i32, _u32, _vp, _vp, _pd]),
    "kmb_histogram": (_i32, [_vp, _u32, _u32, _u32, _vp, _i32, _pd]),
    "kmb_extract_canonical_host": (_i32, [_vp, _vp, _u64, _u64, _u32, _u32, _vp, _vp, _pd]),
    "kmb_extract_canonical_host_packed": (_i32, [_vp, _vp, _vp, _u64, _u64, _u32, _u32, _vp, _vp, _pd]),
    "kmb_host_pack": (_i32, [_vp, _u64, _vp, _vp]),
    "kmb_host_pack_isa": (C.c_char_p, []),
    "kmb_host_read_probe": (_i32, [_vp, _u64, _u32, C.POINTER(C.c_double)]),
    "kmb_ctx_set_host_threads": (_i32, [_vp, _u32]),
    "kmb_ctx_host_stats": (_i32, [_vp, _pu64]),
    "kmb_minimizers": (_i32, [_vp, _u32, _u32, _u32, _u32, _vp, _vp]),
    "kmb_minimizer_words": (_i32, [_vp, _u32, _u32, _u32, _vp, _u64, _vp, _vp]),
    "kmb_batch_repack": (_i32, [_vp, _i32]),
    "kmb_batch_attach_packed": (_i32, [_vp, _vp, _u64, _vp, _vp, _u64, _u64]),
    "kmb_packed_get_kmers": (_i32, [_vp, _u32, _vp, _vp, _u64, _vp]),
    "kmb_batch_new_packed": (_i32, [_vp, _u64]),
    "kmb_packed_push_chars": (_i32, [_vp, _vp, _u64]),
    "kmb_batch_slice": (_i32, [_vp, _u64, _u64, _u64]),
    "kmb_batch_unslice": (_i32, [_vp]),
    "kmb_allreduce_u64": (_i32, [_vp, _i32, _vp, _u64]),
    "kmb_parse_fastx": (_i32, [C.c_char_p, _u64, _vp, _u64, _vp, _u64, _pu64, _pu64]),
    "kmb_batch_ingest_fastx": (_i32, [_vp, C.c_char_p, _u64, _pu64, _pu64]),
    "kmb_pack": (_i32, [_vp, _i32, _u32, _vp, _vp]),
    "kmb_pack_num_words": (_i32, [_vp, _u32, _pu64]),
    "kmb_words_to_strings": (_i32, [_vp, _u32, _vp, _u64, _vp]),
    "kmb_unpack": (_i32, [_vp, _i32, _u32, _vp, _u64, _u32, _u32, _vp]),
    "kmb_revcomp_words": (_i32, [_vp, _i32, _u32, _u32, _u32, _vp, _vp, _u64]),
    "kmb_reverse_complement_words": (_i32, [_vp, _u32, _vp, _vp, _u64]),
    "kmb_canonical_words": (_i32, [_vp, _u32, _vp, _vp, _vp, _u64]),
    "kmb_lexhash_words": (_i32, [_vp, _u32, _vp, _vp, _u64]),
    "kmb_match_words": (_i32, [_vp, _u32, _vp, _vp, _vp, _u64]),
    "kmb_sub_kmer_words": (_i32, [_vp, _u32, _u32, _u32, _vp, _vp, _u64]),
    "kmb_append_base_words": (_i32, [_vp, _u32, _vp, _vp, _i32, _vp, _vp, _u64]),
    "kmb_prepend_base_words": (_i32, [_vp, _u32, _vp, _vp, _i32, _vp, _vp, _u64]),
    "kmb_canonical_append_base_words": (_i32, [_vp, _u32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _u64]),
    "kmb_canonical_prepend_base_words": (_i32, [_vp, _u32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _u64]),
    "kmb_is_fw_canonical_words": (_i32, [_vp, _vp, _vp, _vp, _u64]),
    "kmb_kmer_get": (_i32, [_vp, _u32, _u32, _vp, _u64, _u32, _vp]),
    "kmb_kmer_get_prefix": (_i32, [_vp, _u32, _u32, _vp, _u64, _u32, _vp]),
    "kmb_bitmer_to_bytes": (_i32, [_vp, _u32, _vp, _u64, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libkmers_b200.so; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). kmers_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)  # AttributeError if the ABI and the header drift apart
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(ctx_handle, code: int) -> None:
    if code == OK:
        return
    msg = lib().kmb_last_error(ctx_handle)
    msg = msg.decode(errors="replace") if msg else ""
    raise (KmbPanic if code == ERR_PANIC else KmbError)(code, msg)
