"""The host side of the pinned staging path (kmb_host_pack): ASCII -> 2 bits + 1 validity bit per base.  Pure host
code, runs without a GPU.  Checked against the oracle's restatement of SeqVector::from (naive_impl/seq_vector.rs:230-242),
of encode_binary_u8's valid set (naive_impl/mod.rs:40-50) and of Encoding::encode's (c >> 1) & 3 (encoding/naive.rs:14-16)."""
import numpy as np
import pytest

import oracle as ko


@pytest.fixture(scope="module")
def kb():
    import __graft_entry__ as g
    g.build()
    import kmers_b200
    return kmers_b200


def _reference(b: np.ndarray):
    n = b.size
    nw = (n + 15) // 16
    x = (b >> 1) & 3
    code = (x ^ (x >> 1)).astype(np.uint32)
    u = b & 0xDF
    bad = ~((u == 65) | (u == 67) | (u == 71) | (u == 84))
    pad = nw * 16 - n
    code = np.concatenate([code, np.zeros(pad, dtype=np.uint32)]).reshape(nw, 16)
    bad = np.concatenate([bad, np.ones(pad, dtype=bool)]).reshape(nw, 16)
    bits = np.zeros(nw, dtype=np.uint32)
    inv = np.zeros(nw, dtype=np.uint16)
    for j in range(16):
        bits |= code[:, j] << np.uint32(2 * j)
        inv |= bad[:, j].astype(np.uint16) << np.uint16(j)
    return bits, inv


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 129, 1000, 4097, 100003])
def test_every_length_and_every_byte(kb, n):
    rng = np.random.default_rng(n)
    for b in (rng.integers(0, 256, size=n, dtype=np.uint8),
              np.frombuffer(bytes(rng.choice(list(b"ACGTacgtNnRYKM\n"), size=n).astype(np.uint8)), dtype=np.uint8)):
        bits, inv = kb.host_pack(b)
        rb, ri = _reference(b)
        assert np.array_equal(bits, rb) and np.array_equal(inv, ri)


def test_all_256_byte_values(kb):
    b = np.arange(256, dtype=np.uint8)
    bits, inv = kb.host_pack(b)
    valid = {ord(c) for c in "ACGTacgt"}
    for c in range(256):
        assert ((int(inv[c // 16]) >> (c % 16)) & 1) == (0 if c in valid else 1)
        x = (c >> 1) & 3
        assert ((int(bits[c // 16]) >> (2 * (c % 16))) & 3) == (x ^ (x >> 1))


def test_matches_seqvector_words(kb):
    """On pure ACGT input the packed stream IS SeqVector::from's word image (32 bases per u64, base i at bits 2i+1:2i)."""
    seq = ko.generate_bases(7, 0, 4096).tobytes()
    bits, inv = kb.host_pack(seq)
    assert not inv.any()
    assert np.array_equal(bits.view(np.uint64), ko.sv_from_bytes(seq)[: len(seq) // 32])


def test_isa_is_reported(kb):
    assert kb.host_pack_isa() in ("avx512gfni", "avx512bw", "avx2", "swar")


@pytest.mark.parametrize("isa", ["swar", "avx2", "avx512bw", "avx512gfni"])
def test_every_simd_variant_agrees(kb, isa):
    """Each implementation (capped through KMB_HOST_PACK_ISA in a fresh process) gives the reference bytes."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import kmers_b200 as kb\n"
            "rng = np.random.default_rng(5); b = rng.integers(0, 256, size=100003, dtype=np.uint8)\n"
            "b[1000:90000] = np.frombuffer(bytes(rng.choice(list(b'ACGTacgtN'), size=89000).astype(np.uint8)), dtype=np.uint8)\n"
            "bits, inv = kb.host_pack(b); print(kb.host_pack_isa(), int(bits.astype(np.uint64).sum()), int(inv.astype(np.uint64).sum()))\n" % root)
    env = dict(os.environ, KMB_HOST_PACK_ISA=isa)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split()
    rng = np.random.default_rng(5)
    b = rng.integers(0, 256, size=100003, dtype=np.uint8)
    b[1000:90000] = np.frombuffer(bytes(rng.choice(list(b'ACGTacgtN'), size=89000).astype(np.uint8)), dtype=np.uint8)
    rb, ri = _reference(b)
    assert (int(out[1]), int(out[2])) == (int(rb.astype(np.uint64).sum()), int(ri.astype(np.uint64).sum()))
    # the variant that ran is the requested one, or the best this CPU has below it
    assert out[0] in {"swar": ("swar",), "avx2": ("avx2", "swar"), "avx512bw": ("avx512bw", "avx2", "swar"),
                      "avx512gfni": ("avx512gfni", "avx512bw", "avx2", "swar")}[isa]


def _aligned(n, dtype, align=64):
    raw = np.empty(n * np.dtype(dtype).itemsize + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * np.dtype(dtype).itemsize].view(dtype)


@pytest.mark.parametrize("n", [65536, 65536 + 64, 300_007, 1 << 20])
def test_aligned_and_unaligned_destinations(kb, n):
    """Cache-line aligned outputs of >= 64 Ki bases (the pipeline's pinned staging rings: the interleaved-stream path) and
    destinations 4 bytes off a line give the reference bytes."""
    rng = np.random.default_rng(n)
    b = np.frombuffer(bytes(rng.choice(list(b"ACGTacgtNRY\n"), size=n, p=[.2, .2, .2, .2, .04, .04, .04, .04, .01, .01, .01, .01]).astype(np.uint8)), dtype=np.uint8)
    nw = (n + 15) // 16
    rb, ri = _reference(b)
    bits, inv = kb.host_pack(b, _aligned(nw, np.uint32), _aligned(nw, np.uint16))
    assert bits.ctypes.data % 64 == 0 and inv.ctypes.data % 64 == 0
    assert np.array_equal(bits, rb) and np.array_equal(inv, ri)
    ub = _aligned(nw + 1, np.uint32)[1:]  # 4 bytes off a line
    bits2, inv2 = kb.host_pack(b, ub, _aligned(nw + 1, np.uint16)[1:])
    assert np.array_equal(bits2, rb) and np.array_equal(inv2, ri)
