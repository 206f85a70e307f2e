"""Error behaviour of the C ABI on a GPU box: every misuse is a status code + message, never a crash; where the
reference panics the code is KMB_ERR_PANIC (the reference's error model is panic!/assert!)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx():
    import kmers_b200 as kb
    c = kb.Context(0)
    yield c
    c.close()


def test_no_batch_is_a_state_error(ctx):
    import kmers_b200 as kb
    n = C.c_uint64()
    assert ctx._lib.kmb_batch_num_slots(ctx._h, 31, C.byref(n)) == kb._native.ERR_STATE
    assert b"no read batch" in ctx._lib.kmb_last_error(ctx._h)
    assert ctx._lib.kmb_extract_canonical(ctx._h, 31, 0, None, None, None, None, None) == kb._native.ERR_STATE
    assert ctx._lib.kmb_batch_info(ctx._h, None, None, None) == kb._native.ERR_STATE


def test_bad_shapes_and_arguments(ctx):
    import kmers_b200 as kb
    E = kb._native
    b = np.frombuffer(b"ACGT" * 10, dtype=np.uint8)
    L = ctx._lib
    assert L.kmb_batch_upload(ctx._h, b.ctypes.data, 40, None, 3, 10) == E.ERR_INVALID_ARG          # 3 * 10 != 40
    assert L.kmb_batch_upload(ctx._h, b.ctypes.data, 40, None, 4, 0) == E.ERR_INVALID_ARG           # neither offsets nor fixed_len
    offs = np.array([1, 40], dtype=np.uint64)
    assert L.kmb_batch_upload(ctx._h, b.ctypes.data, 40, offs.ctypes.data, 1, 0) == E.ERR_INVALID_ARG  # offsets[0] != 0
    batch = ctx.upload(b, fixed_len=10)
    out = np.zeros(64, dtype=np.uint8)
    assert L.kmb_pack(ctx._h, 0xFF, 64, out.ctypes.data, None) == E.ERR_INVALID_ARG                  # not a permutation of codes
    assert L.kmb_pack(ctx._h, kb.ENC_ACGT, 24, out.ctypes.data, None) == E.ERR_INVALID_ARG           # word_bits
    assert L.kmb_unpack(ctx._h, kb.ENC_ACGT, 64, out.ctypes.data, 1, 1, 33, out.ctypes.data) == E.ERR_INVALID_ARG  # > capacity
    assert L.kmb_revcomp_words(ctx._h, kb.ENC_ACGT, 33, 64, 1, out.ctypes.data, out.ctypes.data, 1) == E.ERR_PANIC   # 2k > bits
    assert L.kmb_revcomp_words(ctx._h, kb.ENC_ACGT, 0, 64, 1, out.ctypes.data, out.ctypes.data, 1) == E.ERR_PANIC
    assert L.kmb_revcomp_words(ctx._h, kb.ENC_ACGT, 5, 128, 3, out.ctypes.data, out.ctypes.data, 1) == E.ERR_INVALID_ARG  # > 256 bit
    assert L.kmb_histogram(ctx._h, 31, 0, 40, out.ctypes.data, 0, None) == E.ERR_INVALID_ARG
    assert L.kmb_extract_canonical_wide(ctx._h, 65, kb.ENC_ACGT, 0, None, None, None) == E.ERR_INVALID_ARG
    assert L.kmb_extract_canonical_wide(ctx._h, 31, 0x11, 0, None, None, None) == E.ERR_INVALID_ARG
    assert L.kmb_packed_get_kmers(ctx._h, 5, None, out.ctypes.data, 1, out.ctypes.data) == E.ERR_STATE  # batch not packed
    # the context is still usable after all of that
    assert batch.extract_canonical(5, digest=True).digest[0] == 4 * 6


def test_two_contexts_are_independent():
    import kmers_b200 as kb
    import oracle as ko
    a, b = kb.Context(0), kb.Context(0)
    try:
        ba = ko.generate_bases(1, 0, 150 * 2000)
        bb = ko.generate_bases(2, 0, 100 * 3000, 3000)
        ra_batch = a.upload(ba, fixed_len=150)
        rb_batch = b.upload(bb, fixed_len=100)
        ra = ra_batch.extract_canonical(31, digest=True, to="host")
        rb = rb_batch.extract_canonical(21, digest=True, to="host")
        ra2 = ra_batch.extract_canonical(31, digest=True, to="host")  # a's batch is untouched by b's work
        wa = ko.extract_canonical(ba, 31, n_reads=2000, fixed_len=150)
        wb = ko.extract_canonical(bb, 21, n_reads=3000, fixed_len=100)
        assert np.array_equal(ra.canon, wa["canon"]) and np.array_equal(rb.canon, wb["canon"])
        assert ra.digest == ra2.digest == (wa["n_valid"], wa["checksum_canon"], wa["checksum_hash"])
        assert a.stream != b.stream
    finally:
        a.close()
        b.close()


def test_borrowed_streams():
    """A context can run on torch's current stream or the legacy default stream (handle 0 -> cudaStreamLegacy)."""
    import torch
    import kmers_b200 as kb
    import oracle as ko
    bases = ko.generate_bases(3, 0, 150 * 1000)
    want = ko.extract_canonical(bases, 31, n_reads=1000, fixed_len=150)
    s = torch.cuda.Stream()
    for handle in (s.cuda_stream, 0):
        with kb.Context(0, stream=handle) as ctx:
            assert ctx.stream == (handle or 1)
            dev = torch.from_numpy(bases).cuda()
            torch.cuda.synchronize()
            res = ctx.attach(dev, fixed_len=150).extract_canonical(31, to="device")
            ctx.sync()
            assert np.array_equal(res.host("canon"), want["canon"])


def test_launch_count_accounting(ctx):
    b = ctx.generate(1, 100, 150)
    n0 = ctx.launch_count
    b.extract_canonical(31)
    assert ctx.launch_count == n0 + 1  # one kernel per extraction of a fixed-length batch
