"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/kmers_b200.h declares, and refuses to compute without a
GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle as ko

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as g
    g.build()
    from kmers_b200 import _native
    return _native


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "kmers_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(kmb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_all_exported(native):
    L = native.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/kmers_b200.h but not exported"
    assert sorted(native.SIGNATURES) == names, "ctypes signature table and header drifted apart"
    assert L.kmb_version() == 100


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, "include", "kmers_b200.h")).read()
    assert "torch" not in hdr.lower().replace("share a stream with torch", "")
    assert "at::" not in hdr and "c10::" not in hdr
    assert 'extern "C"' in hdr


def test_fails_loudly_without_gpu(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = native.lib()
    assert L.kmb_device_count() == 0
    h = C.c_void_p()
    assert L.kmb_ctx_create(0, None, C.byref(h)) == native.ERR_NO_DEVICE
    assert b"no CPU fallback" in L.kmb_last_error(None)
    import kmers_b200 as kb
    with pytest.raises(kb.KmbError):
        kb.Context(0)


def test_null_ctx_is_an_error_not_a_crash(native):
    L = native.lib()
    assert L.kmb_ctx_sync(None) == native.ERR_INVALID_ARG
    assert L.kmb_extract_canonical(None, 31, 0, None, None, None, None, None) == native.ERR_INVALID_ARG
    assert L.kmb_ctx_destroy(None) == native.OK


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "kmers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "kmers_oracle" not in src, f


def test_naive_enum_matches_reference_table():
    import kmers_b200 as kb
    assert {m.name: int(m) for m in kb.Naive} == ko.NAIVE  # encoding/naive.rs:49-74
    assert int(kb.Xor10) == kb.ENC_XOR10
    # kmer.rs:97-118 choose_number_of_word
    for wb, cases in {8: [(1, 1), (4, 1), (5, 2)], 16: [(1, 1), (8, 1), (9, 2)], 32: [(1, 1), (16, 1), (17, 2)],
                      64: [(1, 1), (32, 1), (64, 2)], 128: [(1, 1), (64, 1), (65, 2)]}.items():
        for k, want in cases:
            assert kb.word_for_k(wb, k) == want == ko.lib().ko_word_for_k(wb, k)
    assert [kb.num_bytes(wb, 15) for wb in (8, 16, 32, 64, 128)] == [4, 4, 4, 8, 16]  # kmer.rs:120-153


def test_static_shared_memory_leaves_room_for_the_tile():
    """fixed_kernel / csr_kernel are launched without opting in to more than 48 KiB of shared memory; the host sizes
    their dynamic tile up to 36 KiB (make_fixed_geom) resp. ~27 KiB (make_csr_geom), so their static part must stay
    below 12 KiB -- a launch that exceeds the limit fails with 'invalid argument' only for unusual read geometries."""
    import re
    import shutil
    import subprocess
    import __graft_entry__ as g
    g.build()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-res-usage", g.SO], capture_output=True, text=True).stdout
    names = re.findall(r"Function (\S+):\n\s*REG:\d+ STACK:\d+ SHARED:(\d+)", out)
    checked = 0
    for name, shared in names:
        if "hist_" in name or "compact_" in name:
            continue
        if "12fixed_kernel" in name:
            assert int(shared) <= 12 * 1024, (name, shared)   # + at most 36 KiB of tile
            checked += 1
        elif "10csr_kernel" in name:
            assert int(shared) <= 20 * 1024, (name, shared)   # + ~27 KiB of tile and offset tables
            checked += 1
    assert checked >= 40


def test_rust_sys_matches_the_header():
    """rust/kmers-b200-sys cannot be compiled here (no Rust toolchain), so its `extern "C"` block is GENERATED from the header
    (scripts/gen_rust_sys.py) and this test fails when the committed file differs from a fresh generation: every entry
    point of include/kmers_b200.h is declared, with the same arity and the pointer / integer types mapped one to one."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(ROOT, "scripts", "gen_rust_sys.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    hdr = open(os.path.join(ROOT, "include", "kmers_b200.h")).read()
    cur = open(gen.LIB_RS).read()
    assert gen.render(cur, gen.generate(hdr)) == cur, "run `python scripts/gen_rust_sys.py` after changing include/kmers_b200.h"
    protos = gen.prototypes(hdr)
    assert sorted(p[0] for p in protos) == _declared_symbols()
    block = cur[cur.index(gen.BEGIN):cur.index(gen.END)]
    for name, _, params in protos:
        m = re.search(r"pub fn %s\((.*?)\)" % name, block)
        assert m, name
        assert (len(m.group(1).split(",")) if m.group(1).strip() else 0) == len(params), f"arity of {name}"
    # the safe wrapper only calls functions the sys crate declares
    wrapper = open(os.path.join(ROOT, "rust", "kmers-b200", "src", "lib.rs")).read()
    for used in set(re.findall(r"sys::(kmb_[a-z0-9_]+)\(", wrapper)):
        assert f"pub fn {used}(" in block, f"rust/kmers-b200 calls sys::{used}, which the header does not declare"
