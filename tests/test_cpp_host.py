"""The C++ host-side mirror (kmers_b200/cpp/kmers_b200.hpp): compiles against include/kmers_b200.h, links the
C-ABI library, and (on a GPU box) passes the reference-style checks in tests/cpp/host_mirror_test.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import __graft_entry__ as g
    g.build()
    exe = str(tmp_path / "host_mirror_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I",
                           os.path.join(ROOT, "kmers_b200", "cpp"), os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
                           "-L", os.path.join(ROOT, "kmers_b200"), "-lkmers_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "kmers_b200"), "-o", exe])
    return exe


def test_cpp_mirror_compiles_links_and_refuses_cpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, r.stdout + r.stderr  # KMB_ERR_NO_DEVICE, loudly; no CPU fallback
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_style_checks(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
