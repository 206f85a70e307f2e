"""examples/fastq_to_kmers.py runs end to end (FASTQ text -> compacted canonical k-mers + minimizers through the ragged path) and
its first printed k-mer is what the restated iterator yields."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fastq_example_runs(tmp_path):
    import ctypes as C
    import oracle as ko
    rng = np.random.default_rng(3)
    letters = np.frombuffer(b"ACGTN", dtype=np.uint8)
    reads = [letters[rng.choice(5, size=int(n), p=[.2495, .2495, .2495, .2495, .002])].tobytes() for n in rng.integers(40, 151, size=3000)]
    fq = tmp_path / "reads.fastq"
    fq.write_bytes(b"".join(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)) for i, s in enumerate(reads)))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "fastq_to_kmers.py"), str(fq), "31"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    m = re.search(r"(\d+) reads, (\d+) bases, k=31: (\d+) canonical k-mers", p.stdout)
    assert m and int(m.group(1)) == len(reads) and int(m.group(2)) == sum(len(s) for s in reads)
    # the count and read 0's first k-mer against the restated iterator
    L = ko.lib()
    total, first = 0, None
    for i, s in enumerate(reads):
        it = ko.Iter(s, 31)
        while not it.exhausted():
            if i == 0 and first is None:
                km = it.km
                first = (it.pos, L.ko_ck_get_canonical_word(C.byref(km)))
            total += 1
            it.inc()
    assert int(m.group(3)) == total
    if first is not None:
        f = re.search(r"first: pos=(\d+) canon=(0x[0-9a-f]+)", p.stdout)
        assert f and int(f.group(1)) == first[0] and int(f.group(2), 16) == first[1]
