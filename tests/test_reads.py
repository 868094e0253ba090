"""SURVEY.md 8f N2, first half — t1k_reads_load: FASTA / FASTQ (.gz) into the strided buffers of t1k_genotype, following the
reference's ReadFiles / kseq grammar (ReadFiles.hpp:155-204).  CPU: against the expected sequences of hand-made files (multi-line
FASTA, CRLF, empty lines, quality strings that start with '@' or '+', a missing final newline, gzip, truncated quality) and,
where the reference harness is present, against the UNMODIFIED reference reader on the same files (oracle/_ref/ref_harness reads).
No device needed."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O
from t1k_b200 import _lib as L
from t1k_b200.reads import load_reads


def seqs_of(arr):
    return [bytes(r[:int(np.argmax(r == 0))]) if (r == 0).any() else bytes(r) for r in arr]


def ref_reader(path):
    if not os.path.exists(O.REF_HARNESS):
        return None
    out = subprocess.run([O.REF_HARNESS, "reads", path], stdout=subprocess.PIPE, check=True).stdout
    return out.split(b"\n")[:-1] if out else []


def rand_reads(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGTN", dtype=np.uint8)
    return [alpha[rng.choice(5, size=int(rng.integers(lo, hi + 1)), p=[.245, .245, .245, .245, .02])].tobytes() for _ in range(n)]


def fastq(reads, qual_char=b"I", crlf=False, final_newline=True):
    nl = b"\r\n" if crlf else b"\n"
    rec = []
    for i, s in enumerate(reads):
        q = qual_char * len(s)
        if i % 7 == 3 and len(s) > 2:
            q = b"@" + q[1:]                 # a quality string that looks like a header
        if i % 11 == 5 and len(s) > 2:
            q = b"+" + q[1:]
        rec.append(b"@r%d/1 comment %d" % (i, i) + nl + s + nl + b"+" + nl + q)
    body = nl.join(rec)
    return body + (nl if final_newline else b"")


def fasta(reads, width=60):
    out = []
    for i, s in enumerate(reads):
        out.append(b">s%d desc" % i)
        if i % 5 == 2:
            out.append(b"")                   # empty line inside a record
        out += [s[k:k + width] for k in range(0, len(s), width)]
    return b"\n".join(out) + b"\n"


CASES = {
    "fastq": lambda r: fastq(r),
    "fastq_crlf": lambda r: fastq(r, crlf=True),
    "fastq_no_final_newline": lambda r: fastq(r, final_newline=False),
    "fasta_multiline": lambda r: fasta(r, 37),
    "fasta_long_lines": lambda r: fasta(r, 100000),
    "leading_garbage": lambda r: b"# not a record\n\n" + fastq(r),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("gz", [False, True])
def test_reader_matches_expected_and_reference(tmp_path, name, gz):
    reads = rand_reads(257, 20, 151, seed=len(name))
    data = CASES[name](reads)
    p = str(tmp_path / (name + (".fq.gz" if gz else ".fq")))
    with (gzip.open(p, "wb") if gz else open(p, "wb")) as f:
        f.write(data)
    a, b = load_reads(p)
    assert b is None and a.shape == (len(reads), max(len(s) for s in reads) + 1)
    got = seqs_of(a)
    assert got == reads
    want = ref_reader(p)
    if want is not None:
        assert got == want


def test_pairs_edge_files_and_errors(tmp_path):
    r1, r2 = rand_reads(1000, 100, 100, seed=1), rand_reads(1000, 90, 120, seed=2)
    p1, p2 = str(tmp_path / "a_1.fq"), str(tmp_path / "a_2.fq.gz")
    with open(p1, "wb") as f:
        f.write(fastq(r1))
    with gzip.open(p2, "wb") as f:
        f.write(fastq(r2))
    a, b = load_reads(p1, p2)
    assert a.shape == b.shape == (1000, 121) and seqs_of(a) == r1 and seqs_of(b) == r2
    # the mate file is shorter: an error, not silently reused data
    p3 = str(tmp_path / "short_2.fq")
    with open(p3, "wb") as f:
        f.write(fastq(r2[:10]))
    with pytest.raises(L.T1KError):
        load_reads(p1, p3)
    with pytest.raises(L.T1KError):
        load_reads(str(tmp_path / "missing.fq"))
    # an empty file
    pe = str(tmp_path / "empty.fq")
    open(pe, "wb").close()
    a, _ = load_reads(pe)
    assert a.shape[0] == 0
    # truncated quality: kseq reports -2 and ReadFiles::Next ends the file there; the records before it are kept
    pt = str(tmp_path / "trunc.fq")
    with open(pt, "wb") as f:
        f.write(fastq(r1[:5]) + b"@last\nACGTACGT\n+\nIII\n")
    a, _ = load_reads(pt)
    got, want = seqs_of(a), ref_reader(pt)
    assert got == r1[:5]
    if want is not None:
        assert got == want
    # a read longer than T1K_MAX_READ_LEN is refused
    pl = str(tmp_path / "long.fa")
    with open(pl, "wb") as f:
        f.write(b">x\n" + b"A" * 1001 + b"\n")
    with pytest.raises(L.T1KError):
        load_reads(pl)
