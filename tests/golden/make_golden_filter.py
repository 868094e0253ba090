"""Golden fixtures of the candidate filter from the UNMODIFIED reference binary `fastq-extractor`
(oracle/_ref/fastq-extractor, compiled from /root/reference/FastqExtractor.cpp by oracle/Makefile): which read pairs it keeps.
Runs only where /root/reference exists.  Usage: python tests/golden/make_golden_filter.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import workloads as W  # noqa: E402
from t1k_b200 import synth  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "fastq-extractor")
ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def mixed_reads(records, n, read_len, seed, paired):
    """true reads of the set (substitutions, a few N and indels), partial reads (half random), random reads, low-complexity
    reads, reads shorter than k — as variable-length byte strings"""
    rng = np.random.default_rng(seed)
    r1, r2 = W.reads_for(records, n, read_len=read_len, seed=seed, err=0.02, n_rate=0.004, indel_rate=0.1,
                         insert=(read_len + 40, read_len + 220), single_end=not paired)
    a = [bytes(r.tobytes()) for r in r1]
    b = [bytes(r.tobytes()) for r in r2] if paired else None

    def rnd(L):
        return ALPHA[rng.integers(0, 4, size=L)].tobytes()

    for i in range(n):
        m = i % 7
        if m == 1:                                  # unrelated pair
            a[i] = rnd(read_len)
            if paired:
                b[i] = rnd(read_len)
        elif m == 2:                                # first mate random, second true (pair kept through mate 2)
            a[i] = rnd(read_len)
        elif m == 3:                                # half of the read is foreign
            cut = int(rng.integers(20, read_len - 20))
            a[i] = a[i][:cut] + rnd(read_len - cut)
            if paired:
                b[i] = rnd(read_len)
        elif m == 4:                                # heavy substitution noise
            x = bytearray(a[i])
            for _ in range(int(rng.integers(5, 40))):
                x[int(rng.integers(0, len(x)))] = b"ACGT"[int(rng.integers(0, 4))]
            a[i] = bytes(x)
            if paired:
                b[i] = rnd(read_len)
        elif m == 5 and i % 3 == 0:                 # low complexity / short
            a[i] = [b"A" * read_len, b"AC" * (read_len // 2), a[i][:8], a[i][:int(rng.integers(12, 40))]][i % 4]
            if paired:
                b[i] = rnd(read_len)
    return a, b


def write_fastq(path, reads):
    with open(path, "wb") as f:
        for i, s in enumerate(reads):
            f.write(b"@r%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")


def kept_ids(path):
    out = set()
    with open(path, "rb") as f:
        for k, line in enumerate(f):
            if k % 4 == 0:
                out.add(int(line[2:].split()[0].split(b"/")[0]))
    return out


CASES = {
    # name: (records factory, paired, n, read_len, similarity, seed)
    "rna_pe": (lambda: W.small_rna_ref(seed=41), True, 700, 100, 0.8, 51),
    "dna_se": (lambda: W.small_dna_ref(seed=42), False, 700, 75, 0.8, 52),
    "rna_pe_s95_150": (lambda: synth.make_hla_rna_ref(genes=[("HLA-A", 900), ("HLA-B", 700), ("HLA-C", 400)], seed=43), True, 500, 150, 0.95, 53),
    # 4.4 Mbases => k = 13 (beyond the direct-address table of the oracle); the reference is regenerated from its recipe
    "recipe_k13": (lambda: big_ref(), True, 300, 150, 0.8, 54),
}


def big_ref():
    return synth.make_hla_rna_ref(genes=[("HLA-A", 1500), ("HLA-B", 1500), ("HLA-C", 1000)], seed=44)


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    for name, (factory, paired, n, rl, sim, seed) in CASES.items():
        recs = factory()
        a, b = mixed_reads([(x[0], x[1], x[2]) for x in recs], n, rl, seed, paired)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "ref.fa")
            synth.write_fasta(fa, recs)
            write_fastq(os.path.join(td, "r1.fq"), a)
            cmd = [EXE, "-f", fa, "-o", os.path.join(td, "out"), "-s", str(sim), "-t", "1"]
            if paired:
                write_fastq(os.path.join(td, "r2.fq"), b)
                cmd += ["-1", os.path.join(td, "r1.fq"), "-2", os.path.join(td, "r2.fq")]
            else:
                cmd += ["-u", os.path.join(td, "r1.fq")]
            subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
            kept = kept_ids(os.path.join(td, "out_1.fq" if paired else "out.fq"))
        flags = np.zeros(n, dtype=np.uint8)
        flags[sorted(kept)] = 1
        total = sum(len(r[2]) for r in recs)
        if name.startswith("recipe_"):
            np.savez_compressed(os.path.join(HERE, "filter", name + ".npz"), fasta=np.zeros(0, np.uint8), paired=int(paired), similarity=sim,
                                reads1=np.asarray(a, dtype=object).astype("S"), reads2=np.asarray(b, dtype=object).astype("S"), kept=flags)
            print(name, "sequences", len(recs), "bases", total, "pairs", n, "kept", int(flags.sum()))
            continue
        np.savez_compressed(os.path.join(HERE, "filter", name + ".npz"),
                            fasta=np.frombuffer(b"".join(b">" + r[0].encode() + (b" " + r[1].encode() if r[1] else b"") + b"\n" + r[2] + b"\n" for r in recs), dtype=np.uint8),
                            paired=int(paired), similarity=sim, reads1=np.asarray(a, dtype=object).astype("S"),
                            reads2=np.asarray(b, dtype=object).astype("S") if paired else np.zeros(0, "S1"), kept=flags)
        print(name, "sequences", len(recs), "bases", total, "pairs", n, "kept", int(flags.sum()))


if __name__ == "__main__":
    main()
