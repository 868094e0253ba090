"""SURVEY 8f N4 goldens: the alleles that stay in their equivalence class after Genotyper::RemoveLowLikelihoodAlleleInEquivalentClass
(Genotyper.hpp:1371-1460, called at Genotyper.cpp:647) — from the UNMODIFIED reference (oracle/_ref/ref_harness genotype
--ecfilter) on the committed golden workloads.  Writes tests/golden/ecfilter/<name>.npz (kept: uint8 per allele).
Runs only where /root/reference exists.  Usage: python tests/golden/make_golden_ecfilter.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_io as G  # noqa: E402
import oracle_py as O  # noqa: E402
import workloads as W  # noqa: E402
from t1k_b200 import synth  # noqa: E402


def run(recs, r1, r2, sim, relax):
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "ref.fa")
        synth.write_fasta(fa, recs)
        W.write_lines(os.path.join(td, "r1.txt"), r1)
        cmd = [O.REF_HARNESS, "genotype", "-f", fa, "-1", os.path.join(td, "r1.txt"), "-o", os.path.join(td, "out"), "-s", str(sim), "--ecfilter"]
        if r2 is not None:
            W.write_lines(os.path.join(td, "r2.txt"), r2)
            cmd += ["-2", os.path.join(td, "r2.txt")]
        if relax:
            cmd.append("--relaxIntronAlign")
        subprocess.check_call(cmd)
        return O.parse_harness(os.path.join(td, "out"))


def twins_case():
    """Alleles with an uncovered extension: two of every three alleles of a small RNA set get a twin = the same sequence + 200..600 random
    bases that no read covers.  Reads come from the originals only, so a twin shares its original's read groups (one class) but
    is covered over a fraction of its length: the likelihood filter removes it where the class is abundant enough."""
    base = W.small_rna_ref(seed=71)
    rng = np.random.default_rng(72)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    recs = list(base)
    for i, (name, comment, seq) in enumerate(base):
        if i % 3 != 1:
            ext = alpha[rng.integers(0, 4, size=int(rng.integers(200, 600)))].tobytes()
            gene = name.split("*")[0]
            recs.append(("%s*9%d:%02d:01" % (gene, i // 99, i % 99 + 1), "1 0 %d" % (len(seq) + len(ext) - 1), seq + ext))
    r1, r2 = W.reads_for(base, 1200, read_len=100, seed=73, err=0.002, n_rate=0.0, indel_rate=0.0, insert=(180, 320))
    return recs, r1, r2, 0.9, False


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    recs, r1, r2, sim, relax = twins_case()
    H = run(recs, r1, r2, sim, relax)
    kept = np.zeros(H["nAlleles"], dtype=np.uint8)
    kept[H["kept"]] = 1
    ec = np.asarray([q[0] for q in H["q"]], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "ecfilter", "twins.npz"), kept=kept, equivalent_class=ec, reads1=r1, reads2=r2, similarity=sim, relax=int(relax),
                        ec_abundance=np.asarray([q[2] for q in H["q"]], dtype=np.float64),
                        fasta=np.frombuffer(b"".join(b">" + n_.encode() + (b" " + c.encode() if c else b"") + b"\n" + s_ + b"\n" for n_, c, s_ in recs), dtype=np.uint8))
    print("twins alleles", H["nAlleles"], "in a class", int((ec >= 0).sum()), "kept", int(kept.sum()), "removed", int((ec >= 0).sum()) - int(kept.sum()))
    for name in G.names():
        g = G.load(name)
        H = run(g["records"], g["reads1"], g["reads2"], g["similarity"], g["relax"])
        n = H["nAlleles"]
        assert np.array_equal(np.asarray([q[0] for q in H["q"]]), g["q"][:, 0].astype(np.int64)), name     # same classes as the committed golden
        kept = np.zeros(n, dtype=np.uint8)
        kept[H["kept"]] = 1
        in_class = int((g["q"][:, 0] >= 0).sum())
        np.savez_compressed(os.path.join(HERE, "ecfilter", name + ".npz"), kept=kept)
        print(name, "alleles", n, "in a class", in_class, "kept", int(kept.sum()), "removed", in_class - int(kept.sum()))


if __name__ == "__main__":
    main()
