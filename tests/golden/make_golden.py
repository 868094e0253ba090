"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/ref_harness, compiled from
/root/reference by oracle/Makefile).  Runs only where /root/reference exists; the .npz files it writes are what
travels to the GPU box.  Usage: python tests/golden/make_golden.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_py as O  # noqa: E402
import workloads as W  # noqa: E402
from t1k_b200 import synth  # noqa: E402

CASES = {
    # name: (records factory, similarity, relax, single_end, n_fragments, read_len, seed)
    "rna_pe": (lambda: W.small_rna_ref(), 0.8, False, False, 90, 100, 21),
    "rna_hla_preset": (lambda: W.small_rna_ref(seed=6), 0.97, False, False, 90, 100, 22),
    "dna_relax_pe": (lambda: W.small_dna_ref(), 0.9, True, False, 120, 100, 23),
    "dna_se": (lambda: W.small_dna_ref(seed=8), 0.9, True, True, 120, 90, 24),
    "cyp2d6_rna": (lambda: W.cyp2d6("rna"), 0.8, False, False, 40, 100, 25),
    "cyp2d6_dna": (lambda: W.cyp2d6("dna"), 0.9, True, False, 40, 125, 26),
}


def flat(rows, width, dtype):
    ptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in rows], out=ptr[1:])
    data = np.asarray([x for r in rows for x in r], dtype=dtype).reshape(-1, width) if ptr[-1] else np.zeros((0, width), dtype=dtype)
    return ptr, data


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    for name, (factory, sim, relax, se, n, rl, seed) in CASES.items():
        recs = factory()
        kept, _ = O.collapse_reference(recs)
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "ref.fa")
            synth.write_fasta(fa, recs)
            r1, r2 = W.reads_for(kept, n, read_len=rl, seed=seed, single_end=se, insert=(rl + 60, rl + 200))
            # duplicates + a read without any hit + a read shorter than k
            r1[n // 2] = r1[0]
            if r2 is not None:
                r2[n // 2] = r2[0]
            rng = np.random.default_rng(seed)
            r1[n - 1] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=rl)]
            W.write_lines(os.path.join(td, "r1.txt"), r1)
            cmd = [O.REF_HARNESS, "genotype", "-f", fa, "-1", os.path.join(td, "r1.txt"), "-o", os.path.join(td, "out"),
                   "-s", str(sim), "--cov"]
            if r2 is not None:
                W.write_lines(os.path.join(td, "r2.txt"), r2)
                cmd += ["-2", os.path.join(td, "r2.txt")]
            if relax:
                cmd.append("--relaxIntronAlign")
            subprocess.check_call(cmd)
            H = O.parse_harness(os.path.join(td, "out"))
        uptr, uov = flat([[o[:10] for o in u["ov"]] for u in H["uniq"]], 10, np.int32)
        fptr, fas = flat([f["as"] for f in H["frag"]], 6, np.float64)
        gptr, grp = flat(H["groups"], 2, np.float64)
        eptr, ecs = flat([[(a,) for a in e] for e in H["ecs"]], 1, np.int32)
        cov = np.concatenate([H["cov"][a] for a in range(H["nAlleles"])])
        q = np.asarray(H["q"], dtype=np.float64)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            fasta=np.frombuffer(b"".join(b">" + n_.encode() + (b" " + c.encode() if c else b"") + b"\n" + s + b"\n" for n_, c, s in recs), dtype=np.uint8),
            similarity=sim, relax=int(relax), single_end=int(se), reads1=r1, reads2=r2 if r2 is not None else np.zeros((0, 0), np.uint8),
            uniq_seq=np.asarray([u["seq"] for u in H["uniq"]]), uniq_weight=np.asarray([u["weight"] for u in H["uniq"]], dtype=np.int32),
            uniq_ptr=uptr, uniq_ov=uov, frag_ptr=fptr, frag_as=fas, group_ptr=gptr, group=grp, ec_ptr=eptr, ec=ecs.reshape(-1),
            missing=np.asarray(H["missing"], dtype=np.int32), iters=H.get("iters", 0), q=q, cov=cov, aligned=H["aligned"])
        print(name, "alleles", H["nAlleles"], "uniq", len(H["uniq"]), "overlaps", len(uov), "assignments", len(fas), "groups", len(H["groups"]),
              "ecs", len(H["ecs"]), "iters", H.get("iters"))


if __name__ == "__main__":
    main()
