"""HLA-scale golden: the UNMODIFIED reference (oracle/_ref/ref_harness) on the bench configuration itself — the synthetic
30,000-allele HLA-RNA-like reference of bench.py (regenerated from its seed, not stored), a few 150 bp pairs, -s 0.97.
Stores the reference's AssignRead records per unique read-end and the whole-flow per-allele outputs.
Runs only where /root/reference exists.  Usage: python tests/golden/make_golden_hla.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
import oracle_py as O  # noqa: E402
import workloads as W  # noqa: E402
from t1k_b200 import synth  # noqa: E402

N_PAIRS, SEED = 12, 4242


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    recs, ref, r1, r2 = bench.make_workload(N_PAIRS, SEED)
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "ref.fa")
        synth.write_fasta(fa, recs)
        W.write_lines(os.path.join(td, "r1.txt"), r1)
        W.write_lines(os.path.join(td, "r2.txt"), r2)
        subprocess.check_call([O.REF_HARNESS, "genotype", "-f", fa, "-1", os.path.join(td, "r1.txt"), "-2", os.path.join(td, "r2.txt"),
                               "-o", os.path.join(td, "out"), "-s", "0.97"])
        H = O.parse_harness(os.path.join(td, "out"))
    ptr = np.zeros(len(H["uniq"]) + 1, dtype=np.int64)
    np.cumsum([len(u["ov"]) for u in H["uniq"]], out=ptr[1:])
    ov = np.asarray([o[:10] for u in H["uniq"] for o in u["ov"]], dtype=np.int32).reshape(-1, 10)
    # records are stored as deltas against the previous record of the list: they compress to a fraction
    q = np.asarray(H["q"], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "hla_scale", "reference_outputs.npz"), n_pairs=N_PAIRS, seed=SEED, reads1=r1, reads2=r2,
                        uniq_seq=np.asarray([u["seq"] for u in H["uniq"]]), uniq_weight=np.asarray([u["weight"] for u in H["uniq"]], dtype=np.int32),
                        uniq_ptr=ptr, uniq_ov=ov, iters=H.get("iters", 0), q=q, aligned=H["aligned"], n_groups=len(H["groups"]), n_ec=len(H["ecs"]))
    print("alleles", H["nAlleles"], "uniq", len(H["uniq"]), "records", len(ov), "groups", len(H["groups"]), "ecs", len(H["ecs"]), "iters", H.get("iters"))


if __name__ == "__main__":
    main()
