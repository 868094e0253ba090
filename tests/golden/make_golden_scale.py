"""Goldens at the scale of BASELINE configs 3 and 4: the UNMODIFIED reference (oracle/_ref/ref_harness genotype) on the bench's
own synthetic references — KIR-DNA-like (17 genes x 90 alleles x ~5 kb, N separators, exon header; -s 0.9 --relaxIntronAlign)
and HLA-DNA-like (30,000 alleles x ~3.5 kb; single-end 100 bp, -s 0.97).  The references are regenerated from
bench.make_reference(config), only the reference's outputs are stored.  Runs only where /root/reference exists.
Usage: python tests/golden/make_golden_scale.py"""
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

CASES = {3: (160, 4343), 4: (48, 4444)}       # config -> (fragments, seed)


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    os.makedirs(os.path.join(HERE, "scale"), exist_ok=True)
    for config, (n, seed) in CASES.items():
        cfg = bench.CONFIGS[config]
        recs, ref, r1, r2 = bench.make_workload(n, seed, config=config)
        # a few duplicates and a read with an N
        r1[n // 2] = r1[0]
        if r2 is not None:
            r2[n // 2] = r2[0]
        r1[3, 40] = ord("N")
        with tempfile.TemporaryDirectory() as td:
            fa = os.path.join(td, "ref.fa")
            synth.write_fasta(fa, recs)
            W.write_lines(os.path.join(td, "r1.txt"), r1)
            cmd = [O.REF_HARNESS, "genotype", "-f", fa, "-1", os.path.join(td, "r1.txt"), "-o", os.path.join(td, "out"), "-s", str(cfg["sim"])]
            if r2 is not None:
                W.write_lines(os.path.join(td, "r2.txt"), r2)
                cmd += ["-2", os.path.join(td, "r2.txt")]
            if cfg["relax"]:
                cmd.append("--relaxIntronAlign")
            subprocess.check_call(cmd)
            H = O.parse_harness(os.path.join(td, "out"))
        ptr = np.zeros(len(H["uniq"]) + 1, dtype=np.int64)
        np.cumsum([len(u["ov"]) for u in H["uniq"]], out=ptr[1:])
        ov = np.asarray([o[:10] for u in H["uniq"] for o in u["ov"]], dtype=np.int32).reshape(-1, 10)
        q = np.asarray(H["q"], dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, "scale", "config%d.npz" % config), config=config, n=n, seed=seed, reads1=r1,
                            reads2=r2 if r2 is not None else np.zeros((0, 0), np.uint8),
                            uniq_seq=np.asarray([u["seq"] for u in H["uniq"]]), uniq_weight=np.asarray([u["weight"] for u in H["uniq"]], dtype=np.int32),
                            uniq_ptr=ptr, uniq_ov=ov, iters=H.get("iters", 0), q=q, aligned=H["aligned"], n_groups=len(H["groups"]), n_ec=len(H["ecs"]),
                            missing=np.asarray(H["missing"], dtype=np.int32))
        print("config", config, "alleles", H["nAlleles"], "uniq", len(H["uniq"]), "records", len(ov), "aligned", H["aligned"], "groups", len(H["groups"]),
              "ecs", len(H["ecs"]), "iters", H.get("iters"))


if __name__ == "__main__":
    main()
