"""Goldens for alleles with internal duplications / tandem repeats (tests/dup_workloads.py): AssignRead records and base
coverage of the UNMODIFIED reference (oracle/_ref/ref_harness assign ..., compiled from /root/reference by oracle/Makefile).
Runs only where /root/reference exists.  Usage: python tests/golden/make_golden_dup.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dup_workloads as D  # noqa: E402
import oracle_py as O  # noqa: E402
from t1k_b200 import synth  # noqa: E402

WEIGHT = 2


def run_reference(recs, reads, sim, relax, weight=WEIGHT):
    """-> (ret[n], ptr[n+1], records[*,10], coverage concatenated over alleles)"""
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "ref.fa")
        synth.write_fasta(fa, recs)
        with open(os.path.join(td, "r.txt"), "wb") as f:
            for r in reads:
                f.write(r + b" %d\n" % weight)
        cmd = [O.REF_HARNESS, "assign", "-f", fa, "-1", os.path.join(td, "r.txt"), "-o", os.path.join(td, "out"), "-s", str(sim), "--cov"]
        if relax:
            cmd.append("--relaxIntronAlign")
        subprocess.check_call(cmd)
        H = O.parse_harness(os.path.join(td, "out"))
    ret = np.asarray([u["ret"] for u in H["uniq"]], dtype=np.int32)
    ptr = np.zeros(len(H["uniq"]) + 1, dtype=np.int64)
    np.cumsum([len(u["ov"]) for u in H["uniq"]], out=ptr[1:])
    ov = np.asarray([o[:10] for u in H["uniq"] for o in u["ov"]], dtype=np.int32).reshape(-1, 10)
    cov = np.concatenate([H["cov"][a] for a in range(H["nAlleles"])])
    return ret, ptr, ov, cov


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    os.makedirs(os.path.join(HERE, "dup"), exist_ok=True)
    for name, recs, reads, sim, relax in D.cases():
        ret, ptr, ov, cov = run_reference(recs, reads, sim, relax)
        np.savez_compressed(os.path.join(HERE, "dup", name + ".npz"), ret=ret, ptr=ptr, ov=ov, cov=cov, weight=WEIGHT,
                            n_reads=len(reads), n_alleles=len(recs))
        multi = sum(1 for i in range(len(reads)) if len(set(ov[ptr[i]:ptr[i + 1], 0].tolist())) < ptr[i + 1] - ptr[i])
        print(name, "reads", len(reads), "records", len(ov), "reads with >1 record on one allele", multi)


if __name__ == "__main__":
    main()
