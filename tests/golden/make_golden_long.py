"""Goldens for reads longer than 255 bases (tests/long_workloads.py): AssignRead records and base coverage of the UNMODIFIED
reference (oracle/_ref/ref_harness assign ...).  Runs only where /root/reference exists.
Usage: python tests/golden/make_golden_long.py"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

import long_workloads as LW  # noqa: E402
from make_golden_dup import run_reference  # noqa: E402


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    os.makedirs(os.path.join(HERE, "long"), exist_ok=True)
    for name, recs, reads, sim, relax in LW.cases():
        ret, ptr, ov, cov = run_reference(recs, reads, sim, relax, weight=2)
        np.savez_compressed(os.path.join(HERE, "long", name + ".npz"), ret=ret, ptr=ptr, ov=ov, cov=cov, weight=2, n_reads=len(reads), n_alleles=len(recs))
        print(name, "reads", len(reads), "read length", len(reads[0]), "records", len(ov), "reads without records", int((ret <= 0).sum()))


if __name__ == "__main__":
    main()
