"""SURVEY 8f N3 goldens: the edit strings SeqSet::AddOverlapAlignmentInfo (SeqSet.hpp:2657-2680) attaches to every record of
AssignRead(read, -1, 0) — from the UNMODIFIED reference (oracle/_ref/ref_harness alninfo), on the unique read-ends of the six
committed golden workloads.  Writes tests/golden/alninfo/<name>.npz (ops: concatenated int8 edit strings, ops_ptr: one entry per
record of the workload's uniq_ov, in its order).  Runs only where /root/reference exists.
Usage: python tests/golden/make_golden_alninfo.py"""
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
from t1k_b200 import synth  # noqa: E402


def run_alninfo(recs, seqs, sim, relax):
    """-> per read (ret, [(record 10-tuple, ops int8 array)])"""
    with tempfile.TemporaryDirectory() as td:
        fa = os.path.join(td, "ref.fa")
        synth.write_fasta(fa, recs)
        with open(os.path.join(td, "r.txt"), "wb") as f:
            for s in seqs:
                f.write(s + b"\n")
        cmd = [O.REF_HARNESS, "alninfo", "-f", fa, "-1", os.path.join(td, "r.txt"), "-o", os.path.join(td, "out"), "-s", str(sim)]
        if relax:
            cmd.append("--relaxIntronAlign")
        subprocess.check_call(cmd)
        out, cur = [], None
        with open(os.path.join(td, "out")) as f:
            for line in f:
                t = line.split()
                if t[0] == "A":
                    continue
                if t[0] == "R":
                    cur = (int(t[2]), [])
                    out.append(cur)
                elif t[0] == "L":
                    ops = np.zeros(0, dtype=np.int8) if t[1] == "-" else np.frombuffer(t[1].encode(), dtype=np.uint8).astype(np.int8) - 48
                    cur[1][-1] = (cur[1][-1][0], ops)
                else:
                    cur[1].append((tuple(int(x) for x in t[:10]), None))
    return out


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    for name in G.names():
        g = G.load(name)
        res = run_alninfo(g["records"], g["uniq_seq"], g["similarity"], g["relax"])
        ops, ptr, k = [], [0], 0
        for i, (ret, recs) in enumerate(res):
            want = G.uniq_overlaps(g, i)
            assert len(recs) == len(want), (name, i)
            for j, (rec, o) in enumerate(recs):
                assert tuple(int(x) for x in want[j]) == rec, (name, i, j)      # weight 0 returns the same records
                ops.append(o)
                k += len(o)
                ptr.append(k)
        allops = np.concatenate(ops) if ops else np.zeros(0, np.int8)
        np.savez_compressed(os.path.join(HERE, "alninfo", name + ".npz"), ops=allops, ops_ptr=np.asarray(ptr, dtype=np.int64))
        nd = sum(1 for o in ops if ((o == 2) | (o == 3)).any())
        print(name, "records", len(ops), "with indels", nd, "ops", k)


if __name__ == "__main__":
    main()
