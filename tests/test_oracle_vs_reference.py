"""Differential test of the CPU oracle against the unmodified reference compiled in-container
(oracle/_ref/ref_harness <- /root/reference, oracle/Makefile).  Skipped where the reference checkout or the
prebuilt harness is absent (the GPU box uses the committed golden vectors instead)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O
import workloads as W
from t1k_b200 import synth

pytestmark = pytest.mark.skipif(not os.path.exists(O.REF_HARNESS), reason="reference harness not built")


def test_global_alignment_random(tmp_path):
    rng = np.random.default_rng(3)
    alpha = np.frombuffer(b"ACGTN", dtype=np.uint8)
    pairs = []
    for it in range(3000):
        n = int(rng.integers(1, 160))
        t = alpha[rng.integers(0, 4, size=n)]
        if it % 3 == 0:
            t = np.resize(alpha[rng.integers(0, 4, size=int(rng.integers(1, 4)))], n)
        p = t.copy()
        for _ in range(int(rng.integers(0, 8))):
            p[rng.integers(0, len(p))] = alpha[rng.integers(0, 5)]
        if it % 4 == 0 and len(p) > 8:
            k = int(rng.integers(1, len(p) - 1))
            d = int(rng.integers(1, 6))
            p = np.delete(p, slice(k, k + d)) if rng.integers(0, 2) else np.insert(p, k, alpha[rng.integers(0, 4, size=d)])
        if len(p) == 0:
            continue
        pairs.append((t.tobytes(), p.tobytes()))
    inp = b"".join(a + b" " + b + b"\n" for a, b in pairs)
    out = subprocess.run([O.REF_HARNESS, "align"], input=inp, stdout=subprocess.PIPE, check=True).stdout.decode().split("\n")
    for (t, p), line in zip(pairs, out):
        score, ops = line.split()
        s, o = O.global_alignment(t, p)
        assert s == int(score) and "".join(map(str, o.tolist())) == (ops if ops != "-" else ""), (t, p)


@pytest.mark.parametrize("kind,sim,relax,se", [("rna", 0.8, False, False), ("dna", 0.9, True, False), ("rna", 0.97, False, True)])
def test_pipeline_random(tmp_path, kind, sim, relax, se):
    recs = W.small_rna_ref(seed=51) if kind == "rna" else W.small_dna_ref(seed=52)
    kept, w = O.collapse_reference(recs)
    fa = str(tmp_path / "ref.fa")
    synth.write_fasta(fa, recs)
    r1, r2 = W.reads_for(kept, 250, seed=53, single_end=se, err=0.015, n_rate=0.003, indel_rate=0.1)
    W.write_lines(str(tmp_path / "r1.txt"), r1)
    cmd = [O.REF_HARNESS, "genotype", "-f", fa, "-1", str(tmp_path / "r1.txt"), "-o", str(tmp_path / "out"), "-s", str(sim), "--cov"]
    if not se:
        W.write_lines(str(tmp_path / "r2.txt"), r2)
        cmd += ["-2", str(tmp_path / "r2.txt")]
    if relax:
        cmd.append("--relaxIntronAlign")
    subprocess.check_call(cmd)
    H = O.parse_harness(str(tmp_path / "out"))
    sw = O.seq_weights(kept, w)
    orc = O.Oracle(kept, sim, relax, sw)
    R = O.genotype_pipeline(orc, r1, r2, [k[0] for k in kept], sw)
    for u in H["uniq"]:
        got = [tuple(int(x) for x in o) for o in R["uniq"][u["seq"]]]
        assert got == [o[:10] for o in u["ov"]]
    for i, f in enumerate(H["frag"]):
        mine = [(int(a["alleleIdx"]), int(a["start"]), int(a["end"]), float(a["weight"]), float(a["qual"]), float(a["adjustWeight"]))
                for a in R["frags"][i]]
        assert mine == f["as"], i
    assert H["ecs"] == R["ecs"] and H["iters"] == R["iters"] and H["missing"] == R["missing"].tolist()
    assert np.array_equal(np.array([q[1] for q in H["q"]]), R["abundance"])
    cov = np.concatenate([orc.coverage(a) for a in range(len(kept))])
    assert np.array_equal(cov, np.concatenate([H["cov"][a] for a in range(len(kept))]))
