"""The CPU oracle against the committed outputs of the unmodified reference (tests/golden/*.npz): integer fields and
float32 weights bit-exact, EM iteration count equal, abundances bit-exact (same summation order on the CPU)."""
import numpy as np
import pytest

import golden_io as G
import oracle_py as O


@pytest.fixture(scope="module", params=G.names())
def case(request):
    g = G.load(request.param)
    kept, w = O.collapse_reference(g["records"])
    sw = O.seq_weights(kept, w)
    orc = O.Oracle(kept, g["similarity"], g["relax"], sw)
    return g, kept, sw, orc


def test_assign_read_matches_reference(case):
    g, kept, sw, orc = case
    orc.coverage_reset()
    for i, s in enumerate(g["uniq_seq"]):
        ret, ov = orc.assign(s, int(g["uniq_weight"][i]))
        want = G.uniq_overlaps(g, i)
        got = np.asarray([tuple(int(x) for x in o) for o in ov], dtype=np.int32).reshape(-1, 10)
        assert np.array_equal(got, want), "read-end %d" % i
    cov = np.concatenate([orc.coverage(a) for a in range(len(kept))])
    assert np.array_equal(cov, g["cov"])
    assert [orc.missing_coverage(a) for a in range(len(kept))] == g["missing"].tolist()


def test_pipeline_matches_reference(case):
    g, kept, sw, orc = case
    R = O.genotype_pipeline(orc, g["reads1"], g["reads2"], [k[0] for k in kept], sw)
    for i in range(len(g["reads1"])):
        want = G.frag_rows(g, i)
        a = R["frags"][i]
        got = np.stack([a["alleleIdx"], a["start"], a["end"], a["weight"], a["qual"], a["adjustWeight"]], axis=1).astype(np.float64) \
            if len(a) else np.zeros((0, 6))
        assert np.array_equal(got, want), "fragment %d" % i
    assert R["assigned"] == g["aligned"]
    assert len(R["groups"]) == len(g["group_ptr"]) - 1
    for k, grp in enumerate(R["groups"]):
        want = g["group"][g["group_ptr"][k]:g["group_ptr"][k + 1]]
        assert np.array_equal(grp["alleleIdx"], want[:, 0].astype(np.int32))
        assert np.array_equal(grp["weight"].astype(np.float64), want[:, 1])
    assert [len(e) for e in R["ecs"]] == np.diff(g["ec_ptr"]).tolist()
    assert [a for e in R["ecs"] for a in e] == g["ec"].tolist()
    assert R["iters"] == g["iters"]
    q = g["q"]
    assert np.array_equal(R["allele_ec"], q[:, 0].astype(np.int32))
    assert np.array_equal(R["abundance"], q[:, 1])
    assert np.array_equal(R["ec_abundance"], q[:, 2])
    assert np.array_equal(R["eff_len"], q[:, 3].astype(np.int32))
    assert np.array_equal(np.asarray(sw), q[:, 4].astype(np.int32))


def test_global_alignment_known_answers():
    # hand-checked properties of AlignAlgo::GlobalAlignment (AlignAlgo.hpp:215-421)
    s, ops = O.global_alignment(b"ACGTACGT", b"ACGTACGT")
    assert s == 16 and ops.tolist() == [0] * 8
    s, ops = O.global_alignment(b"ACGTACGT", b"ACGAACGT")
    assert s == 12 and ops.tolist() == [0, 0, 0, 1, 0, 0, 0, 0]
    s, ops = O.global_alignment(b"ACGTNCGT", b"ACGTACGT")       # N matches anything (Q6)
    assert s == 16 and ops.tolist() == [0] * 8
    s, ops = O.global_alignment(b"ACGTTACGTACGT", b"ACGTACGTACGT")   # one deletion: 12 matches - gap open 4 - extend 1
    assert s == 24 - 5 and sorted(ops.tolist()) == [0] * 12 + [3]
    assert O.global_alignment(b"", b"ACGT")[0] == 0
    assert O.global_alignment(b"A", b"C") == (-2, pytest.approx(np.asarray([1], dtype=np.int8)))
