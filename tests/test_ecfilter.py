"""SURVEY.md 8f N4 — Genotyper::RemoveLowLikelihoodAlleleInEquivalentClass (Genotyper.hpp:1371-1460, Genotyper.cpp:647).
Goldens (tests/golden/ecfilter, make_golden_ecfilter.py): the alleles the UNMODIFIED reference keeps in their classes on the six
golden workloads and on "twins" (alleles with an uncovered extension, where the filter does remove members).  CPU: the
host-side ABI entry over the oracle's read groups; `-m gpu`: allele_kept of t1k_genotype."""
import os

import numpy as np
import pytest

import golden_io as G
import oracle_py as O
from t1k_b200 import _lib as L
from t1k_b200 import dist_em
from t1k_b200.refset import RefSet

CASES = G.names() + ["twins"]


def load_case(name):
    z = np.load(os.path.join(G.GOLDEN, "ecfilter", name + ".npz"))
    if name == "twins":
        recs = G.parse_fasta_bytes(z["fasta"].tobytes())
        return dict(records=recs, reads1=z["reads1"], reads2=z["reads2"], similarity=float(z["similarity"]), relax=bool(int(z["relax"])),
                    kept=z["kept"], equivalent_class=z["equivalent_class"], ec_abundance=z["ec_abundance"])
    g = G.load(name)
    return dict(records=g["records"], reads1=g["reads1"], reads2=g["reads2"], similarity=g["similarity"], relax=g["relax"], kept=z["kept"],
                equivalent_class=g["q"][:, 0].astype(np.int32), ec_abundance=None)


@pytest.mark.parametrize("name", CASES)
def test_ec_filter_over_the_oracles_groups(name):
    """t1k_groups_ec_filter (host ABI) on the read groups / classes / ecAbundance the oracle flow produces == the reference"""
    c = load_case(name)
    ref = RefSet(c["records"])
    kept_recs, w = O.collapse_reference(c["records"])
    sw = O.seq_weights(kept_recs, w)
    R = O.genotype_pipeline(O.Oracle(kept_recs, c["similarity"], c["relax"], sw), c["reads1"], c["reads2"], ref.names, sw)
    assert np.array_equal(R["allele_ec"], c["equivalent_class"])
    g = dist_em.ReadGroups()
    ptr = np.zeros(len(R["frags"]) + 1, dtype=np.uint64)
    ptr[1:] = np.cumsum([len(f) for f in R["frags"]])
    ent = (np.concatenate(R["frags"]) if ptr[-1] else np.zeros(0, dtype=L.ASSIGN_DT)).astype(L.ASSIGN_DT)
    g.add_fragments(ptr, ent)
    ecp = np.zeros(len(R["ecs"]) + 1, dtype=np.int32)
    ecp[1:] = np.cumsum([len(e) for e in R["ecs"]])
    eca = np.asarray([a for e in R["ecs"] for a in e], dtype=np.int32)
    lens = np.asarray([len(s) for s in ref.seqs], dtype=np.int32)
    kept, span = g.ec_filter(lens, R["ec_abundance"], ecp, eca)
    assert np.array_equal(kept, c["kept"])
    if name == "twins":
        assert 0 < int((c["equivalent_class"] >= 0).sum()) - int(kept.sum())       # the filter does remove members here
    # the covered ranges are the per-allele min start / max end over the groups
    _, gent, _ = g.fetch()
    for a in np.unique(gent["alleleIdx"])[:50]:
        m = gent["alleleIdx"] == a
        assert span[a] == gent["start"][m].min() and span[ref.n + a] == gent["end"][m].max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_flow_reports_the_kept_alleles(name):
    from t1k_b200.genotyper import Genotyper
    c = load_case(name)
    ref = RefSet(c["records"])
    out = Genotyper(ref, c["similarity"], c["relax"]).Genotype(c["reads1"], c["reads2"])
    assert np.array_equal(out["equivalent_class"], c["equivalent_class"])
    assert np.array_equal(out["allele_kept"], c["kept"])
    if c["ec_abundance"] is not None:
        assert np.array_equal(out["ec_abundance"], c["ec_abundance"])
