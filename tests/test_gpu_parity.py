"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the committed reference outputs.
Integer fields and float32 weights bit-exact; EM iteration count equal; abundances within 1e-5 relative
(BASELINE.json north_star; the device sums in a different order than the reference's sequential loop)."""
import os

import numpy as np
import pytest

import golden_io as G
import oracle_py as O
import workloads as W
from t1k_b200 import synth
from t1k_b200._lib import ASSIGN_DT, OVERLAP_DT, T1KError
from t1k_b200.genotyper import Genotyper, QuantifyAlleleEquivalentClass, SeqSet
from t1k_b200.refset import RefSet

pytestmark = pytest.mark.gpu

ABUND_RTOL = 1e-5


def assert_abundance_close(got, want):
    """north_star: abundances within 1e-5 relative.  Components many orders below the top allele are EM dust whose
    trajectory amplifies any rounding difference (the reference itself moves them when the read order changes), so the
    bound is taken relative to the largest abundance of the sample as well."""
    scale = float(np.abs(want).max()) if len(want) else 0.0
    np.testing.assert_allclose(got, want, rtol=ABUND_RTOL, atol=ABUND_RTOL * scale + 1e-12)


def as_rows(ov):
    return np.asarray([tuple(int(x) for x in o) for o in ov], dtype=np.int32).reshape(-1, 10)


def rec_rows(rec):
    return np.stack([rec[n] for n in OVERLAP_DT.names], axis=1).astype(np.int32) if len(rec) else np.zeros((0, 10), np.int32)


def ent_rows(ent):
    return np.stack([ent[n].astype(np.float64) for n in ASSIGN_DT.names], axis=1) if len(ent) else np.zeros((0, 6))


def uniq_batch(reads1, reads2):
    """unique read-ends, weights and the per-fragment indices into them"""
    seqs = [r.tobytes() for r in reads1] + ([r.tobytes() for r in reads2] if reads2 is not None else [])
    index, uniq, w = {}, [], []
    ids = []
    for s in seqs:
        k = index.get(s)
        if k is None:
            k = index[s] = len(uniq)
            uniq.append(s)
            w.append(0)
        w[k] += 1
        ids.append(k)
    n = len(reads1)
    e1 = np.asarray(ids[:n], dtype=np.uint32)
    e2 = np.asarray(ids[n:], dtype=np.uint32) if reads2 is not None else None
    return uniq, np.asarray(w, dtype=np.int32), e1, e2


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module", params=G.names())
def golden(request):
    g = G.load(request.param)
    ref = RefSet(g["records"])
    ss = SeqSet(ref, g["similarity"], g["relax"])
    return g, ref, ss


def test_refset_matches_reference(golden):
    g, ref, ss = golden
    q = g["q"]
    assert ss.Size() == len(q)
    assert np.array_equal(ref.effective_len, q[:, 3].astype(np.int32))
    assert np.array_equal(ref.seq_weight, q[:, 4].astype(np.int32))


def test_assign_read_golden(golden):
    """SeqSet::AssignRead on the GPU == the unmodified reference, record for record, plus base coverage."""
    g, ref, ss = golden
    ss.ResetBaseCoverage()
    a = ss.AssignRead(g["uniq_seq"], g["uniq_weight"])
    row, ret, rec = a.fetch()
    rows = rec_rows(rec)
    for i in range(len(g["uniq_seq"])):
        want = G.uniq_overlaps(g, i)
        got = rows[int(row[i]):int(row[i + 1])]
        assert np.array_equal(got, want), "read-end %d" % i
    assert np.array_equal(ss.GetBaseCoverage(), g["cov"])
    assert np.array_equal(ss.GetSeqMissingBaseCoverage(), g["missing"])


def test_fragment_assignment_golden(golden):
    g, ref, ss = golden
    uniq, w, e1, e2 = uniq_batch(g["reads1"], g["reads2"])
    a = ss.AssignRead(uniq, w)
    has_n = np.asarray([(b"N" in g["reads1"][i].tobytes()) or (g["reads2"] is not None and b"N" in g["reads2"][i].tobytes())
                        for i in range(len(e1))], dtype=np.uint8)
    row, ent = ss.ReadAssignmentToFragmentAssignment(a, e1, e2, has_n, 2000)
    rows = ent_rows(ent)
    for i in range(len(e1)):
        assert np.array_equal(rows[int(row[i]):int(row[i + 1])], G.frag_rows(g, i)), "fragment %d" % i


def test_genotype_golden(golden):
    """The whole Genotyper.cpp:450-646 flow in one C-ABI call against the reference's per-allele outputs."""
    g, ref, ss = golden
    gt = Genotyper(ref, g["similarity"], g["relax"])
    out = gt.Genotype(g["reads1"], g["reads2"])
    q = g["q"]
    assert out["assigned_fragments"] == g["aligned"]
    assert out["n_groups"] == len(g["group_ptr"]) - 1
    assert out["n_ec"] == len(g["ec_ptr"]) - 1
    assert np.array_equal(out["equivalent_class"], q[:, 0].astype(np.int32))
    assert np.array_equal(out["missing_coverage"], g["missing"])
    assert out["em_iterations"] == g["iters"]
    # default EM: every sum in the reference's order -> the reference's doubles, bit for bit
    assert np.array_equal(out["abundance"], q[:, 1])
    assert np.array_equal(out["ec_abundance"], q[:, 2])
    # `fragmentAssigned` (Genotyper.cpp:564) is set from the pairing result BEFORE SetReadAssignments' cuts, so it
    # covers every fragment that kept rows (tests/test_dropin.py checks the flag itself through _aligned*.fa)
    frag_has = np.diff(g["frag_ptr"]) > 0
    assert (out["fragment_assigned"].astype(bool) | ~frag_has).all()
    # tree-reduction EM: within the north-star tolerance
    fast = Genotyper(ref, g["similarity"], g["relax"], em_fast_sums=True).Genotype(g["reads1"], g["reads2"])
    assert np.array_equal(fast["equivalent_class"], out["equivalent_class"])
    assert fast["em_iterations"] > 0 and np.isfinite(fast["abundance"]).all()
    assert_abundance_close(fast["abundance"], out["abundance"])          # 1e-5 of the top allele (tolerance of the north star)
    assert_abundance_close(fast["ec_abundance"], out["ec_abundance"])


# ------------------------------------------------------------------------------------------------
# seeded workloads against the oracle (sizes the oracle finishes in seconds)
WORKLOADS = {
    "rna_s80": (lambda: W.small_rna_ref(seed=31), 0.8, False, dict(read_len=100, err=0.02, n_rate=0.004, indel_rate=0.15)),
    "rna_s97_150": (lambda: W.small_rna_ref(seed=32), 0.97, False, dict(read_len=150, err=0.004, n_rate=0.0, indel_rate=0.02, insert=(200, 420))),
    "dna_relax": (lambda: W.small_dna_ref(seed=33), 0.9, True, dict(read_len=125, err=0.01, n_rate=0.002, indel_rate=0.08, insert=(200, 420))),
    "dna_se_short": (lambda: W.small_dna_ref(seed=34), 0.9, True, dict(read_len=60, err=0.01, n_rate=0.0, indel_rate=0.05, single_end=True)),
}


@pytest.fixture(scope="module", params=sorted(WORKLOADS))
def workload(request):
    factory, sim, relax, kw = WORKLOADS[request.param]
    recs = factory()
    ref = RefSet(recs)
    kept, w = O.collapse_reference(recs)
    orc = O.Oracle(kept, sim, relax, O.seq_weights(kept, w))
    r1, r2 = W.reads_for(kept, 400, seed=77, **kw)
    return dict(ref=ref, kept=kept, orc=orc, sim=sim, relax=relax, r1=r1, r2=r2, sw=O.seq_weights(kept, w))


def test_assign_read_vs_oracle(workload):
    wl = workload
    ss = SeqSet(wl["ref"], wl["sim"], wl["relax"])
    uniq, w, e1, e2 = uniq_batch(wl["r1"], wl["r2"])
    a = ss.AssignRead(uniq, w)
    row, ret, rec = a.fetch()
    rows = rec_rows(rec)
    wl["orc"].coverage_reset()
    for i, s in enumerate(uniq):
        oret, ov = wl["orc"].assign(s, int(w[i]))
        assert ret[i] == oret, "read-end %d" % i
        assert np.array_equal(rows[int(row[i]):int(row[i + 1])], as_rows(ov)), "read-end %d" % i
    cov = np.concatenate([wl["orc"].coverage(k) for k in range(wl["ref"].n)])
    assert np.array_equal(ss.GetBaseCoverage(), cov)
    assert np.array_equal(ss.GetSeqMissingBaseCoverage(), [wl["orc"].missing_coverage(k) for k in range(wl["ref"].n)])
    # analyzer mode (weight 0, Analyzer.cpp:476): same records, no coverage
    ss.ResetBaseCoverage()
    a0 = ss.AssignRead(uniq, np.zeros(len(uniq), dtype=np.int32))
    row0, ret0, rec0 = a0.fetch()
    assert np.array_equal(row0, row) and np.array_equal(rec_rows(rec0), rows)
    assert not ss.GetBaseCoverage().any()


def test_genotype_vs_oracle(workload):
    wl = workload
    R = O.genotype_pipeline(wl["orc"], wl["r1"], wl["r2"], wl["ref"].names, wl["sw"])
    gt = Genotyper(wl["ref"], wl["sim"], wl["relax"])
    out = gt.Genotype(wl["r1"], wl["r2"])
    assert out["assigned_fragments"] == R["assigned"]
    assert out["n_groups"] == len(R["groups"]) and out["n_ec"] == len(R["ecs"])
    assert np.array_equal(out["equivalent_class"], R["allele_ec"])
    assert np.array_equal(out["missing_coverage"], R["missing"])
    assert out["em_iterations"] == R["iters"]
    assert np.array_equal(out["abundance"], R["abundance"])
    assert np.array_equal(out["ec_abundance"], R["ec_abundance"])
    assert out["n_launches"] > 0


def test_chunking_and_store_growth_do_not_change_results(workload, monkeypatch):
    """Small fragment chunks (read-ends re-aligned per chunk with split weights; the three-stage host pipeline runs
    with many chunks), a record store that overflows and is grown (deferred read-ends re-run), a pairing row buffer
    that overflows and the multi-threaded host tail must give the very same integers."""
    wl = workload
    gt = Genotyper(wl["ref"], wl["sim"], wl["relax"])
    base = gt.Genotype(wl["r1"], wl["r2"])
    monkeypatch.setenv("T1K_CHUNK_FRAGMENTS", "37")
    monkeypatch.setenv("T1K_STORE_RECORDS", "2000")
    monkeypatch.setenv("T1K_PAIR_ROWS", "64")          # pairing rows overflow their first buffer: exact re-run
    monkeypatch.setenv("T1K_PAR_MIN", "1")             # the threaded host tail (EC build, EM inputs, CSC) on small inputs too
    out = gt.Genotype(wl["r1"], wl["r2"])
    for k in ("equivalent_class", "missing_coverage", "fragment_assigned"):
        assert np.array_equal(out[k], base[k]), k
    assert out["em_iterations"] == base["em_iterations"]
    assert out["n_assignments"] == base["n_assignments"]
    assert np.array_equal(out["abundance"], base["abundance"])


def test_fast_path_and_hit_list_path_agree_on_device(workload, monkeypatch):
    """T1K_NO_FAST=1 sends every allele group through the hit-list path (chain_allele, extend_cand, full_align) instead of
    the mismatch-mask fast path: records, coverage and the whole-flow outputs must be identical."""
    wl = workload
    uniq, w, e1, e2 = uniq_batch(wl["r1"], wl["r2"])
    res = []
    for no_fast in ("0", "1"):
        monkeypatch.setenv("T1K_NO_FAST", no_fast)
        ss = SeqSet(wl["ref"], wl["sim"], wl["relax"])
        row, ret, rec = ss.AssignRead(uniq, w).fetch()
        out = Genotyper(wl["ref"], wl["sim"], wl["relax"]).Genotype(wl["r1"], wl["r2"])
        res.append((row, ret, rec_rows(rec), ss.GetBaseCoverage(), out))
    a, b = res
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[3], b[3])
    for k in ("equivalent_class", "missing_coverage", "fragment_assigned", "abundance"):
        assert np.array_equal(a[4][k], b[4][k]), k
    assert a[4]["em_iterations"] == b[4]["em_iterations"] and a[4]["n_assignments"] == b[4]["n_assignments"]


def test_bench_configuration_against_the_reference_golden():
    """The device path on the bench configuration (30,000 alleles, > 1000 cut active) against the UNMODIFIED reference's
    records (tests/golden/hla_scale) and whole-flow summary."""
    import bench
    g = G.load_hla_scale()
    recs, ref, r1, r2 = bench.make_workload(g["n_pairs"], g["seed"])
    ss = SeqSet(ref, 0.97, False)
    row, ret, rec = ss.AssignRead(g["uniq_seq"], g["uniq_weight"]).fetch()
    rows = rec_rows(rec)
    for i in range(len(g["uniq_seq"])):
        want = g["uniq_ov"][g["uniq_ptr"][i]:g["uniq_ptr"][i + 1]]
        assert np.array_equal(rows[int(row[i]):int(row[i + 1])], want), "read-end %d" % i
    out = Genotyper(ref, 0.97, False).Genotype(r1, r2)
    assert out["assigned_fragments"] == g["aligned"]
    assert out["n_groups"] == g["n_groups"] and out["n_ec"] == g["n_ec"]
    assert out["em_iterations"] == g["iters"]
    q = g["q"]
    assert np.array_equal(out["equivalent_class"], q[:, 0].astype(np.int32))
    assert np.array_equal(out["abundance"], q[:, 1])


def test_em_vs_oracle(workload):
    wl = workload
    R = O.genotype_pipeline(wl["orc"], wl["r1"], wl["r2"], wl["ref"].names, wl["sw"])
    P = R["problem"]
    for mask in (False, True):
        kw = dict(ec_allele_ptr=P["ec_allele_ptr"], ec_alleles=P["ec_alleles"], allele_major=R["major"], allele_gene=R["gene"]) if mask else {}
        it, x, rc = O.em(P["rowptr"], P["col"], P["count"], P["eclen"], P["x0"], 0.0, 0.15, **kw)
        git, gx, grc, info = QuantifyAlleleEquivalentClass(P["rowptr"], P["col"], P["count"], P["eclen"], P["x0"], 0.0, 0.15, **kw)
        assert git == it
        assert np.array_equal(gx, x) and np.array_equal(grc, rc)          # reference-order sums: bit-identical
        assert info["n_launches"] > 0
        git, gx, grc, info = QuantifyAlleleEquivalentClass(P["rowptr"], P["col"], P["count"], P["eclen"], P["x0"], 0.0, 0.15,
                                                           fast_sums=True, **kw)
        # tree reductions: the SQUAREM trajectory (and so the iteration count, and the split between alleles the reads
        # cannot tell apart) is free to differ; the expected read counts still add up to the reads
        assert git > 0
        np.testing.assert_allclose(grc.sum(), rc.sum(), rtol=1e-9)
        # ... and both end on the same likelihood: where the reads cannot tell two equivalence classes apart the likelihood is
        # flat and the point an EM run stops at depends on every rounding (the abundances themselves are compared, to 1e-5 of
        # the top allele, on the golden samples and on the bench configuration, where they are identifiable)
        def loglik(xv):
            rp, cl = np.asarray(P["rowptr"]), np.asarray(P["col"])
            tot = np.add.reduceat(np.asarray(xv)[cl], rp[:-1])[np.diff(rp) > 0]
            return float((np.asarray(P["count"])[np.diff(rp) > 0] * np.log(np.maximum(tot, 1e-300))).sum())
        np.testing.assert_allclose(loglik(gx), loglik(x), rtol=2e-6)      # (the stop test |dx| < 1e-5 leaves the likelihood ~1e-4 short of its maximum)
    # --squaremMinAlpha (Genotyper.hpp:1243-1244)
    it, x, rc = O.em(P["rowptr"], P["col"], P["count"], P["eclen"], P["x0"], -2.0, 0.15)
    git, gx, grc, _ = QuantifyAlleleEquivalentClass(P["rowptr"], P["col"], P["count"], P["eclen"], P["x0"], -2.0, 0.15)
    assert git == it and np.array_equal(gx, x)


# ------------------------------------------------------------------------------------------------
# edge cases
def test_edge_reads():
    recs = W.small_dna_ref(seed=41)
    ref = RefSet(recs)
    kept, w = O.collapse_reference(recs)
    orc = O.Oracle(kept, 0.9, True, w)
    ss = SeqSet(ref, 0.9, True)
    a0 = kept[0][2]
    reads = [
        b"ACGTACGTAC",                       # shorter than k: AssignRead returns -1 (SeqSet.hpp:1598)
        b"N" * 80,                           # no valid k-mer
        b"ACGT" * 20,                        # low complexity / no hit
        a0[100:355],                         # maximum supported length (255), exact copy
        a0[10:95],                           # ragged lengths
        a0[300:420][::-1],                   # reversed (not complemented): no alignment
        bytes(synth.revcomp(np.frombuffer(a0[200:330], dtype=np.uint8))),   # reverse strand
        a0[:70],                             # starts at the allele boundary (left clip territory)
        a0[-70:],                            # ends at the allele boundary
        a0[150:200] + b"N" + a0[201:260],    # N inside the read (Q6)
        a0[400:440] + b"ACGTTGCA" + a0[440:500],   # insertion
        a0[500:560] + a0[566:640],           # deletion
    ]
    # a read spanning an N separator of the reference
    npos = a0.find(b"N")
    if npos > 60:
        reads.append(a0[npos - 50:npos + 50])
    a = ss.AssignRead(reads, np.arange(1, len(reads) + 1, dtype=np.int32))
    row, ret, rec = a.fetch()
    rows = rec_rows(rec)
    for i, s in enumerate(reads):
        oret, ov = orc.assign(s, i + 1)
        assert ret[i] == oret, i
        assert np.array_equal(rows[int(row[i]):int(row[i + 1])], as_rows(ov)), i
    cov = np.concatenate([orc.coverage(k) for k in range(ref.n)])
    assert np.array_equal(ss.GetBaseCoverage(), cov)


def test_empty_and_invalid_inputs():
    ref = RefSet(W.small_rna_ref(seed=42))
    ss = SeqSet(ref, 0.8, False)
    a = ss.AssignRead([], [])
    row, ret, rec = a.fetch()
    assert row.tolist() == [0] and len(rec) == 0
    row2, ent = ss.ReadAssignmentToFragmentAssignment(a, [], None)
    assert row2.tolist() == [0] and len(ent) == 0
    with pytest.raises(T1KError) as e:
        ss.AssignRead([b"ACGTACGTACGTXACGTACGTACGT"], [1])
    assert e.value.code == 3
    with pytest.raises(T1KError) as e:
        ss.AssignRead([b"A" * 1001], [1])                  # T1K_MAX_READ_LEN = 1000
    assert e.value.code == 3
    gt = Genotyper(ref, 0.8, False)
    rng = np.random.default_rng(1)
    noise = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(16, 100))]
    out = gt.Genotype(noise, noise[::-1].copy())           # nothing aligns: zero-assignment path
    assert out["assigned_fragments"] == 0 and out["n_ec"] == 0 and out["em_iterations"] == 0
    assert not out["abundance"].any()


def test_coverage_is_linear_in_weight():
    """posWeight adds commute and scale (Genotyper.cpp:149: weight = number of duplicates)."""
    recs = W.small_rna_ref(seed=43)
    ref = RefSet(recs)
    ss = SeqSet(ref, 0.8, False)
    r1, r2 = W.reads_for([(n, c, s) for n, c, s in zip(ref.names, ref.comments, ref.seqs)], 300, seed=5)
    uniq, w, _, _ = uniq_batch(r1, r2)
    ss.AssignRead(uniq, np.ones(len(uniq), dtype=np.int32))
    c1 = ss.GetBaseCoverage().astype(np.int64)
    ss.ResetBaseCoverage()
    ss.AssignRead(uniq, np.full(len(uniq), 3, dtype=np.int32))
    c3 = ss.GetBaseCoverage().astype(np.int64)
    assert c1.any() and np.array_equal(c3, 3 * c1)
    ss.ResetBaseCoverage()
    perm = np.random.default_rng(0).permutation(len(uniq))
    ss.AssignRead([uniq[i] for i in perm[:len(perm) // 2]], np.ones(len(perm) // 2, dtype=np.int32))
    ss.AssignRead([uniq[i] for i in perm[len(perm) // 2:]], np.ones(len(perm) - len(perm) // 2, dtype=np.int32))
    assert np.array_equal(ss.GetBaseCoverage().astype(np.int64), c1)


@pytest.mark.skipif(os.environ.get("T1K_SKIP_LARGE") == "1", reason="large-size property test disabled")
def test_hla_scale_properties():
    """BASELINE config-2 shaped reference (30k alleles) on a bounded read sample: size-independent properties.
    Every simulated error-free fragment must be assigned to its source allele with full-length matchCnt; shuffling
    the fragments changes no integer output and moves abundances by < 1e-5 relative."""
    recs = synth.make_hla_rna_ref(seed=11)
    ref = RefSet(recs)
    gt = Genotyper(ref, 0.97, False)
    kept = [(n, c, s) for n, c, s in zip(ref.names, ref.comments, ref.seqs)]
    r1, r2, src = synth.simulate_pairs(kept, 2000, read_len=150, insert=(300, 450), err=0.0, seed=3)
    out = gt.Genotype(r1, r2)
    assert out["assigned_fragments"] == 2000
    assert out["n_unique_ends"] <= 4000
    # the source alleles' equivalence classes carry the abundance
    assert sum(out["ec_abundance"][s] > 0 for s in src) >= 0.75 * len(src)
    perm = np.random.default_rng(9).permutation(2000)
    out2 = gt.Genotype(r1[perm], r2[perm])
    assert np.array_equal(out2["missing_coverage"], out["missing_coverage"])
    assert out2["n_assignments"] == out["n_assignments"] and out2["n_ec"] == out["n_ec"]
    assert np.array_equal(out2["fragment_assigned"], out["fragment_assigned"][perm])
    assert_abundance_close(out2["abundance"], out["abundance"])
    # the direct pairing output for a few fragments: the source allele is among the tied best
    ss = gt.refSet
    uniq, w, e1, e2 = uniq_batch(r1[:64], r2[:64])
    a = ss.AssignRead(uniq, w)
    row, ent = ss.ReadAssignmentToFragmentAssignment(a, e1, e2, None, 2000)
    assert (np.diff(row) > 0).all()
    assert (ent["weight"] == 1.0).all() and (ent["qual"] == 1.0).all()


def test_async_assign_batches_equal_the_synchronous_calls():
    """t1k_assign_batch_async / t1k_assign_wait (the async variant): two batches in flight give the records, return
    values and accumulated coverage of the same two synchronous calls; a failing batch reports its error at wait()."""
    factory, sim, relax, kw = WORKLOADS[sorted(WORKLOADS)[0]]
    recs = factory()
    ref = RefSet(recs)
    kept, _ = O.collapse_reference(recs)
    r1, r2 = W.reads_for(kept, 300, seed=91, **kw)
    uniq, w, _, _ = uniq_batch(r1, r2)
    half = len(uniq) // 2
    ss = SeqSet(ref, sim, relax)
    a0, a1 = ss.AssignRead(uniq[:half], w[:half]), ss.AssignRead(uniq[half:], w[half:])
    want = [a0.fetch(), a1.fetch()]
    cov = ss.GetBaseCoverage()
    ss2 = SeqSet(ref, sim, relax)
    j0 = ss2.AssignReadAsync(uniq[:half], w[:half])
    j1 = ss2.AssignReadAsync(uniq[half:], w[half:])          # queued behind j0 while the host is free
    got = [j0.wait().fetch(), j1.wait().fetch()]
    for g, x in zip(got, want):
        assert np.array_equal(g[0], x[0]) and np.array_equal(g[1], x[1]) and np.array_equal(g[2], x[2])
    assert np.array_equal(ss2.GetBaseCoverage(), cov)
    bad = ss2.AssignReadAsync([b"ACGTACGTACGTXACGTACGTACGT"], np.ones(1, dtype=np.int32))
    with pytest.raises(T1KError):
        bad.wait()
    ok = ss2.AssignReadAsync(uniq[:4], w[:4]).wait().fetch()        # the queue keeps working after a failed job
    assert np.array_equal(ok[2], want[0][2][:len(ok[2])])


@pytest.mark.parametrize("switch", ["T1K_HOST_DEDUP", "T1K_HOST_TAIL", "T1K_SYNC_ROWS", "T1K_EM_NO_GRAPH"])
def test_device_stages_equal_their_host_counterparts(workload, switch, monkeypatch):
    """The stages that moved to the device this round keep an A/B switch back to the host code: read-end de-duplication
    (ingest kernels vs unique_read_ends), the global tail (bit-matrix equivalence classes + EM matrix vs EquivalenceClasses /
    EmInputs), the copy-stream hand-over of the fragment rows, the CUDA graph of the SQUAREM iteration.  Every output of the whole
    flow must be identical either way (the EM runs in reference order, so abundances too)."""
    wl = workload
    res = []
    for on in (False, True):
        if on:
            monkeypatch.setenv(switch, "1")
        else:
            monkeypatch.delenv(switch, raising=False)
        monkeypatch.setenv("T1K_CHUNK_FRAGMENTS", "150")          # several chunks: the pipeline stages hand over more than once
        res.append(Genotyper(wl["ref"], wl["sim"], wl["relax"]).Genotype(wl["r1"], wl["r2"]))
    a, b = res
    for k in ("equivalent_class", "missing_coverage", "fragment_assigned", "abundance", "ec_abundance", "allele_kept", "allele_span"):
        assert np.array_equal(a[k], b[k]), (switch, k)
    for k in ("em_iterations", "n_groups", "n_ec", "assigned_fragments", "n_unique_ends", "n_overlaps", "n_assignments", "em_nnz"):
        assert a[k] == b[k], (switch, k)
