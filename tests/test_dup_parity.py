"""Alleles with internal duplications and tandem repeats (tests/dup_workloads.py): several overlaps of one read on one
allele tie down to the seqStart / seqEnd tail of `_overlap::operator<` (SeqSet.hpp:103-127).  The goldens under
tests/golden/dup/ are the UNMODIFIED reference's AssignRead records and coverage (tests/golden/make_golden_dup.py).
CPU part: the oracle and the lane-code emulation reproduce them; a seeded fuzz compares the emulation with the live
reference harness where it is built.  GPU part: the device path through the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

import dup_workloads as D
import oracle_py as O
from t1k_b200.refset import RefSet

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {name: (recs, reads, sim, relax) for name, recs, reads, sim, relax in D.cases()}


def _golden(name):
    z = np.load(os.path.join(HERE, "golden", "dup", name + ".npz"))
    return {k: z[k] for k in z.files}


def _rows(buf, n):
    return np.stack([buf[k][:max(n, 0)] for k in O.OVERLAP_DT.names], axis=1) if n > 0 else np.zeros((0, 10), np.int32)


def _emu_run(emu, recs, reads, sim, relax, weight, fast=1):
    ref = RefSet(recs)
    bases, off, ptr, se = ref.packed()
    E = emu.emu_create(ref.n, bases, O._p(off), O._p(ptr), O._p(se), sim, int(relax))
    assert E
    emu.emu_set_fast(E, fast)
    buf = np.zeros(1 << 16, dtype=O.OVERLAP_DT)
    out = []
    for s in reads:
        err = C.c_int32(0)
        n = emu.emu_assign(E, s, weight, O._p(buf), len(buf), C.byref(err))
        assert err.value == 0
        out.append((n, _rows(buf, n).copy()))
    cov = []
    for a in range(ref.n):
        c = np.zeros(len(ref.seqs[a]), dtype=np.int32)
        emu.emu_coverage(E, a, O._p(c))
        cov.append(c)
    emu.emu_destroy(E)
    return out, np.concatenate(cov)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_on_duplications(name):
    recs, reads, sim, relax = CASES[name]
    g = _golden(name)
    orc = O.Oracle(recs, sim, relax)
    for i, s in enumerate(reads):
        ret, ov = orc.assign(s, int(g["weight"]))
        assert ret == g["ret"][i], i
        got = np.stack([ov[k] for k in O.OVERLAP_DT.names], axis=1) if len(ov) else np.zeros((0, 10), np.int32)
        assert np.array_equal(got, g["ov"][g["ptr"][i]:g["ptr"][i + 1]]), i
    assert np.array_equal(np.concatenate([orc.coverage(a) for a in range(len(recs))]), g["cov"])


@pytest.mark.parametrize("fast", [1, 0])
@pytest.mark.parametrize("name", sorted(CASES))
def test_lane_code_matches_reference_on_duplications(emu, name, fast):
    recs, reads, sim, relax = CASES[name]
    g = _golden(name)
    out, cov = _emu_run(emu, recs, reads, sim, relax, int(g["weight"]), fast)
    for i, (n, rows) in enumerate(out):
        assert n == g["ret"][i], (i, reads[i])
        assert np.array_equal(rows, g["ov"][g["ptr"][i]:g["ptr"][i + 1]]), (i, reads[i])
    assert np.array_equal(cov, g["cov"])


@pytest.mark.skipif(not os.path.exists(O.REF_HARNESS), reason="reference harness not built")
def test_lane_code_fuzz_against_reference_harness(emu):
    """Seeded fuzz of the emulation against the live reference: duplication lengths, tandem units, thresholds."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_dup import run_reference
    cfgs = []
    for seed in range(4):
        cfgs.append(D.duplication_set(n_alleles=12, dup=40 + 45 * seed, seed=100 + seed, n_reads=120, read_len=70 + 25 * seed) + (0.8 + 0.05 * seed,))
        cfgs.append(D.tandem_set(seed=200 + seed, n_reads=120) + (0.8 + 0.04 * seed,))
    for recs, reads, sim in cfgs:
        ret, ptr, ov, cov = run_reference(recs, reads, sim, False, weight=1)
        out, ecov = _emu_run(emu, recs, reads, sim, False, 1)
        for i, (n, rows) in enumerate(out):
            assert n == ret[i], (i, reads[i])
            assert np.array_equal(rows, ov[ptr[i]:ptr[i + 1]]), (i, reads[i])
        assert np.array_equal(ecov, cov)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_matches_reference_on_duplications(name):
    from t1k_b200.genotyper import SeqSet
    recs, reads, sim, relax = CASES[name]
    g = _golden(name)
    ref = RefSet(recs)
    for no_fast in ("0", "1"):
        os.environ["T1K_NO_FAST"] = no_fast
        try:
            ss = SeqSet(ref, sim, relax)
            a = ss.AssignRead(reads, [int(g["weight"])] * len(reads))
            row_ptr, ret, rec = a.fetch()
            assert np.array_equal(ret, g["ret"])
            assert np.array_equal(row_ptr.astype(np.int64), g["ptr"])
            got = np.stack([rec[k] for k in O.OVERLAP_DT.names], axis=1) if len(rec) else np.zeros((0, 10), np.int32)
            assert np.array_equal(got, g["ov"]), no_fast
            assert np.array_equal(ss.GetBaseCoverage(), g["cov"])
        finally:
            os.environ.pop("T1K_NO_FAST", None)


# ------------------------------------------------------------------------------------------------
# reads longer than 255 bases (tests/long_workloads.py, goldens from the unmodified reference)
import long_workloads as LW  # noqa: E402

LONG = {name: (recs, reads, sim, relax) for name, recs, reads, sim, relax in LW.cases()}


def _golden_long(name):
    z = np.load(os.path.join(HERE, "golden", "long", name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", sorted(LONG))
def test_oracle_matches_reference_on_long_reads(name):
    recs, reads, sim, relax = LONG[name]
    g = _golden_long(name)
    kept, w = O.collapse_reference(recs)
    orc = O.Oracle(kept, sim, relax, O.seq_weights(kept, w))
    for i, s in enumerate(reads):
        ret, ov = orc.assign(s, int(g["weight"]))
        assert ret == g["ret"][i], i
        got = np.stack([ov[k] for k in O.OVERLAP_DT.names], axis=1) if len(ov) else np.zeros((0, 10), np.int32)
        assert np.array_equal(got, g["ov"][g["ptr"][i]:g["ptr"][i + 1]]), i
    assert np.array_equal(np.concatenate([orc.coverage(a) for a in range(len(kept))]), g["cov"])


@pytest.mark.parametrize("name", sorted(LONG))
def test_lane_code_matches_reference_on_long_reads(emu, name):
    recs, reads, sim, relax = LONG[name]
    g = _golden_long(name)
    out, cov = _emu_run(emu, recs, reads, sim, relax, int(g["weight"]))
    for i, (n, rows) in enumerate(out):
        assert n == g["ret"][i], i
        assert np.array_equal(rows, g["ov"][g["ptr"][i]:g["ptr"][i + 1]]), i
    assert np.array_equal(cov, g["cov"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(LONG))
def test_device_matches_reference_on_long_reads(name):
    from t1k_b200.genotyper import SeqSet
    recs, reads, sim, relax = LONG[name]
    g = _golden_long(name)
    ref = RefSet(recs)
    ss = SeqSet(ref, sim, relax)
    a = ss.AssignRead(reads, [int(g["weight"])] * len(reads))
    row_ptr, ret, rec = a.fetch()
    assert np.array_equal(ret, g["ret"])
    assert np.array_equal(row_ptr.astype(np.int64), g["ptr"])
    got = np.stack([rec[k] for k in O.OVERLAP_DT.names], axis=1) if len(rec) else np.zeros((0, 10), np.int32)
    assert np.array_equal(got, g["ov"])
    assert np.array_equal(ss.GetBaseCoverage(), g["cov"])
    # a batch that mixes short and long reads takes the long-read geometry for all of them
    mixed = [reads[0][:120], reads[1], reads[2][:200]]
    b = ss.AssignRead(mixed)
    _, ret2, _ = b.fetch()
    assert ret2[1] == g["ret"][1]
