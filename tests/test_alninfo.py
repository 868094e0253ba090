"""SURVEY.md 8f N3 — SeqSet::AddOverlapAlignmentInfo (SeqSet.hpp:2657-2680): the edit strings the analyzer attaches to the
overlaps it kept.  Goldens (tests/golden/alninfo, tests/golden/make_golden_alninfo.py) are the UNMODIFIED reference's strings
for every record of the six golden workloads; checked here against the oracle, the product's lane code on the CPU
(host emulation) and — `-m gpu` — the device path through t1k_align_info_batch, each with the certified-diagonal shortcut on and off."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_io as G
import oracle_py as O
from t1k_b200.refset import RefSet

_RC = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _RC[_a] = _b


def load_alninfo(name):
    z = np.load(os.path.join(G.GOLDEN, "alninfo", name + ".npz"))
    return z["ops"], z["ops_ptr"]


def items_of(g):
    """(read index, record) of every golden record, in the order of the golden's ops_ptr"""
    idx, recs = [], []
    for i in range(len(g["uniq_seq"])):
        ov = G.uniq_overlaps(g, i)
        idx += [i] * len(ov)
        recs.append(ov)
    rec = np.concatenate(recs) if recs else np.zeros((0, 10), np.int32)
    out = np.zeros(len(rec), dtype=O.OVERLAP_DT)
    for k, name in enumerate(O.OVERLAP_DT.names):
        out[name] = rec[:, k]
    return np.asarray(idx, dtype=np.uint32), out


def slices(g, ref, idx, ov, k):
    read = np.frombuffer(g["uniq_seq"][int(idx[k])], dtype=np.uint8)
    if ov["strand"][k] == -1:
        read = _RC[read[::-1]]
    t = ref.seqs[int(ov["seqIdx"][k])][int(ov["seqStart"][k]):int(ov["seqEnd"][k]) + 1]
    p = read[int(ov["readStart"][k]):int(ov["readEnd"][k]) + 1].tobytes()
    return bytes(t), p


@pytest.mark.parametrize("name", G.names())
def test_oracle_edit_strings_match_the_reference(name):
    """the oracle's GlobalAlignment on the record's slices == the reference's AddOverlapAlignmentInfo string"""
    g = G.load(name)
    ref = RefSet(g["records"])
    ops, ptr = load_alninfo(name)
    idx, ov = items_of(g)
    assert len(ptr) == len(ov) + 1
    step = max(1, len(ov) // 4000)            # every record with an indel + a stride of the rest
    indel = np.zeros(len(ov), dtype=bool)
    bad = np.flatnonzero((ops == 2) | (ops == 3))
    indel[np.unique(np.searchsorted(ptr, bad, side="right") - 1)] = True
    for k in range(len(ov)):
        if not indel[k] and k % step:
            continue
        t, p = slices(g, ref, idx, ov, k)
        _, got = O.global_alignment(t, p)
        assert np.array_equal(got, ops[ptr[k]:ptr[k + 1]]), (name, k)


@pytest.mark.parametrize("name", G.names())
def test_lane_code_edit_strings_match_the_reference(emu, name):
    """align_info of t1k_core.cuh (CPU emulation) == the reference, with the certified-diagonal shortcut and through the band DP"""
    g = G.load(name)
    ref = RefSet(g["records"])
    bases, off, eptr, se = ref.packed()
    E = emu.emu_create(ref.n, bases, O._p(off), O._p(eptr), O._p(se), g["similarity"], int(g["relax"]))
    assert E
    ops, ptr = load_alninfo(name)
    idx, ov = items_of(g)
    out = np.zeros(4096, dtype=np.int8)
    n_diag = 0
    step = max(1, len(ov) // 3000)
    for k in range(0, len(ov)):
        want = ops[ptr[k]:ptr[k + 1]]
        has_indel = bool(((want == 2) | (want == 3)).any())
        if not has_indel and k % step:
            continue
        rec = ov[k:k + 1]
        for no_diag in (0, 1):
            ran = C.c_int32(0)
            n = emu.emu_align_info(E, g["uniq_seq"][int(idx[k])], O._p(rec), no_diag, O._p(out), C.byref(ran))
            assert n == len(want) and np.array_equal(out[:n], want), (name, k, no_diag)
            if no_diag == 0 and not ran.value:
                n_diag += 1
                assert not has_indel
    assert n_diag > 100
    emu.emu_destroy(E)


# ------------------------------------------------------------------------------------------------ device
@pytest.mark.gpu
@pytest.mark.parametrize("name", G.names())
def test_device_edit_strings_match_the_reference(name):
    from t1k_b200.genotyper import SeqSet
    g = G.load(name)
    ref = RefSet(g["records"])
    ss = SeqSet(ref, g["similarity"], g["relax"])
    ops, ptr = load_alninfo(name)
    idx, ov = items_of(g)
    # the records themselves come from the device too (weight 0 = the analyzer's call, Analyzer.cpp:476)
    a = ss.AssignRead(g["uniq_seq"], np.zeros(len(g["uniq_seq"]), dtype=np.int32))
    row, ret, rec = a.fetch()
    assert np.array_equal(rec, ov)
    for force_dp in (False, True):
        got, st = ss.AddOverlapAlignmentInfo(g["uniq_seq"], idx, rec, force_dp=force_dp, with_stats=True)
        assert st["n_diagonal"] + st["n_dp"] == len(ov)
        assert (st["n_diagonal"] == 0) == force_dp
        assert st["dp_cells"] > 0
        for k in range(len(ov)):
            assert np.array_equal(got[k], ops[ptr[k]:ptr[k + 1]]), (name, k, force_dp)
    assert not ss.GetBaseCoverage().any()


@pytest.mark.gpu
def test_device_edit_strings_long_reads_and_edges():
    """reads of 300 and 1000 bases with indels and N against the oracle; seqIdx == -1 items; bad coordinates; empty batch"""
    import long_workloads as LW
    from t1k_b200._lib import OVERLAP_DT, T1KError
    from t1k_b200.genotyper import SeqSet
    n_indel = 0
    for name, recs, reads, sim, relax in LW.cases():
        if name not in ("rna_300", "dna_320_relax", "rna_1000"):
            continue
        ref = RefSet(recs)
        ss = SeqSet(ref, sim, relax)
        a = ss.AssignRead(reads, np.zeros(len(reads), dtype=np.int32))
        row, ret, rec = a.fetch()
        idx = np.repeat(np.arange(len(reads), dtype=np.uint32), np.diff(row).astype(np.int64))
        assert len(rec) > 0, name
        got = ss.AddOverlapAlignmentInfo(reads, idx, rec)
        for k in range(0, len(rec), max(1, len(rec) // 400)):
            read = np.frombuffer(reads[int(idx[k])], dtype=np.uint8)
            if rec["strand"][k] == -1:
                read = _RC[read[::-1]]
            t = bytes(ref.seqs[int(rec["seqIdx"][k])][int(rec["seqStart"][k]):int(rec["seqEnd"][k]) + 1])
            _, want = O.global_alignment(t, read[int(rec["readStart"][k]):int(rec["readEnd"][k]) + 1].tobytes())
            assert np.array_equal(got[k], want), (name, k)
            n_indel += int(((want == 2) | (want == 3)).any())
    assert n_indel > 0
    # seqIdx == -1: no string, as the reference leaves `align` unset
    one = np.zeros(2, dtype=OVERLAP_DT)
    one[0] = rec[0]
    one[1]["seqIdx"] = -1
    got = ss.AddOverlapAlignmentInfo(reads, idx[:2], one)
    assert got[1] is None and got[0] is not None
    bad = rec[:1].copy()
    bad["seqEnd"] = 10 ** 6
    with pytest.raises(T1KError):
        ss.AddOverlapAlignmentInfo(reads, idx[:1], bad)
    assert ss.AddOverlapAlignmentInfo(reads, np.zeros(0, np.uint32), np.zeros(0, dtype=OVERLAP_DT)) == []


@pytest.mark.gpu
def test_dpx_peak_is_measurable():
    from t1k_b200 import _lib as L
    g = C.c_double(0)
    L.check(L.lib().t1k_dpx_peak(-1, C.byref(g)))
    assert g.value > 1000.0          # G ops/s; a B200 issues tens of T ops/s
