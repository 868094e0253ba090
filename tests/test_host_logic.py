"""CPU-only checks of the product's host logic and of the lane-level device code compiled for the host
(tests/host_emu.cpp runs t1k_core.cuh sequentially — test harness only) against the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import golden_io as G
import oracle_py as O
import workloads as W
from t1k_b200 import _lib as L
from t1k_b200.refset import RefSet, parse_allele_name, parse_exons

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", G.names())
def test_lane_code_matches_reference_golden(emu, name):
    """chaining, diagonal certificate, banded DP, extension and full-read alignment of t1k_core.cuh reproduce the
    reference's AssignRead records and base coverage"""
    g = G.load(name)
    ref = RefSet(g["records"])
    bases, off, ptr, se = ref.packed()
    E = emu.emu_create(ref.n, bases, O._p(off), O._p(ptr), O._p(se), g["similarity"], int(g["relax"]))
    assert E
    buf = np.zeros(1 << 14, dtype=O.OVERLAP_DT)
    for i, s in enumerate(g["uniq_seq"]):
        err = C.c_int32(0)
        n = emu.emu_assign(E, s, int(g["uniq_weight"][i]), O._p(buf), len(buf), C.byref(err))
        assert err.value == 0
        want = G.uniq_overlaps(g, i)
        got = np.stack([buf[k][:max(n, 0)] for k in O.OVERLAP_DT.names], axis=1) if n > 0 else np.zeros((0, 10), np.int32)
        assert np.array_equal(got, want), i
    cov = []
    for a in range(ref.n):
        out = np.zeros(len(ref.seqs[a]), dtype=np.int32)
        emu.emu_coverage(E, a, O._p(out))
        cov.append(out)
    assert np.array_equal(np.concatenate(cov), g["cov"])
    emu.emu_destroy(E)


def _emu_assign_all(emu, E, reads, weight=1):
    buf = np.zeros(1 << 16, dtype=O.OVERLAP_DT)
    out = []
    for s in reads:
        err = C.c_int32(0)
        n = emu.emu_assign(E, s, weight, O._p(buf), len(buf), C.byref(err))
        assert err.value == 0
        out.append((n, np.stack([buf[k][:max(n, 0)] for k in O.OVERLAP_DT.names], axis=1).copy()))
    return out


def _ab_workloads():
    from t1k_b200 import synth
    rng = np.random.default_rng(77)
    # (records, reads, similarity, relax)
    rna = W.small_rna_ref(seed=21)
    r1, r2 = W.reads_for(rna, 150, read_len=150, seed=22, err=0.004, n_rate=0.0005, indel_rate=0.02, insert=(200, 400))
    yield "rna150", rna, [r.tobytes() for r in r1] + [r.tobytes() for r in r2], 0.97, False
    dna = W.small_dna_ref(seed=23)
    r1, r2 = W.reads_for(dna, 150, read_len=100, seed=24, err=0.01, n_rate=0.002, indel_rate=0.05)
    yield "dna100_relax", dna, [r.tobytes() for r in r1] + [r.tobytes() for r in r2], 0.9, True
    # low-complexity / homopolymer / tandem-repeat alleles and reads: the cases the eligibility rule must catch
    base = bytearray(synth.make_hla_rna_ref(genes=[("G", 1)], length=600, n_sites=10, min_sub=0, max_sub=0, seed=25)[0][2])
    base[100:130] = b"A" * 30
    base[200:240] = b"AC" * 20
    base[300:345] = b"ACG" * 15
    base[400:415] = b"T" * 15
    recs = []
    for i in range(40):
        s = bytearray(base)
        for _ in range(int(rng.integers(0, 6))):
            s[int(rng.integers(0, len(s)))] = b"ACGT"[int(rng.integers(0, 4))]
        if i % 7 == 3:
            s = s[:int(rng.integers(300, 600))]           # truncated alleles: clips at the allele end
        recs.append(("G*%02d:01" % i, "", bytes(s)))
    reads = []
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    for _ in range(400):
        a = recs[int(rng.integers(0, len(recs)))][2]
        L = int(rng.integers(40, 200))
        st = int(rng.integers(-20, len(a) - 20))
        r = bytearray(a[max(st, 0):max(st, 0) + L])
        if st < 0:
            r = bytearray(bytes(rng.choice(list(b"ACGT"), size=-st).astype(np.uint8))) + r     # overhang before the allele start
        if len(r) < 20:
            continue
        for _ in range(int(rng.integers(0, 5))):
            r[int(rng.integers(0, len(r)))] = b"ACGTN"[int(rng.integers(0, 5 if rng.integers(0, 4) == 0 else 4))]
        if rng.integers(0, 10) == 0 and len(r) > 30:
            k = int(rng.integers(5, len(r) - 5)); del r[k:k + int(rng.integers(1, 4))]
        r = bytes(r)
        reads.append(r.translate(comp)[::-1] if rng.integers(0, 2) else r)
    yield "repeats", recs, reads, 0.8, False


def test_diagonal_fast_path_equals_general_path(emu):
    """diag_fast (mismatch-mask path of t1k_core.cuh) and chain_allele + extend_cand + full_align (hit-list path) give the
    same AssignRead records and the same coverage on every workload; the fast path must actually be taken."""
    for name, recs, reads, sim, relax in _ab_workloads():
        ref = RefSet(recs)
        bases, off, ptr, se = ref.packed()
        res, covs, taken = [], [], []
        for fast in (1, 0):
            E = emu.emu_create(ref.n, bases, O._p(off), O._p(ptr), O._p(se), sim, int(relax))
            emu.emu_set_fast(E, fast)
            c0 = emu.emu_counters()[17]
            res.append(_emu_assign_all(emu, E, reads, weight=2))
            taken.append(emu.emu_counters()[17] - c0)
            cov = []
            for a in range(ref.n):
                out = np.zeros(len(ref.seqs[a]), dtype=np.int32)
                emu.emu_coverage(E, a, O._p(out))
                cov.append(out)
            covs.append(np.concatenate(cov))
            emu.emu_destroy(E)
        assert taken[0] > 0 and taken[1] == 0, (name, taken)
        for i, ((n1, a), (n0, b)) in enumerate(zip(*res)):
            assert n1 == n0 and np.array_equal(a, b), (name, i, reads[i])
        assert np.array_equal(covs[0], covs[1]), name
        # ... and both equal the oracle (a comparison of the product with itself proves nothing about the reference)
        kept, _ = O.collapse_reference(recs)
        assert len(kept) == ref.n
        orc = O.Oracle(kept, sim, relax)
        for i, s in enumerate(reads):
            oret, ov = orc.assign(s, 2)
            want = np.stack([ov[k] for k in O.OVERLAP_DT.names], axis=1) if len(ov) else np.zeros((0, 10), np.int32)
            assert res[0][i][0] == oret and np.array_equal(res[0][i][1], want), (name, i, s)
        assert np.array_equal(covs[0], np.concatenate([orc.coverage(a) for a in range(ref.n)])), name


def test_lane_code_matches_oracle_on_the_bench_workload(emu):
    """The bench configuration itself (30,000 HLA-RNA-like alleles x 1,100 bp, 150 bp reads, -s 0.97; ~3.5 k records per
    read-end, the > 1000 cut active): AssignRead records and coverage of the product's lane code == the oracle's."""
    import bench
    recs, ref, r1, r2 = bench.make_workload(20, 321)
    kept, w = O.collapse_reference(recs)
    orc = O.Oracle(kept, 0.97, False, O.seq_weights(kept, w))
    bases, off, ptr, se = ref.packed()
    E = emu.emu_create(ref.n, bases, O._p(off), O._p(ptr), O._p(se), 0.97, 0)
    buf = np.zeros(1 << 16, dtype=O.OVERLAP_DT)
    n_rec = 0
    for i, s in enumerate([r.tobytes() for r in r1] + [r.tobytes() for r in r2]):
        err = C.c_int32(0)
        n = emu.emu_assign(E, s, 2, O._p(buf), len(buf), C.byref(err))
        assert err.value == 0
        oret, ov = orc.assign(s, 2)
        assert n == oret, i
        got = np.stack([buf[k][:max(n, 0)] for k in O.OVERLAP_DT.names], axis=1) if n > 0 else np.zeros((0, 10), np.int32)
        want = np.stack([ov[k] for k in O.OVERLAP_DT.names], axis=1) if len(ov) else np.zeros((0, 10), np.int32)
        assert np.array_equal(got, want), i
        n_rec += max(n, 0)
    assert n_rec > 50000
    cov = []
    for a in range(ref.n):
        out = np.zeros(len(ref.seqs[a]), dtype=np.int32)
        emu.emu_coverage(E, a, O._p(out))
        cov.append(out)
    assert np.array_equal(np.concatenate(cov), np.concatenate([orc.coverage(k) for k in range(ref.n)]))
    emu.emu_destroy(E)


def test_oracle_and_lane_code_match_the_reference_on_the_bench_configuration(emu):
    """tests/golden/hla_scale: AssignRead records of the UNMODIFIED reference on the 30,000-allele bench reference (the > 1000
    cut of SeqSet.hpp:2290-2298 is active on every read-end).  The oracle and the product's lane code reproduce them."""
    import bench
    g = G.load_hla_scale()
    recs, ref, r1, r2 = bench.make_workload(g["n_pairs"], g["seed"])
    assert np.array_equal(r1, g["reads1"]) and np.array_equal(r2, g["reads2"])       # the generator still makes the golden's inputs
    kept, w = O.collapse_reference(recs)
    orc = O.Oracle(kept, 0.97, False, O.seq_weights(kept, w))
    bases, off, ptr, se = ref.packed()
    E = emu.emu_create(ref.n, bases, O._p(off), O._p(ptr), O._p(se), 0.97, 0)
    buf = np.zeros(1 << 16, dtype=O.OVERLAP_DT)
    assert len(g["uniq_seq"]) > 20 and g["uniq_ptr"][-1] > 100000
    for i, s in enumerate(g["uniq_seq"]):
        want = g["uniq_ov"][g["uniq_ptr"][i]:g["uniq_ptr"][i + 1]]
        oret, ov = orc.assign(s, int(g["uniq_weight"][i]))
        got_o = np.stack([ov[k] for k in O.OVERLAP_DT.names], axis=1) if len(ov) else np.zeros((0, 10), np.int32)
        assert np.array_equal(got_o, want), ("oracle", i)
        err = C.c_int32(0)
        n = emu.emu_assign(E, s, int(g["uniq_weight"][i]), O._p(buf), len(buf), C.byref(err))
        assert err.value == 0
        got = np.stack([buf[k][:max(n, 0)] for k in O.OVERLAP_DT.names], axis=1) if n > 0 else np.zeros((0, 10), np.int32)
        assert np.array_equal(got, want), ("lane code", i)
    emu.emu_destroy(E)
    # pairing, coalescing, equivalence classes and EM of the oracle against the reference's whole-flow outputs
    orc.coverage_reset()
    R = O.genotype_pipeline(orc, r1, r2, ref.names, O.seq_weights(kept, w))
    assert R["assigned"] == g["aligned"] and len(R["groups"]) == g["n_groups"] and len(R["ecs"]) == g["n_ec"]
    assert R["iters"] == g["iters"]
    assert np.array_equal(R["allele_ec"], g["q"][:, 0].astype(np.int32))
    assert np.array_equal(R["abundance"], g["q"][:, 1])


def test_banded_dp_and_diagonal_certificate(emu):
    """dp_align == AlignAlgo::GlobalAlignment op for op (oracle) on random pairs incl. N, indels and band edges;
    whenever the diagonal certificate fires, the reference alignment is the pure diagonal."""
    rng = np.random.default_rng(12)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    certified = 0
    for it in range(1500):
        n = int(rng.integers(1, 150))
        t = alpha[rng.integers(0, 4, size=n)].copy()
        if it % 3 == 0:                       # repetitive sequence stresses tie-breaking
            unit = alpha[rng.integers(0, 4, size=int(rng.integers(1, 5)))]
            t = np.resize(unit, n).copy()
        p = t.copy()
        for _ in range(int(rng.integers(0, 7))):
            p[rng.integers(0, len(p))] = alpha[rng.integers(0, 4)]
        if it % 4 == 1 and len(p) > 12:
            k = int(rng.integers(1, len(p) - 1))
            d = int(rng.integers(1, 5))
            p = np.delete(p, slice(k, k + d)) if rng.integers(0, 2) else np.insert(p, k, alpha[rng.integers(0, 4, size=d)])
        if it % 5 == 2:
            p[rng.integers(0, len(p))] = ord("N")
            t[rng.integers(0, len(t))] = ord("N")
        if len(p) == 0 or len(p) > 250:
            continue
        tb, pb = t.tobytes(), p.tobytes()
        score, ops = O.global_alignment(tb, pb)
        out = np.zeros(len(tb) + len(pb) + 16, dtype=np.int8)
        cert, matches = C.c_int32(0), C.c_int32(0)
        k = emu.emu_align(tb, len(tb), pb, len(pb), O._p(out), C.byref(cert), C.byref(matches))
        assert k == len(ops) and np.array_equal(out[:k], ops), (tb, pb)
        assert matches.value == int((ops == 0).sum())
        if cert.value:
            certified += 1
            assert len(tb) == len(pb) and set(ops.tolist()) <= {0, 1}
    assert certified > 300


def test_diagonal_certificates_are_sound_on_clustered_mismatches(emu):
    """Equal-length pairs with 3-13 clustered substitutions, periodic / low-complexity sequence and displaced segments (the
    inputs where a gapped path can beat the diagonal): dp_align_eq == GlobalAlignment op for op, and whenever a
    certificate (<= 3, exact 4-5 enumeration, shift histogram, interval) fires the reference alignment has no indel."""
    rng = np.random.default_rng(99)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    certified_by_mm = {}
    for it in range(5000):
        n = int(rng.integers(8, 170))
        mode = it % 4
        if mode == 0:
            t = alpha[rng.integers(0, 4, size=n)].copy()
        elif mode == 1:
            t = np.resize(alpha[rng.integers(0, 4, size=int(rng.integers(1, 7)))], n).copy()
        elif mode == 2:
            t = alpha[rng.integers(0, 2, size=n)].copy()
        else:
            t = np.resize(alpha[rng.integers(0, 4, size=int(rng.integers(2, 12)))], n).copy()
            for _ in range(int(rng.integers(0, 4))):
                t[rng.integers(0, n)] = alpha[rng.integers(0, 4)]
        p = t.copy()
        if it % 5 == 0 and n > 20:             # a displaced segment: the indel-favouring case
            k = int(rng.integers(1, n - 12)); d = int(rng.integers(1, 5)); k2 = int(rng.integers(k, n - d))
            p = np.concatenate([p[:k], p[k + d:k2 + d], alpha[rng.integers(0, 4, size=d)], p[k2 + d:]])[:n]
            if len(p) < n:
                p = np.concatenate([p, alpha[rng.integers(0, 4, size=n - len(p))]])
        c0 = int(rng.integers(0, n))
        for _ in range(int(rng.integers(3, 14))):
            j = int(np.clip(c0 + rng.integers(-12, 13), 0, n - 1)) if it % 2 else int(rng.integers(0, n))
            p[j] = alpha[rng.integers(0, 4)]
        tb, pb = t.tobytes(), p.tobytes()
        _, ops = O.global_alignment(tb, pb)
        out = np.zeros(2 * n + 16, dtype=np.int8)
        cert, matches = C.c_int32(0), C.c_int32(0)
        k = emu.emu_align(tb, n, pb, n, O._p(out), C.byref(cert), C.byref(matches))
        assert k == len(ops) and np.array_equal(out[:k], ops), (tb, pb)
        assert matches.value == int((ops == 0).sum())
        if cert.value:
            mm = int((t != p).sum())
            certified_by_mm[mm] = certified_by_mm.get(mm, 0) + 1
            assert set(ops.tolist()) <= {0, 1}, (tb, pb)
    assert sum(v for m, v in certified_by_mm.items() if m >= 6) > 500      # the interval certificate is exercised


def test_threaded_host_model_equals_single_thread():
    """tests/host_model_check.cpp: sharded coalescing + gather, CSR transposition, equivalence classes, EM inputs and the
    rank-table merge of t1k_model.hpp give, for every thread count, exactly the single-threaded result (and the sharded
    coalescing gives the plain fragment-order one)."""
    src = os.path.join(ROOT, "tests", "host_model_check.cpp")
    so = os.path.join(ROOT, "tests", "_build", "libhostmodelcheck.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    deps = [src, os.path.join(ROOT, "t1k_b200", "csrc", "t1k_model.hpp")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-w", "-std=c++14", "-pthread", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    lib.host_model_check.restype = C.c_int
    lib.host_model_check.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    for seed, n_frag, n_alleles, n_sets in ((1, 4000, 300, 60), (2, 20000, 2000, 700), (3, 300, 50, 5), (4, 2, 10, 1)):
        assert lib.host_model_check(seed, n_frag, n_alleles, n_sets) == 0, (seed, n_frag)
    # the two-phase read-end de-duplication of the chunk pipeline (unique_read_ends)
    lib.dedup_check.restype = C.c_int
    lib.dedup_check.argtypes = [C.c_int, C.c_int, C.c_int]
    for seed, n_frag, paired in ((1, 9000, 1), (2, 5000, 0), (3, 40, 1), (4, 1, 1)):
        assert lib.dedup_check(seed, n_frag, paired) == 0, (seed, n_frag, paired)


def test_parse_helpers():
    assert parse_exons("7 50 221 623 783", 3000) == [(50, 221), (623, 783)]
    assert parse_exons("", 100) == [(0, 99)]
    assert parse_exons("1 0 1099", 1100) == [(0, 1099)]
    assert parse_allele_name("HLA-A*01:02:03:04") == ("HLA-A", "HLA-A*01:02:03")
    assert parse_allele_name("HLA-A*01:02") == ("HLA-A", "HLA-A*01:02")
    assert parse_allele_name("KIR2DL1*0010101") == ("KIR2DL1", "KIR2DL1*001")
    assert parse_allele_name("CYP2D6*4.001", 1, ".") == ("CYP2D6", "CYP2D6*4")
    for name in ("HLA-B*07:02:01", "KIR3DL2*00701", "X", "CYP2D6*10.002"):
        for du, dl in ((-1, ""), (1, "."), (2, ":")):
            assert parse_allele_name(name, du, dl) == O.parse_allele_name(name, du, dl)


@pytest.mark.parametrize("name", G.names())
def test_refset_matches_reference_golden(name):
    g = G.load(name)
    ref = RefSet(g["records"])
    q = g["q"]
    assert ref.n == len(q)
    assert np.array_equal(ref.effective_len, q[:, 3].astype(np.int32))
    assert np.array_equal(ref.seq_weight, q[:, 4].astype(np.int32))


def test_abi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "t1k_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(t1k_[a-z_0-9]+)\s*\(", hdr)))
    assert sorted(L.EXPORTS) == declared
    lib = L.lib()
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the compute entry points fail loudly (T1K_ERR_NO_DEVICE); nothing runs on the CPU."""
    n = C.c_int(0)
    L.lib().t1k_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a CUDA device is present")
    from t1k_b200.genotyper import QuantifyAlleleEquivalentClass, SeqSet
    ref = RefSet(W.small_rna_ref())
    with pytest.raises(L.T1KError) as e:
        SeqSet(ref, 0.8, False)
    assert e.value.code == L.T1K_ERR_NO_DEVICE
    with pytest.raises(L.T1KError) as e:
        QuantifyAlleleEquivalentClass([0, 1], [0], [1.0], [100], [1.0])
    assert e.value.code == L.T1K_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "t1k_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "_never_.py", os.path.join(dirpath, f)
