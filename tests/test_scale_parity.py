"""BASELINE configs 3 and 4 at their own scale (tests/golden/scale/, made by tests/golden/make_golden_scale.py from the
UNMODIFIED reference): KIR-DNA-like reference with N separators, -s 0.9 --relaxIntronAlign, paired-end; HLA-DNA-like reference
(30,000 alleles x ~3.5 kb), single-end 100 bp, -s 0.97.  CPU: the oracle and the lane-code emulation reproduce the reference's
AssignRead records and whole-flow outputs; GPU: the device path through the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

import bench
import oracle_py as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(config):
    z = np.load(os.path.join(HERE, "golden", "scale", "config%d.npz" % config))
    g = {k: z[k] for k in z.files}
    g["uniq_seq"] = [s.encode() if isinstance(s, str) else bytes(s) for s in g["uniq_seq"].tolist()]
    for k in ("config", "n", "seed", "iters", "aligned", "n_groups", "n_ec"):
        g[k] = int(g[k])
    if g["reads2"].size == 0:
        g["reads2"] = None
    return g


@pytest.mark.parametrize("config", [3, 4])
def test_oracle_and_lane_code_match_the_reference_at_config_scale(emu, config):
    g = _load(config)
    cfg = bench.CONFIGS[config]
    recs, ref = bench.make_reference(config)
    kept, w = O.collapse_reference(recs)
    assert len(kept) == ref.n
    sw = O.seq_weights(kept, w)
    orc = O.Oracle(kept, cfg["sim"], cfg["relax"], sw)
    bases, off, ptr, se = ref.packed()
    E = emu.emu_create(ref.n, bases, O._p(off), O._p(ptr), O._p(se), cfg["sim"], int(cfg["relax"]))
    buf = np.zeros(1 << 16, dtype=O.OVERLAP_DT)
    step = max(1, len(g["uniq_seq"]) // 60)              # every read-end through the emulation, a sample through the (slow) oracle
    for i, s in enumerate(g["uniq_seq"]):
        want = g["uniq_ov"][g["uniq_ptr"][i]:g["uniq_ptr"][i + 1]]
        err = C.c_int32(0)
        n = emu.emu_assign(E, s, int(g["uniq_weight"][i]), O._p(buf), len(buf), C.byref(err))
        assert err.value == 0
        got = np.stack([buf[k][:max(n, 0)] for k in O.OVERLAP_DT.names], axis=1) if n > 0 else np.zeros((0, 10), np.int32)
        assert np.array_equal(got, want), ("lane code", i)
        if i % step == 0:
            _, ov = orc.assign(s, int(g["uniq_weight"][i]))
            got_o = np.stack([ov[k] for k in O.OVERLAP_DT.names], axis=1) if len(ov) else np.zeros((0, 10), np.int32)
            assert np.array_equal(got_o, want), ("oracle", i)
    emu.emu_destroy(E)


@pytest.mark.gpu
@pytest.mark.parametrize("no_share", ["0", "1"])
@pytest.mark.parametrize("config", [3, 4])
def test_device_matches_the_reference_at_config_scale(config, no_share, monkeypatch):
    """no_share = 1: every deferred allele / full-read alignment evaluated on its own instead of once per group of alleles with
    identical window content (T1K_NO_SHARE, the A/B switch of the sharing in k_deferred / k_align)"""
    from t1k_b200.genotyper import Genotyper
    monkeypatch.setenv("T1K_NO_SHARE", no_share)
    g = _load(config)
    cfg = bench.CONFIGS[config]
    recs, ref = bench.make_reference(config)
    out = Genotyper(ref, cfg["sim"], cfg["relax"]).Genotype(g["reads1"], g["reads2"])
    q = g["q"]
    assert out["assigned_fragments"] == g["aligned"]
    assert out["n_groups"] == g["n_groups"] and out["n_ec"] == g["n_ec"]
    assert out["em_iterations"] == g["iters"]
    assert np.array_equal(out["equivalent_class"], q[:, 0].astype(np.int32))
    assert np.array_equal(out["missing_coverage"], g["missing"])
    assert np.array_equal(out["abundance"], q[:, 1])                  # reference-order EM: the reference's doubles
    # the AssignRead records themselves
    from t1k_b200.genotyper import SeqSet
    ss = SeqSet(ref, cfg["sim"], cfg["relax"])
    a = ss.AssignRead(g["uniq_seq"], g["uniq_weight"])
    row_ptr, ret, rec = a.fetch()
    got = np.stack([rec[k] for k in O.OVERLAP_DT.names], axis=1) if len(rec) else np.zeros((0, 10), np.int32)
    assert np.array_equal(row_ptr.astype(np.int64), g["uniq_ptr"])
    assert np.array_equal(got, g["uniq_ov"])
