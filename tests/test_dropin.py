"""End-to-end drop-in check: integration/_build/genotyper_b200 (the reference's driver with phases A/B/C forwarded to
the C ABI) against the unmodified reference binary oracle/_ref/genotyper on the same FASTA/FASTQ files.  Every output
file must be byte-identical (the EM runs in the reference's summation order, so even the %lf abundances agree)."""
import filecmp
import os
import subprocess
import tempfile

import pytest

import workloads as W
from t1k_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "integration", "_build", "genotyper_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "genotyper")

pytestmark = pytest.mark.gpu

CASES = {
    "rna_pe": (lambda: W.small_rna_ref(seed=71), ["-s", "0.8"], dict(read_len=100, err=0.01, n_rate=0.002, indel_rate=0.05), False),
    "rna_hla_preset": (lambda: W.small_rna_ref(seed=72), ["-s", "0.97"], dict(read_len=150, err=0.003, n_rate=0.0, indel_rate=0.02, insert=(200, 420)), False),
    "dna_kir_wgs_preset": (lambda: W.small_dna_ref(seed=73), ["-s", "0.9", "--relaxIntronAlign"], dict(read_len=125, err=0.01, n_rate=0.002, indel_rate=0.05, insert=(200, 420)), False),
    "dna_se": (lambda: W.small_dna_ref(seed=74), ["-s", "0.9", "--relaxIntronAlign", "-n", "40"], dict(read_len=80, err=0.01, single_end=True), True),
}


def _hla_scale():
    """the bench configuration itself: 30,000 HLA-RNA-like alleles, 2,000 x 150 bp pairs, -s 0.97"""
    import bench
    recs, ref, r1, r2 = bench.make_workload(2000, 4711)
    return recs, r1, r2


@pytest.mark.parametrize("with_assign", [True, False], ids=["synchronous+assign_tsv", "pipelined"])
@pytest.mark.parametrize("name", sorted(CASES) + ["hla_scale_2000"])
def test_outputs_identical_to_reference_binary(name, with_assign):
    """with --outputReadAssignment the driver takes the synchronous entry points (t1k_assign_batch / t1k_pair_batch, rows in the
    reference's order, the reference's own CoalesceReadAssignments); without it the pipelined t1k_genotype call."""
    if not (os.path.exists(OURS) and os.path.exists(REF)):
        pytest.skip("integration driver / reference binary not built (needs the T1K checkout at build time)")
    if name == "hla_scale_2000":
        recs, r1, r2 = _hla_scale()
        flags, single = ["-s", "0.97"], False
    else:
        factory, flags, kw, single = CASES[name]
        recs = factory()
        r1, r2 = W.reads_for(recs, 500, seed=75, **kw)
    td = tempfile.mkdtemp(prefix="t1kdrop_")
    fa = os.path.join(td, "ref.fa")
    synth.write_fasta(fa, recs)
    p1, p2 = os.path.join(td, "r_1.fq"), os.path.join(td, "r_2.fq")
    synth.write_fastq(p1, r1)
    if not single:
        synth.write_fastq(p2, r2)
    outs = {}
    for tag, exe in (("ref", REF), ("ours", OURS)):
        prefix = os.path.join(td, tag)
        cmd = [exe, "-f", fa] + (["-u", p1] if single else ["-1", p1, "-2", p2]) + ["-o", prefix, "-t", str(os.cpu_count() or 2)] + flags
        if with_assign:
            cmd.append("--outputReadAssignment")
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert res.returncode == 0, res.stderr[-2000:]
        outs[tag] = (prefix, res.stderr)
    suffixes = ["_genotype.tsv", "_allele.tsv"] + (["_assign.tsv"] if with_assign else []) + (["_aligned.fa"] if single else ["_aligned_1.fa", "_aligned_2.fa"])
    for sfx in suffixes:
        a, b = outs["ref"][0] + sfx, outs["ours"][0] + sfx
        assert os.path.exists(a) and os.path.exists(b), sfx
        assert filecmp.cmp(a, b, shallow=False), "%s differs:\n%s" % (sfx, subprocess.run(["diff", a, b], stdout=subprocess.PIPE, text=True).stdout[:2000])
    assert os.path.getsize(outs["ref"][0] + "_genotype.tsv") > 0
    # the log lines a wrapper may parse: same counts, same EM iteration count
    def tail(s):
        return [l.split("] ", 1)[1] for l in s.strip().split("\n") if "] " in l]
    assert tail(outs["ours"][1]) == tail(outs["ref"][1])


def test_cli_error_behaviour():
    if not os.path.exists(OURS):
        pytest.skip("integration driver not built")
    assert subprocess.run([OURS], stderr=subprocess.PIPE).returncode == 0                      # usage, exit 0 (Genotyper.cpp:199-203)
    assert subprocess.run([OURS, "--nope"], stderr=subprocess.PIPE).returncode != 0            # unknown flag (Genotyper.cpp:321-325)
    assert subprocess.run([OURS, "-u", "/dev/null"], stderr=subprocess.PIPE).returncode != 0   # no -f (Genotyper.cpp:327-331)
