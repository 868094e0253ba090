"""SURVEY.md 8f N1 (next row, oracle stage): the oracle's restatement of fastq-extractor's candidate filter
(IsLowComplexity + SeqSet::HasHitInSet with the extractor's own k / hitLenRequired set-up) against what the UNMODIFIED
reference binary kept (tests/golden/filter/*.npz, made by tests/golden/make_golden_filter.py)."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

import golden_io as G
import oracle_py as O

FILTER_DIR = os.path.join(G.GOLDEN, "filter")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(FILTER_DIR, "*.npz")))


@pytest.mark.parametrize("name", NAMES)
def test_candidate_filter_matches_reference_binary(name):
    z = np.load(os.path.join(FILTER_DIR, name + ".npz"))
    if name.startswith("recipe_"):                      # large reference: regenerated from the recipe of make_golden_filter.py
        import importlib.util
        spec = importlib.util.spec_from_file_location("make_golden_filter", os.path.join(G.GOLDEN, "make_golden_filter.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        recs = mod.big_ref()
    else:
        recs = G.parse_fasta_bytes(z["fasta"].tobytes())
    paired = bool(int(z["paired"]))
    r1 = [bytes(x) for x in z["reads1"]]
    r2 = [bytes(x) for x in z["reads2"]] if paired else None
    f = O.CandidateFilter(recs, r1, paired, float(z["similarity"]))
    if name == "recipe_k13":
        assert f.k == 13                                # the sorted-code index of the oracle (k > 12)
    got = np.asarray([f.keep_pair(r1[i], r2[i] if paired else None) for i in range(len(r1))], dtype=np.uint8)
    want = z["kept"]
    assert 0 < want.sum() < len(want)                  # the fixture has both outcomes
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:10]


def test_kmer_length_inference_and_low_complexity():
    L = O.lib()
    # SeqSet::InferKmerLength (SeqSet.hpp:2830-2845): number of base-4 digits of the total length + 1
    assert L.t1ko_infer_kmer_length(0) == 1
    assert L.t1ko_infer_kmer_length(3) == 2
    assert L.t1ko_infer_kmer_length(98000) == 10
    assert L.t1ko_infer_kmer_length(33_000_000) == 14
    # IsLowComplexity (FastqExtractor.cpp:89-112)
    assert L.t1ko_is_low_complexity(b"A" * 60) == 1
    assert L.t1ko_is_low_complexity(b"ACGT" * 20) == 0
    assert L.t1ko_is_low_complexity(b"ACACACACACACACACACACACAC") == 1            # two letters absent
    assert L.t1ko_is_low_complexity(b"ACGTNNNNNNNNNNACGTACGTACGTACGTACGT") == 1   # >= 10 % N


@pytest.fixture(scope="module")
def femu():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "filter_emu.cpp")
    so = os.path.join(root, "tests", "_build", "libfilteremu.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    deps = [src] + [os.path.join(root, "t1k_b200", "csrc", f) for f in ("t1k_filter_lane.cuh", "t1k_core.cuh", "t1k_host.hpp")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-w", "-std=c++14", "-x", "c++", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    lib.femu_create.restype = C.c_void_p
    lib.femu_create.argtypes = [C.c_int32, C.c_char_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double]
    lib.femu_destroy.argtypes = [C.c_void_p]
    lib.femu_good_candidate.restype = C.c_int32
    lib.femu_good_candidate.argtypes = [C.c_void_p, C.c_char_p]
    return lib


@pytest.mark.parametrize("name", NAMES)
def test_draft_lane_code_of_the_filter_matches_reference_binary(femu, name):
    """t1k_filter_lane.cuh (runtime-k seeds, bucket chaining, verdict; not yet in the product library) run sequentially on
    the CPU keeps exactly the pairs the unmodified fastq-extractor keeps."""
    z = np.load(os.path.join(FILTER_DIR, name + ".npz"))
    if name.startswith("recipe_"):
        import importlib.util
        spec = importlib.util.spec_from_file_location("make_golden_filter", os.path.join(G.GOLDEN, "make_golden_filter.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        recs = mod.big_ref()
    else:
        recs = G.parse_fasta_bytes(z["fasta"].tobytes())
    paired = bool(int(z["paired"]))
    r1 = [bytes(x) for x in z["reads1"]]
    r2 = [bytes(x) for x in z["reads2"]] if paired else None
    f = O.CandidateFilter(recs, r1, paired, float(z["similarity"]))          # for k / hitLenRequired (and as a second opinion)
    seqs = [r[2] for r in recs]
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in seqs], out=off[1:])
    E = femu.femu_create(len(seqs), b"".join(seqs), O._p(off), f.k, f.hit_len, float(z["similarity"]))
    assert E
    got = np.zeros(len(r1), dtype=np.uint8)
    for i in range(len(r1)):
        g = femu.femu_good_candidate(E, r1[i])
        assert g >= 0
        if not g and paired:
            g = femu.femu_good_candidate(E, r2[i])
            assert g >= 0
        got[i] = g
    femu.femu_destroy(E)
    assert np.array_equal(got, z["kept"]), np.flatnonzero(got != z["kept"])[:10]


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_device_filter_matches_reference_binary(name):
    """k_filter through the C ABI (t1k_filter_batch) keeps exactly the pairs the unmodified fastq-extractor keeps."""
    from t1k_b200.extractor import CandidateFilter
    z = np.load(os.path.join(FILTER_DIR, name + ".npz"))
    if name.startswith("recipe_"):
        import importlib.util
        spec = importlib.util.spec_from_file_location("make_golden_filter", os.path.join(G.GOLDEN, "make_golden_filter.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        recs = mod.big_ref()
    else:
        recs = G.parse_fasta_bytes(z["fasta"].tobytes())
    paired = bool(int(z["paired"]))
    r1 = [bytes(x) for x in z["reads1"]]
    r2 = [bytes(x) for x in z["reads2"]] if paired else None
    f = CandidateFilter(recs, r1, paired, float(z["similarity"]))
    orc = O.CandidateFilter(recs, r1, paired, float(z["similarity"]))
    assert (f.k, f.hit_len) == (orc.k, orc.hit_len)
    got = f.keep_pairs(r1, r2)
    assert np.array_equal(got, z["kept"]), np.flatnonzero(got != z["kept"])[:10]
    # per read against the oracle (both mates, incl. the reads of kept pairs the pair rule never looks at)
    reads = r1 + (r2 or [])
    g, st = f.IsGoodCandidate(reads, with_stats=True)
    want = np.asarray([orc.good(r) for r in reads], dtype=np.uint8)
    assert np.array_equal(g, want), np.flatnonzero(g != want)[:10]
    assert st["windows"] > 0 and st["kmer_length"] == f.k
