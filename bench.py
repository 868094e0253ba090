#!/usr/bin/env python
"""Benchmark of the genotyping hot path (align + EM): BASELINE.json metric "candidate reads/sec through align+EM".

A step = one pass of the hot path (de-duplicate read-ends -> k-mer seeded banded alignment -> fragment pairing ->
read-group coalescing -> equivalence classes -> SQUAREM EM) over one batch of synthetic fragments:

  --config 2 (default)  150 bp PE vs the HLA-RNA-like reference (30,000 alleles x 1,100 bp), -s 0.97         SURVEY.md §8d cfg 2
  --config 3            150 bp PE vs the KIR-DNA-like reference (17 genes x 90 alleles x ~5 kb, N separators),
                        -s 0.9 --relaxIntronAlign                                                            cfg 3
  --config 4            100 bp SE vs the HLA-DNA-like reference (30,000 alleles x ~3.5 kb), -s 0.97           cfg 4

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--config C] [--impl ours|reference]

One JSON line on stdout (rank 0).
  value  whole-job fragments/s of the step with the read-ends staged: wall clock of the C-ABI call `t1k_genotype` (alignment
         kernels, pairing, D2H of the fragment rows, coalescing, multi-GPU exchange, equivalence classes, EM) minus the part
         of the host-side read de-duplication / staging the device had to wait for; max over ranks.
  e2e    the same call from HOST buffers, everything inside (wall clock, max over ranks).
Both include every host phase and collective, so the 1 -> N curve of either measures the system, not the kernels.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "fragments/s"
CONFIGS = {
    2: dict(metric="candidate reads/sec through align+EM (150bp PE, HLA ref)",
            workload="configs[1]: synthetic 150bp PE vs HLA-RNA-like ref (30000 alleles x 1100bp), -s 0.97",
            sim=0.97, relax=False, read_len=150, single_end=False, insert=(300, 450), err=0.002, pairs=1000000, ref_pairs=6000),
    3: dict(metric="candidate reads/sec through align+EM (150bp PE, KIR DNA ref, kir-wgs)",
            workload="configs[2]: synthetic 150bp PE vs KIR-DNA-like ref (17 genes x 90 alleles x ~5kb, N separators), -s 0.9 --relaxIntronAlign",
            sim=0.9, relax=True, read_len=150, single_end=False, insert=(300, 450), err=0.005, pairs=1000000, ref_pairs=20000),
    4: dict(metric="candidate reads/sec through align+EM (100bp SE, HLA DNA ref, hla-wgs)",
            workload="configs[3]: synthetic 100bp SE vs HLA-DNA-like ref (30000 alleles x ~3.5kb, N separators), -s 0.97",
            sim=0.97, relax=False, read_len=100, single_end=True, insert=(100, 100), err=0.002, pairs=2000000, ref_pairs=6000),
}
METRIC = CONFIGS[2]["metric"]
WORKLOAD = CONFIGS[2]["workload"]

_REF_CACHE = {}


def make_reference(config):
    """the synthetic allele reference of a config -> (records, RefSet)"""
    if config not in _REF_CACHE:
        from t1k_b200 import synth
        from t1k_b200.refset import RefSet
        if config == 2:
            recs = synth.make_hla_rna_ref(seed=11)
        elif config == 3:
            recs = synth.make_dna_ref(n_genes=17, alleles_per_gene=90, n_exons=9, exon_mean=300, pad=150, n_sites=120, min_sub=2, max_sub=10,
                                      seed=23, prefix="KIR", family_div=0.05)
        elif config == 4:
            recs = []
            for g, (name, cnt) in enumerate((("HLA-A", 8000), ("HLA-B", 9000), ("HLA-C", 8000), ("HLA-DRB1", 3000), ("HLA-DQB1", 1500), ("HLA-DPB1", 500))):
                part = synth.make_dna_ref(n_genes=1, alleles_per_gene=cnt, n_exons=8, exon_mean=240, pad=100, n_sites=300, min_sub=3, max_sub=14,
                                          seed=41 + g, prefix="X")
                recs += [("%s*%02d:%02d:01" % (name, i // 99 + 1, i % 99 + 1), c, s) for i, (_, c, s) in enumerate(part)]
        else:
            raise SystemExit("unknown --config")
        _REF_CACHE[config] = (recs, RefSet(recs))
    return _REF_CACHE[config]


def make_workload(n_pairs, seed, ref_records=None, config=2, start_frac=(0.0, 1.0), insert=None):
    from t1k_b200 import synth
    from t1k_b200.refset import RefSet
    cfg = CONFIGS[config]
    if ref_records is not None:
        recs, ref = ref_records, RefSet(ref_records)
    else:
        recs, ref = make_reference(config)
    kept = list(zip(ref.names, ref.comments, ref.seqs))
    r1, r2, _ = synth.simulate_pairs(kept, n_pairs, read_len=cfg["read_len"], insert=insert or cfg["insert"], err=cfg["err"], alleles_per_gene=2,
                                     seed=seed, src_seed=7, single_end=cfg["single_end"], start_frac=start_frac)
    return recs, ref, r1, r2


def unique_ends(r1, r2):
    s = set(r.tobytes() for r in r1)
    if r2 is not None:
        s |= set(r.tobytes() for r in r2)
    return len(s)


def matched_sample(config, n, target, seed=1234):
    """n fragments of the config's generator with fragment starts / inserts restricted to a window chosen (bisection) so that
    the sample's unique read-ends per fragment equals `target`, the rate of the deep set the GPU arm runs: both arms
    de-duplicate read-ends and alignment cost is per unique read-end, so a shallow uniform sample (nearly every read-end
    unique) would overstate the CPU's cost per fragment."""
    cfg = CONFIGS[config]
    lo, hi = 0.0005, 1.0
    best = None
    for _ in range(12):
        w = (lo * hi) ** 0.5
        ins0 = cfg["insert"][0]
        ins = (ins0, ins0 + max(0, min(cfg["insert"][1] - ins0, int(round((cfg["insert"][1] - ins0) * min(1.0, 4 * w))))))
        recs, ref, r1, r2 = make_workload(n, seed, config=config, start_frac=(0.3, min(1.0, 0.3 + w)), insert=ins)
        rate = unique_ends(r1, r2) / float(n)
        if best is None or abs(rate - target) < abs(best[0] - target):
            best = (rate, w, recs, ref, r1, r2)
        if abs(rate - target) <= 0.01 * target:
            break
        if rate < target:
            lo = w
        else:
            hi = w
    return best


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.p = None
        self.idx = gpu_index

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(config, fragments_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum per kernel group for ONE chunk of `fragments_per_launch` fragments, from the
    committed ncu capture (profiles/kernel_traffic.json, written by tools/ncu_traffic.py); {} unless the capture matches."""
    p = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if not os.path.exists(p):
        return {}, None
    with open(p) as f:
        t = json.load(f)
    key = "config%d" % config
    if key not in t or int(t[key].get("fragments_per_launch", -1)) != int(fragments_per_launch):
        return {}, "profiles/kernel_traffic.json has no capture for config %d at %d fragments per launch" % (config, fragments_per_launch)
    return t[key]["dram_bytes_per_launch"], t[key].get("source")


def reference_arm(args, rank, target_rate=None):
    """The reference's own CPU implementation (oracle/_ref/genotyper, compiled unmodified from /root/reference) on the
    box's host cores, all threads, on a bounded sample of the same workload whose duplicate rate matches the deep set of
    the GPU arm.  Rate = slope between two sample sizes so that reference loading / index build / FASTQ parsing (outside the
    metric) cancel."""
    if rank != 0:
        return None
    from t1k_b200 import synth
    cfg = CONFIGS[args.config]
    exe = os.path.join(ROOT, "oracle", "_ref", "genotyper")
    cores = os.cpu_count() or 1
    if not os.path.exists(exe):
        return {"impl": "reference", "unavailable": "oracle/_ref/genotyper was not built (reference checkout absent at build time)"}
    if target_rate is None:
        target_rate = deep_set_rate(args.config, args.pairs)
    # sample size: the config's, shrunk when many steps are asked for so that the whole run stays within a few minutes
    n = args.ref_pairs or max(1000, min(cfg["ref_pairs"], cfg["ref_pairs"] * 8 // max(1, args.warmup + args.steps)))
    rate_u, w, recs, ref, r1, r2 = matched_sample(args.config, n, target_rate)
    small = max(50, n // 5)
    mates = 1 if cfg["single_end"] else 2
    u_full, u_small = unique_ends(r1, r2), unique_ends(r1[:small], r2[:small] if r2 is not None else None)
    td = tempfile.mkdtemp(prefix="t1kref_")
    try:
        fa = os.path.join(td, "ref.fa")
        synth.write_fasta(fa, recs)
        paths = {}
        for tag, k in (("small", small), ("full", n)):
            p1, p2 = os.path.join(td, tag + "_1.fq"), os.path.join(td, tag + "_2.fq")
            synth.write_fastq(p1, r1[:k])
            if r2 is not None:
                synth.write_fastq(p2, r2[:k])
            paths[tag] = (p1, p2, k)

        def run(tag):
            p1, p2, k = paths[tag]
            cmd = [exe, "-f", fa, "-t", str(cores), "-s", str(cfg["sim"]), "-o", os.path.join(td, tag)]
            cmd += ["-u", p1] if r2 is None else ["-1", p1, "-2", p2]
            if cfg["relax"]:
                cmd.append("--relaxIntronAlign")
            t0 = time.perf_counter()
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            return time.perf_counter() - t0

        times = []
        for it in range(args.warmup + args.steps):
            ts, tf = run("small"), run("full")
            if it >= args.warmup:
                times.append(tf - ts)
        dt = float(np.mean(times))
        rate = (n - small) / dt if dt > 0 else 0.0
    finally:
        shutil.rmtree(td, ignore_errors=True)
    sample = ("%d vs %d fragments of the config's generator with starts / inserts restricted to a window (fraction %.4f) so that the sample has "
              "%.3f unique read-ends per fragment (GPU arm's deep set: %.3f); stock reference genotyper -t %d -s %s%s, rate from the wall-clock "
              "difference" % (n, small, w, rate_u, target_rate, cores, cfg["sim"], " --relaxIntronAlign" if cfg["relax"] else ""))
    return {"impl": "reference", "metric": cfg["metric"], "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "fragments_per_step": n - small, "unique_read_ends_per_fragment": rate_u},
            "read_ends_per_s": rate * mates, "unique_read_ends_per_s": (u_full - u_small) / dt if dt > 0 else 0.0,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample,
                             "unique_read_ends_per_fragment": rate_u, "unique_read_ends_per_s": (u_full - u_small) / dt if dt > 0 else 0.0},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def deep_set_rate(config, pairs):
    """unique read-ends per fragment of the GPU arm's workload (rank 0's shard)"""
    _, _, r1, r2 = make_workload(pairs, seed=100, config=config)
    v1 = np.ascontiguousarray(r1).view(np.dtype((np.void, r1.shape[1])))
    allv = v1 if r2 is None else np.concatenate([v1, np.ascontiguousarray(r2).view(np.dtype((np.void, r2.shape[1])))])
    return len(np.unique(allv)) / float(pairs)


def filter_stage(args, real_stdout):
    """the candidate filter of fastq-extractor (k_filter through t1k_filter_batch) on a WGS-like read set:
    150 bp pairs, 2 % drawn from the HLA-RNA-like reference (candidates), 98 % random sequence (the bulk of a genome).  The
    CPU baseline is the unmodified fastq-extractor binary (oracle/_ref/fastq-extractor) on a bounded sample."""
    import torch
    from t1k_b200 import synth
    from t1k_b200.extractor import CandidateFilter
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    n = args.pairs or 4000000
    recs, ref = make_reference(2)
    rng = np.random.default_rng(5)
    n_cand = n // 50
    _, _, c1, c2 = make_workload(n_cand, seed=77, config=2)
    alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    r1 = alpha[rng.integers(0, 4, size=(n, 150), dtype=np.uint8)]
    r2 = alpha[rng.integers(0, 4, size=(n, 150), dtype=np.uint8)]
    where = rng.choice(n, size=n_cand, replace=False)
    r1[where] = c1
    r2[where] = c2
    f = CandidateFilter(recs, [r.tobytes() for r in r1[:1000]], True, 0.8)
    both = np.concatenate([r1, r2])

    def step():
        return f.IsGoodCandidate(both, with_stats=True)

    for _ in range(args.warmup):
        good, st = step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ms_k = 0.0
    for _ in range(args.steps):
        good, st = step()
        ms_k += st["ms_kernel"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    kept = int((good[:n] | good[n:]).sum())
    peak, peak_src = measured_peaks()
    # algorithmic bytes per read: its packed bases + two 8-byte table entries per k-mer window of both strands (+ 16 B per entry swept)
    alg = (both.shape[0] * ((150 + 3) // 4) + 16 * st["windows"] + 16 * st["entries"])
    ms = ms_k / args.steps
    line = {"metric": "reads/sec through the fastq-extractor candidate filter (150bp PE, HLA ref, WGS-like mix)", "value": 2 * n * args.steps / (ms_k * 1e-3),
            "unit": "reads/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 / u32 (2-bit bases, k-mer codes)", "data": "synthetic",
            "config": {"workload": "SURVEY 8f N1: %d read pairs x 150 bp, 2 %% from the HLA-RNA-like reference, 98 %% random; k = %d, hitLenRequired = %d" % (n, f.k, f.hit_len),
                       "kept_pairs": kept, "reads_chained": int(st["chained"]), "value_basis": "CUDA-event time of k_filter, reads resident in HBM"},
            "e2e": {"value": 2 * n * args.steps / wall, "unit": "reads/s", "h2d_bytes_per_step": int(both.nbytes), "d2h_bytes_per_step": int(both.shape[0])},
            "gpu_launches": 2 * args.steps,
            "roofline": {"kernel": "k_filter", "bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg),
                         "note": "bytes = packed read + 16 B per k-mer window looked up (two 8-byte entries of the 4^k table, random access) + 16 B per index entry swept"}}
    exe = os.path.join(ROOT, "oracle", "_ref", "fastq-extractor")
    if not args.no_cpu_baseline and os.path.exists(exe):
        m = min(n, 200000)
        td = tempfile.mkdtemp(prefix="t1kfilt_")
        try:
            fa = os.path.join(td, "ref.fa")
            synth.write_fasta(fa, recs)
            synth.write_fastq(os.path.join(td, "a_1.fq"), r1[:m])
            synth.write_fastq(os.path.join(td, "a_2.fq"), r2[:m])
            cores = os.cpu_count() or 1
            t0 = time.perf_counter()
            subprocess.run([exe, "-f", fa, "-1", os.path.join(td, "a_1.fq"), "-2", os.path.join(td, "a_2.fq"), "-t", str(cores), "-o", os.path.join(td, "out")],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 2 * m / dt, "unit": "reads/s", "cores": cores, "kind": "reference",
                                    "sample": "first %d pairs of the same set through the stock fastq-extractor -t %d (wall clock incl. index build and FASTQ parsing)" % (m, cores)}
        finally:
            shutil.rmtree(td, ignore_errors=True)
    print(json.dumps(line), file=real_stdout, flush=True)


def dropin_stage(args, real_stdout):
    """The drop-in `genotyper` binary (integration/_build/genotyper_b200: the reference's driver with the hot path forwarded to the
    C ABI) on FASTQ files of the config's workload: wall clock of its hot path (between the reference's own log points
    "Start read assignment" and the EM) and of the whole process incl. FASTQ parsing and the reference's allele selection."""
    import re
    from t1k_b200 import synth
    exe = os.path.join(ROOT, "integration", "_build", "genotyper_b200")
    if not os.path.exists(exe):
        print(json.dumps({"stage": "dropin", "unavailable": "integration/_build/genotyper_b200 was not built (needs the T1K checkout at build time)"}), file=real_stdout, flush=True)
        return
    cfg = CONFIGS[args.config]
    n = args.pairs or cfg["pairs"]
    recs, ref, r1, r2 = make_workload(n, seed=100, config=args.config)
    td = tempfile.mkdtemp(prefix="t1kdropin_")
    try:
        fa = os.path.join(td, "ref.fa")
        synth.write_fasta(fa, recs)
        p1, p2 = os.path.join(td, "r_1.fq"), os.path.join(td, "r_2.fq")
        synth.write_fastq(p1, r1)
        if r2 is not None:
            synth.write_fastq(p2, r2)
        cmd = [exe, "-f", fa, "-s", str(cfg["sim"]), "-o", os.path.join(td, "out"), "-t", str(os.cpu_count() or 1)] + (["-u", p1] if r2 is None else ["-1", p1, "-2", p2])
        if cfg["relax"]:
            cmd.append("--relaxIntronAlign")
        runs = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = subprocess.run(cmd, env=dict(os.environ, T1K_TIMING="1"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
            wall = time.perf_counter() - t0
            m = re.search(r"hot path \+ FinalizeReadAssignments: (\d+) ms wall \((\w+) path\)", res.stderr)
            if it >= args.warmup:
                runs.append((wall, float(m.group(1)) / 1e3 if m else None, m.group(2) if m else None))
    finally:
        shutil.rmtree(td, ignore_errors=True)
    hot = float(np.mean([r[1] for r in runs if r[1] is not None])) if runs and runs[0][1] is not None else None
    wall = float(np.mean([r[0] for r in runs]))
    print(json.dumps({"stage": "dropin", "metric": cfg["metric"], "unit": UNIT, "value": n / hot if hot else None, "fragments": n, "path": runs[0][2],
                      "hot_path_s": hot, "process_wall_s": wall, "process_fragments_per_s": n / wall,
                      "config": {"workload": cfg["workload"], "what": "drop-in genotyper binary on FASTQ files: hot path = reads in RAM -> abundances set (t1k_genotype + the reference's "
                                 "FinalizeReadAssignments); process wall adds reference loading, FASTQ parsing, allele selection and the writers"}}), file=real_stdout, flush=True)


def alninfo_stage(args, real_stdout):
    """the analyzer's edit-string pass (k_align_info through t1k_align_info_batch) over the AssignRead records of
    config-2 read-ends, and §8d's band-DP roofline: the same items with the certified-diagonal shortcut off, band cells/s and
    DPX operations/s (two max(a+b,c) + one three-way max per cell) against the device's measured DPX issue rate."""
    import ctypes as C
    import torch
    from t1k_b200 import _lib as L
    from t1k_b200.genotyper import SeqSet
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    n = args.pairs or 600
    cfg = CONFIGS[2]
    recs, ref, r1, r2 = make_workload(n, seed=100, config=2)
    reads = sorted(set(r.tobytes() for r in r1) | set(r.tobytes() for r in r2))
    ss = SeqSet(ref, cfg["sim"], cfg["relax"])
    a = ss.AssignRead(reads, np.zeros(len(reads), dtype=np.int32))
    row, ret, rec = a.fetch()
    idx = np.repeat(np.arange(len(reads), dtype=np.uint32), np.diff(row).astype(np.int64))
    del a
    gops = C.c_double(0)
    L.check(L.lib().t1k_dpx_peak(-1, C.byref(gops)))

    def run(force_dp):
        ms, wall, st = 0.0, 0.0, None
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            _, st = ss.AddOverlapAlignmentInfo(reads, idx, rec, force_dp=force_dp, with_stats=True, raw=True)
            if it >= args.warmup:
                ms += st["ms_kernel"]
                wall += time.perf_counter() - t0
        return ms / args.steps, wall / args.steps, st

    ms, wall, st = run(False)
    ms_dp, _, st_dp = run(True)
    peak, peak_src = measured_peaks()
    out_bytes = int((rec["seqEnd"] - rec["seqStart"] + 1 + rec["readEnd"] - rec["readStart"] + 1 + 2 + 15).astype(np.int64).sum() // 16 * 16)
    alg = 44 * len(rec) + int((rec["seqEnd"] - rec["seqStart"] + 1 + 1).astype(np.int64).sum())     # item + the string (one byte per column + -1)
    cells = float(st_dp["dp_cells"])
    line = {"metric": "overlaps/sec through the analyzer's edit-string pass (AddOverlapAlignmentInfo, 150bp reads, HLA ref)", "value": len(rec) / (ms * 1e-3),
            "unit": "overlaps/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32 (affine-gap scores, DPX)", "data": "synthetic",
            "config": {"workload": "SURVEY 8f N3: every AssignRead record (weight 0) of %d unique read-ends of configs[1] (%d overlaps)" % (len(reads), len(rec)),
                       "certified_diagonal": int(st["n_diagonal"]), "band_dp": int(st["n_dp"]), "value_basis": "CUDA-event time of k_align_info, items resident in HBM"},
            "e2e": {"value": len(rec) / wall, "unit": "overlaps/s", "h2d_bytes_per_step": int(44 * len(rec) + sum(len(r) for r in reads)), "d2h_bytes_per_step": out_bytes},
            "gpu_launches": 2 * args.steps,
            "roofline": {"kernel": "k_align_info", "bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                         "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg),
                         "note": "bytes = 44 B per item + one byte per alignment column; the allele / read windows are L2-resident"},
            "dp_roofline": {"kernel": "k_align_info with the band DP forced for every item (dp_align_eq / dp_align + traceback)", "bound": "dpx",
                            "band_cells_per_s": cells / (ms_dp * 1e-3), "achieved": 3 * cells / (ms_dp * 1e-3) / 1e9, "peak": gops.value, "unit": "G DPX ops/s",
                            "frac": 3 * cells / (ms_dp * 1e-3) / 1e9 / gops.value if gops.value else None, "ms": ms_dp, "band_dps": int(st_dp["n_dp"]),
                            "peak_source": "t1k_dpx_peak: 8 independent VIADDMNMX chains per thread, 8 blocks of 256 threads per SM, CUDA events",
                            "note": "3 DPX operations per band cell (e, f: max(a+b,c); m: three-way max); a cell also costs ~12 integer / logic instructions "
                                    "(direction nibble, base compare, band bookkeeping), so the DPX pipe cannot be the binding limit"}}
    if not args.no_cpu_baseline:
        import oracle_py as O        # checker only: the CPU restatement of GlobalAlignment as the 1-core baseline
        rc = np.zeros(256, dtype=np.uint8)
        for x, y in zip(b"ACGTN", b"TGCAN"):
            rc[x] = y
        m = min(len(rec), 20000)
        pairs = []
        for k in range(m):
            rd = np.frombuffer(reads[int(idx[k])], dtype=np.uint8)
            if rec["strand"][k] == -1:
                rd = rc[rd[::-1]]
            pairs.append((bytes(ref.seqs[int(rec["seqIdx"][k])][int(rec["seqStart"][k]):int(rec["seqEnd"][k]) + 1]),
                          rd[int(rec["readStart"][k]):int(rec["readEnd"][k]) + 1].tobytes()))
        t0 = time.perf_counter()
        for t, q in pairs:
            O.global_alignment(t, q)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": m / dt, "unit": "overlaps/s", "cores": 1, "kind": "port",
                                "sample": "the first %d items through the oracle's GlobalAlignment (full three-matrix DP as AlignAlgo.hpp:215-421), one core" % m}
    print(json.dumps(line), file=real_stdout, flush=True)


def _claim_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on stdout.
    Everything else is sent to stderr; the returned file object is the real stdout for the JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def output_digest(out):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(out["equivalent_class"]).tobytes())
    h.update(np.ascontiguousarray(out["missing_coverage"]).tobytes())
    h.update(np.round(np.asarray(out["abundance"]), 6).tobytes())
    return h.hexdigest()[:16]


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="fragments per GPU per step (default: the config's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-pairs", type=int, default=0, help="fragments of the CPU reference sample (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the multi-GPU check against a one-GPU run of the union")
    ap.add_argument("--stage", default="genotype", choices=["genotype", "filter", "dropin", "alninfo"],
                    help="filter: SURVEY 8f N1, the extractor's candidate filter; dropin: the drop-in genotyper binary on FASTQ files; "
                         "alninfo: SURVEY 8f N3, the analyzer's edit strings + the band-DP / DPX roofline")
    args = ap.parse_args()
    if args.stage == "filter":
        filter_stage(args, real_stdout)
        return
    if args.stage == "dropin":
        dropin_stage(args, real_stdout)
        return
    if args.stage == "alninfo":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        alninfo_stage(args, real_stdout)
        return
    cfg = CONFIGS[args.config]
    if not args.pairs:
        args.pairs = int(os.environ.get("T1K_BENCH_PAIRS", cfg["pairs"]))
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        line = reference_arm(args, rank)
        if line is not None:
            print(json.dumps(line), file=real_stdout, flush=True)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from t1k_b200.genotyper import Genotyper
    comm = None
    if world > 1:
        from t1k_b200 import dist_em
        comm = dist_em.Comm.from_torch_distributed(local)       # NCCL communicator of the C library (id via torch.distributed)

    recs, ref, r1, r2 = make_workload(args.pairs, seed=100 + rank, config=args.config)          # weak scaling: every rank its own shard
    gt = Genotyper(ref, cfg["sim"], cfg["relax"], device=local)
    h2d = int(r1.nbytes + (r2.nbytes if r2 is not None else 0))
    mates = 1 if r2 is None else 2

    def step():
        return dist_em.genotype_sharded(gt, r1, r2, comm) if world > 1 else gt.Genotype(r1, r2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    wait_ms, launches, outs = 0.0, 0, []
    for _ in range(args.steps):
        out = step()
        outs.append(out)
        wait_ms += out["ms_prep_wait"]
        launches += int(out["n_launches"])
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([wall, wall - wait_ms / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_max, staged_max = float(t[0]), float(t[1])
    total_frag = args.pairs * world * args.steps
    uniq = torch.tensor([float(outs[-1]["n_unique_ends"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(uniq, op=dist.ReduceOp.SUM)
    uniq_all = float(uniq[0])

    out = outs[-1]
    digests = sorted(set(output_digest(o) for o in outs))
    verified = {"steps_identical": len(digests) == 1, "digest": digests[0]}
    # read-sharded run: the per-allele outputs must equal a one-GPU run over the union of the shards (outside the timed region)
    if world > 1 and not args.no_verify:
        if rank == 0:
            shards = [make_workload(args.pairs, seed=100 + r, config=args.config) for r in range(world)]
            u1 = np.concatenate([s[2] for s in shards])
            u2 = None if r2 is None else np.concatenate([s[3] for s in shards])
            one = gt.Genotype(u1, u2)
            scale = float(np.abs(one["abundance"]).max()) or 1.0
            verified.update({
                "vs_one_gpu_union": True,
                "equivalent_class_equal": bool(np.array_equal(one["equivalent_class"], out["equivalent_class"])),
                "missing_coverage_equal": bool(np.array_equal(one["missing_coverage"], out["missing_coverage"])),
                "groups_ecs_equal": bool(one["n_groups"] == out["n_groups"] and one["n_ec"] == out["n_ec"]),
                "abundance_max_abs_diff_over_max": float(np.abs(one["abundance"] - out["abundance"]).max() / scale),
                "abundance_within_1e-5": bool(np.abs(one["abundance"] - out["abundance"]).max() <= 1e-5 * scale),
                "em_iterations": [int(one["em_iterations"]), int(out["em_iterations"])]})
        dist.barrier()

    # ---- rooflines (three stages with different roofs, never blended).  Algorithmic bytes per launch = per
    # chunk of fragments (T1K_CHUNK_FRAGMENTS = 2^18): per-step counters x (chunk / fragments per step).
    peak, peak_src = measured_peaks()
    chunk = int(os.environ.get("T1K_CHUNK_FRAGMENTS", 1 << 18))
    per_launch = min(args.pairs, chunk) / float(args.pairs)
    n_chunks = -(-args.pairs // chunk)
    traffic, traffic_src = measured_traffic(args.config, min(args.pairs, chunk))
    read_bytes = (cfg["read_len"] + 3) // 4

    def roof(name, alg_bytes, ms, launches_per_step, note, key):
        achieved = alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        return {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get(key), "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": ms, "launches_per_step": launches_per_step,
                "note": note}

    k_assign = roof("AssignRead kernels (k_seed + k_deferred + k_passes + k_align, one chunk of fragments)",
                    (read_bytes * out["n_unique_ends"] + 8 * out["n_postings"] + 40 * out["n_overlaps"]) * per_launch, out["ms_align_kernel"] * per_launch, n_chunks,
                    "bytes = ceil(L/4) per read-end + 8 B per posting the reference's index would visit + 40 B per record kept (SURVEY §8d B_seed); "
                    "the tile index moves far fewer bytes than that; the kernels are integer-ALU / latency bound, not HBM bound", "k_assign")
    k_pair = roof("k_pair", (32 * out["n_pair_records"] + 24 * out["n_assignments"]) * per_launch, out["ms_pair_kernel"] * per_launch, n_chunks,
                  "bytes = 32 B per overlap record of both mates + 24 B per fragment-row entry", "k_pair")
    em_bytes = (4 * out["em_nnz"] + 12 * out["n_groups"] + 28 * out["n_ec"]) * max(1, out["em_updates"])
    k_em = roof("EM kernels (all EMupdates + SQUAREM vector steps of the step)", em_bytes, out["ms_em_kernel"], 1,
                "bytes = (4 nnz + 12 G + 28 E) per EMupdate (SURVEY §8d B_em) x EMupdates; the matrix fits L2, latency bound", "k_em")
    roofline = dict(k_assign)
    roofline.update({"traffic_source": traffic_src, "peak_source": peak_src, "kernels": [k_assign, k_pair, k_em]})

    d2h = int(out["n_assignments"] * 24 + args.pairs * 28)
    value = total_frag / staged_max if staged_max > 0 else 0.0
    e2e = total_frag / wall_max
    line = {
        "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": staged_max * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 (alignment scores) + f64 (EM)", "data": "synthetic",
        "config": {"workload": cfg["workload"], "fragments_per_gpu_per_step": args.pairs, "unique_read_ends": int(out["n_unique_ends"]),
                   "unique_read_ends_per_fragment": uniq_all / float(args.pairs * world),
                   "overlap_records": int(out["n_overlaps"]), "assignments": int(out["n_assignments"]), "read_groups": int(out["n_groups"]),
                   "equivalence_classes": int(out["n_ec"]), "em_iterations": int(out["em_iterations"]),
                   "value_basis": "wall clock of t1k_genotype per step (kernels, D2H of fragment rows, coalescing, multi-GPU exchange, equivalence classes, EM) "
                                  "minus the read de-duplication / staging time the device waited for; max over ranks",
                   "l2_policy": "inputs larger than L2 (candidate pool + record store >> 126 MB per chunk); no flush needed",
                   "parallelism": "read-shard x%d" % world},
        "read_ends_per_s": value * mates, "unique_read_ends_per_s": uniq_all * args.steps / staged_max if staged_max > 0 else 0.0,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": wall_max * 1e3 / args.steps, "read_ends_per_s": e2e * mates, "unique_read_ends_per_s": uniq_all * args.steps / wall_max,
                "phases_ms": {k: float(out[k]) for k in ("ms_dedup", "ms_prep_wait", "ms_align", "ms_pair", "ms_coalesce", "ms_exchange", "ms_em",
                                                            "ms_align_kernel", "ms_pair_kernel", "ms_em_kernel")}},
        "gpu_launches": launches,
        "roofline": roofline,
        "verified": verified,
        "clocks": clocks,
    }
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            ref_line = reference_arm(argparse.Namespace(**{**vars(args), "steps": 1, "warmup": 0}), 0, target_rate=uniq_all / float(args.pairs * world))
            line["cpu_baseline"] = ref_line.get("cpu_baseline") or {"value": None, "unavailable": ref_line.get("unavailable")}
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
