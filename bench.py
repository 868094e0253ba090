#!/usr/bin/env python
"""Benchmark of the genotyping hot path (align + EM): BASELINE.json metric "candidate reads/sec through align+EM".

A step = one pass of the hot path (de-duplicate read-ends -> k-mer seeded banded alignment -> fragment pairing ->
read-group coalescing -> equivalence classes -> SQUAREM EM) over one batch of synthetic 150 bp paired-end fragments
against the synthetic HLA-RNA-like allele reference (30,000 alleles x 1,100 bp).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl ours|reference]

One JSON line on stdout (rank 0).  `value` = fragments/s with the read-ends resident on the device (CUDA-event time of
the kernels of the step), `e2e` = the same metric through the C-ABI call `t1k_genotype` from HOST buffers (wall clock,
host<->device copies and the host-side model steps inside the timed region).
"""
from __future__ import annotations

import argparse
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

METRIC = "candidate reads/sec through align+EM (150bp PE, HLA ref)"
UNIT = "fragments/s"
WORKLOAD = "configs[1]: synthetic 150bp PE vs HLA-RNA-like ref (30000 alleles x 1100bp), -s 0.97"


def make_workload(n_pairs, seed, ref_records=None):
    from t1k_b200 import synth
    from t1k_b200.refset import RefSet
    recs = ref_records if ref_records is not None else synth.make_hla_rna_ref(seed=11)
    ref = RefSet(recs)
    kept = list(zip(ref.names, ref.comments, ref.seqs))
    r1, r2, _ = synth.simulate_pairs(kept, n_pairs, read_len=150, insert=(300, 450), err=0.002, alleles_per_gene=2, seed=seed, src_seed=7)
    return recs, ref, r1, r2


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


def measured_traffic(fragments_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE k_assign launch from the committed ncu capture
    (profiles/k_assign_traffic.json, written by tools/ncu_traffic.py); None unless the capture's launch size is the bench's."""
    p = os.path.join(ROOT, "profiles", "k_assign_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        t = json.load(f)
    if int(t.get("fragments_per_launch", -1)) != int(fragments_per_launch):
        return None, "profiles/k_assign_traffic.json is for %s fragments per launch" % t.get("fragments_per_launch")
    return int(t["dram_bytes_per_launch"]), t.get("source")


def reference_arm(args, rank):
    """The reference's own CPU implementation (oracle/_ref/genotyper, compiled unmodified from /root/reference) on the
    box's host cores, all threads, on a bounded sample of the same workload.  Rate = slope between two sample sizes so
    that reference loading / index build / FASTQ parsing (outside the metric) cancel."""
    if rank != 0:
        return None
    from t1k_b200 import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "genotyper")
    cores = os.cpu_count() or 1
    if not os.path.exists(exe):
        return {"impl": "reference", "unavailable": "oracle/_ref/genotyper was not built (reference checkout absent at build time)"}
    recs, ref, r1, r2 = make_workload(args.ref_pairs, seed=1234)
    small = max(50, args.ref_pairs // 5)
    td = tempfile.mkdtemp(prefix="t1kref_")
    try:
        fa = os.path.join(td, "ref.fa")
        synth.write_fasta(fa, recs)
        paths = {}
        for tag, n in (("small", small), ("full", args.ref_pairs)):
            p1, p2 = os.path.join(td, tag + "_1.fq"), os.path.join(td, tag + "_2.fq")
            synth.write_fastq(p1, r1[:n])
            synth.write_fastq(p2, r2[:n])
            paths[tag] = (p1, p2, n)

        def run(tag):
            p1, p2, n = paths[tag]
            t0 = time.perf_counter()
            subprocess.run([exe, "-f", fa, "-1", p1, "-2", p2, "-t", str(cores), "-s", "0.97", "-o", os.path.join(td, tag)],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            return time.perf_counter() - t0

        times = []
        for it in range(args.warmup + args.steps):
            ts, tf = run("small"), run("full")
            if it >= args.warmup:
                times.append(tf - ts)
        dt = float(np.mean(times))
        rate = (args.ref_pairs - small) / dt if dt > 0 else 0.0
    finally:
        shutil.rmtree(td, ignore_errors=True)
    sample = "%d vs %d fragments of the same synthetic workload, stock reference genotyper -t %d -s 0.97, rate from the wall-clock difference" % (
        args.ref_pairs, small, cores)
    return {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "fragments_per_step": args.ref_pairs - small},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def _claim_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on stdout.
    Everything else is sent to stderr; the returned file object is the real stdout for the JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("T1K_BENCH_PAIRS", 1000000)), help="fragments per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-pairs", type=int, default=1500, help="fragments of the CPU reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
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

    recs, ref, r1, r2 = make_workload(args.pairs, seed=100 + rank)          # weak scaling: every rank its own shard
    gt = Genotyper(ref, 0.97, False, device=local)
    h2d = int(r1.nbytes + r2.nbytes)

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
    dev_ms, launches, outs = 0.0, 0, []
    for _ in range(args.steps):
        out = step()
        outs.append(out)
        dev_ms += out["ms_align_kernel"] + out["ms_pair_kernel"] + out["ms_em_kernel"]
        launches += int(out["n_launches"])
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([wall, dev_ms / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_max, dev_max = float(t[0]), float(t[1])
    total_frag = args.pairs * world * args.steps

    out = outs[-1]
    # roofline of the dominant kernel (k_assign): algorithmic bytes per read-end = ceil(L/4) + 8 B per posting read +
    # 40 B per record kept, over the CUDA-event duration of its launches in the last timed step
    peak, peak_src = measured_peaks()
    # k_assign is launched once per chunk of fragments (T1K_CHUNK_FRAGMENTS = 2^18, a small first chunk to start the
    # pipeline): per-launch figures are quoted for a full-size launch = per-step x (chunk / fragments per step)
    chunk = int(os.environ.get("T1K_CHUNK_FRAGMENTS", 1 << 18))
    per_launch = min(args.pairs, chunk) / float(args.pairs)
    first = min(chunk, int(os.environ.get("T1K_FIRST_CHUNK", chunk)))
    n_chunks = 1 + max(0, -(-(args.pairs - first) // chunk)) if args.pairs > first else 1
    alg_bytes = (38 * out["n_unique_ends"] + 8 * out["n_postings"] + 40 * out["n_overlaps"]) * per_launch
    k_ms = out["ms_align_kernel"] * per_launch
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic, traffic_src = measured_traffic(min(args.pairs, chunk))
    d2h = int(out["n_assignments"] * 24 + out["n_unique_ends"] * 16)
    line = {
        "metric": METRIC, "value": total_frag / dev_max if dev_max > 0 else 0.0, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_max * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 (alignment scores) + f64 (EM)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "fragments_per_gpu_per_step": args.pairs, "unique_read_ends": int(out["n_unique_ends"]),
                   "overlap_records": int(out["n_overlaps"]), "assignments": int(out["n_assignments"]), "read_groups": int(out["n_groups"]),
                   "equivalence_classes": int(out["n_ec"]), "em_iterations": int(out["em_iterations"]),
                   "value_basis": "CUDA-event time of k_assign + k_pair + EM kernels per step, read-ends resident in HBM",
                   "l2_policy": "inputs larger than L2 (posting lists 264 MB + record store >> 126 MB); no flush needed",
                   "parallelism": "read-shard x%d" % world},
        "e2e": {"value": total_frag / wall_max, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": wall_max * 1e3 / args.steps,
                "phases_ms": {k: float(out[k]) for k in ("ms_dedup", "ms_align", "ms_pair", "ms_coalesce", "ms_em", "ms_align_kernel",
                                                            "ms_pair_kernel", "ms_em_kernel")}},
        "gpu_launches": launches,
        "roofline": {"kernel": "k_assign", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(alg_bytes),
                     "kernel_ms": k_ms, "launches_per_step": n_chunks,
                     "note": "k_assign is integer-ALU/latency bound (chaining + banded alignment per (read, allele)); HBM fraction is honest but not its roof"},
        "clocks": clocks,
    }
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            ref_line = reference_arm(argparse.Namespace(**{**vars(args), "steps": 1, "warmup": 0}), 0)
            line["cpu_baseline"] = ref_line.get("cpu_baseline") or {"value": None, "unavailable": ref_line.get("unavailable")}
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
