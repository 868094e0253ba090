#!/usr/bin/env python
"""One-screen summary of an .ncu-rep (raw page): duration, DRAM traffic, occupancy, issue rate, top stall reasons.
Usage: ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_alu.sum"]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h, u = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(h, r))
            print("== %s :: %s  grid %s block %s" % (rep, d["Kernel Name"][:60], d["Grid Size"], d["Block Size"]))
            for k in KEYS:
                for hk in h:
                    if hk == k or hk.endswith("." + k):
                        print("   %-70s %s %s" % (k, d[hk], u[h.index(hk)]))
                        break
            stalls = []
            for hk in h:
                if "average_warps_issue_stalled_" in hk and hk.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(d[hk].replace(",", "")), hk.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("   stalls (warps stalled per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:8]))


if __name__ == "__main__":
    main()
