#!/usr/bin/env python
"""DRAM traffic per kernel group of ONE chunk of fragments, from an ncu launch list taken with
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file X.csv \
      python bench.py --config C --pairs <chunk> --steps 1 --warmup 0 --no-cpu-baseline
-> profiles/kernel_traffic.json (read by bench.py for roofline.traffic).  Usage: ncu_traffic.py X.csv config fragments_per_launch"""
import collections
import csv
import json
import os
import sys

GROUPS = {"k_assign": ("k_seed", "k_deferred", "k_defer_", "k_passes", "k_align"), "k_pair": ("k_pair",), "k_em": ("k_em_",)}


def main():
    path, config, frags = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    per = collections.defaultdict(float)
    ms = collections.defaultdict(float)
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        d = dict(zip(h, r))
        v = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "")
        for g, pats in GROUPS.items():
            if any(p in d["Kernel Name"] for p in pats):
                if d["Metric Name"].startswith("dram__bytes"):
                    per[g] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
                elif d["Metric Name"] == "gpu__time_duration.sum":
                    ms[g] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1e-6)
    out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_traffic.json")
    t = json.load(open(out_path)) if os.path.exists(out_path) else {}
    t["config%d" % config] = {"fragments_per_launch": frags, "dram_bytes_per_launch": {g: int(v) for g, v in per.items()},
                              "kernel_ms_under_ncu": {g: round(v, 3) for g, v in ms.items()},
                              "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of each kernel group, one chunk of %d fragments (%s)" % (frags, os.path.basename(path))}
    json.dump(t, open(out_path, "w"), indent=1)
    print(json.dumps(t["config%d" % config]))


if __name__ == "__main__":
    main()
