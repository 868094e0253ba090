#!/usr/bin/env python
"""Writes profiles/k_assign_traffic.json from an ncu capture of ONE k_assign launch taken at the bench's launch size:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_assign -c 1 \\
      --csv --log-file gpurun_out/traffic.csv python bench.py --pairs 262144 --steps 1 --warmup 1 --no-cpu-baseline
Usage: ncu_traffic.py traffic.csv fragments_per_launch out.json"""
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
iN, iU, iV = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tot, dur = 0.0, None
for r in rows[1:]:
    v = float(r[iV].replace(",", ""))
    if r[iN].startswith("dram__bytes_"):
        tot += v * scale.get(r[iU], 1)
    if r[iN] == "gpu__time_duration.sum":
        dur = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU], 1.0)
json.dump({"kernel": "k_assign", "fragments_per_launch": int(sys.argv[2]), "dram_bytes_per_launch": int(tot), "ncu_duration_ms": dur,
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch, " + sys.argv[1].split("/")[-1]}, open(sys.argv[3], "w"), indent=1)
print(open(sys.argv[3]).read())
