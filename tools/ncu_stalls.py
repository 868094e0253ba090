#!/usr/bin/env python
"""Warp-stall samples of one reason per CUDA function / source line, from an .ncu-rep captured with --import-source on.
Usage: ncu_stalls.py report.ncu-rep [reason=no_inst] [top=30]"""
import csv, io, re, subprocess, sys
FUNC = re.compile(r"^(?!//|#|\}|namespace|struct|enum|typedef|constexpr|extern|using|template)\S.*\(.*[{,)]\s*(//.*)?$")
rep = sys.argv[1]
reason = "stall_" + (sys.argv[2] if len(sys.argv) > 2 else "no_inst")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; cur_file = "?"; disk = []
lines = []; funcs = {}; tot = {}
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        if r[0] == "File Path":
            disk = []; name = "?"
            try:
                for src_line in open(r[1], errors="replace").read().split("\n"):
                    if FUNC.match(src_line):
                        mm = re.search(r"([A-Za-z_0-9]+)\s*\(", src_line.replace("__launch_bounds__(", "launch_bounds "))
                        name = "%s:%s" % (cur_file, mm.group(1) if mm else "?")
                    disk.append(name)
            except OSError:
                pass
        continue
    if r[0] == "Line No":
        hdr = r; cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}; continue
    if hdr is None or not r[0].isdigit(): continue
    ln = int(r[0])
    vals = {}
    for h, i in cols.items():
        try: vals[h] = int(r[i] or 0)
        except (ValueError, IndexError): pass
    s = sum(vals.values())
    if s == 0: continue
    for h, v in vals.items(): tot[h] = tot.get(h, 0) + v
    v = vals.get(reason, 0)
    func = disk[ln - 1] if 0 < ln <= len(disk) else cur_file + ":?"
    f = funcs.setdefault(func, [0, 0]); f[0] += v; f[1] += s
    lines.append((v, s, cur_file, ln, r[1].strip()[:100]))
allsum = sum(tot.values()) or 1
print("all samples %d: " % allsum + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / allsum) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
rs = tot.get(reason, 0) or 1
print("-- %s by function (share of the reason | share of the function's own samples)" % reason[6:])
for k, (v, s) in sorted(funcs.items(), key=lambda kv: -kv[1][0])[:16]:
    print("%5.1f%% | %4.1f%% | %s" % (100.0 * v / rs, 100.0 * v / max(s, 1), k))
print("-- by line")
for v, s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% | %4.1f%% | %s:%d | %s" % (100.0 * v / rs, 100.0 * v / max(s, 1), f, ln, src))
