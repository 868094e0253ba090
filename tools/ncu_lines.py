#!/usr/bin/env python
"""Per-source-line summary of an ncu report: joins `ncu --page source --csv` (SASS rows, in address order) with
`nvdisasm -g` line info of the same cubin.  Usage: ncu_lines.py report.ncu-rep lib.so kernel_mangled_substr [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    td = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(td, cubin)], stdout=subprocess.PIPE, text=True).stdout.split("\n")
    # line table of the kernel's .text section
    start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":")][0]
    table = {}
    cur = ("?", 0)
    for l in sass[start + 1:]:
        if l.startswith(".text.") or l.startswith("//-----"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    base = None
    agg = {}
    tot_i = tot_s = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        if base is None:
            base = a
        key, _ = table.get(a - base, (("?", 0), ""))
        ins, thr, smp = int(r[iI] or 0), int(r[iT] or 0), int(r[iS] or 0)
        g = agg.setdefault(key, [0, 0, 0])
        g[0] += ins; g[1] += thr; g[2] += smp
        tot_i += ins; tot_s += smp
    src_cache = {}

    def src(f, n):
        if f not in src_cache:
            src_cache[f] = []
            for root, _, files in os.walk(os.path.dirname(os.path.dirname(os.path.abspath(so)))):
                if f in files:
                    src_cache[f] = open(os.path.join(root, f), errors="replace").read().split("\n")
                    break
        L = src_cache[f]
        return L[n - 1].strip()[:100] if 0 < n <= len(L) else ""

    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for key, (ins, thr, smp) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        print("%5.1f%% smp %5.1f%% inst  lanes %4.1f | %s:%d | %s" % (100.0 * smp / max(tot_s, 1), 100.0 * ins / max(tot_i, 1), thr / max(ins, 1), key[0], key[1],
                                                                     src(*key)))


if __name__ == "__main__":
    main()
