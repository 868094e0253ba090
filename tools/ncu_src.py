#!/usr/bin/env python
"""Per-CUDA-line and per-function summary of an .ncu-rep captured with --import-source on (compile with -lineinfo).
Parses `ncu --page source --print-source cuda,sass --csv`; functions are found from the source text embedded in the
report (nearest preceding line that looks like a function header).  Usage: ncu_src.py report.ncu-rep [top_lines]"""
import csv
import io
import re
import subprocess
import sys

FUNC = re.compile(r"^(?!//|#|\}|namespace|struct|enum|typedef|constexpr|extern|using|template)\S.*\(.*[{,)]\s*(//.*)?$")


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, func = "?", None, "?"
    disk = []
    lines, funcs = [], {}
    tot_s = tot_i = 0
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur_file = r[1].split("/")[-1]; func = "?"
            if r[0] == "File Path":
                disk = []                                    # line -> enclosing function, from the file on disk
                try:
                    name = "?"
                    for src_line in open(r[1], errors="replace").read().split("\n"):
                        if FUNC.match(src_line):
                            mm = re.search(r"([A-Za-z_0-9]+)\s*\(", src_line.replace("__launch_bounds__(", "launch_bounds "))
                            name = "%s:%s" % (cur_file, mm.group(1) if mm else "?")
                        disk.append(name)
                except OSError:
                    pass
            continue
        if r[0] == "Line No":
            hdr = r
            iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or not r[0].isdigit():
            continue
        ln = int(r[0])
        func = disk[ln - 1] if 0 < ln <= len(disk) else cur_file + ":?"
        try:
            s, i, t = int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0)
        except (ValueError, IndexError):
            continue
        if s == 0 and i == 0:
            continue
        lines.append((s, i, t, cur_file, int(r[0]), r[1].strip()[:110]))
        f = funcs.setdefault(func, [0, 0, 0])
        f[0] += s; f[1] += i; f[2] += t
        tot_s += s; tot_i += i
    print("total: %d warp instructions, %d samples" % (tot_i, tot_s))
    print("-- by function")
    for k, (s, i, t) in sorted(funcs.items(), key=lambda kv: -kv[1][0])[:25]:
        print("%5.1f%% smp %5.1f%% inst  lanes %4.1f | %s" % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), t / max(i, 1), k))
    print("-- by line")
    for s, i, t, f, n, src in sorted(lines, reverse=True)[:top]:
        print("%5.1f%% smp %5.1f%% inst  lanes %4.1f | %s:%d | %s" % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), t / max(i, 1), f, n, src))


if __name__ == "__main__":
    main()
