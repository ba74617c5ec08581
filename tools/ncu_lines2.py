#!/usr/bin/env python
"""Top CUDA source lines by executed instructions: ncu_lines2.py <rep> <kernel-regex> <launch-skip> [n]"""
import csv, subprocess, sys
rep, kre, skip = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
cur = None; hdr = None; res = []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
        try: ex = int(r[ie])
        except ValueError: continue
        res.append((ex, cur, r[0], r[1].strip()[:105], r[sm]))
tot = sum(x[0] for x in res)
print("total warp-instructions", tot)
for x in sorted(res, reverse=True)[:n]:
    print("%5.1f%% %-16s:%-5s %s  [samples %s]" % (100.0 * x[0] / tot, x[1], x[2], x[3], x[4]))
