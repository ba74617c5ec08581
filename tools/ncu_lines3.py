#!/usr/bin/env python
"""Per-source-line totals of one kernel launch: ncu_lines3.py <rep> <kernel-regex> <launch-skip> [n]
(sums the SASS rows of ncu's source page under each CUDA line: executed warp instructions and stall samples)"""
import csv, subprocess, sys, collections
rep, kre, skip = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
cur = None; hdr = None
inst = collections.Counter(); samp = collections.Counter(); text = {}
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); continue
    if hdr and len(r) == len(hdr):
        try: ln = int(r[0])
        except ValueError: continue
        key = (cur, ln)
        if r[1].strip(): text[key] = r[1].strip()[:100]
        try: inst[key] += int(r[ie]); samp[key] += int(r[sm])
        except ValueError: pass
tot = sum(inst.values()) or 1; ts = sum(samp.values()) or 1
print("total warp-instructions", tot, "samples", ts)
for key, v in sorted(inst.items(), key=lambda kv: -samp[kv[0]])[:n]:
    print("%5.1f%% inst %5.1f%% samples  %-16s:%-5d %s" % (100.0 * v / tot, 100.0 * samp[key] / ts, key[0], key[1], text.get(key, "")))
