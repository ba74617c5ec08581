#!/usr/bin/env python3
"""Join an ncu SASS source page (ncu -i rep --page source --csv --print-source sass) with
nvdisasm -g line annotations of the same kernel and aggregate samples / executed instructions
per source line.  usage: ncu_lines.py sass.csv disasm.txt [top]"""
import csv, re, sys, collections
sass_csv, dis_txt = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(sass_csv)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
ins = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
base = int(ins[0][0], 16)
lines = {}
cur = None
for l in open(dis_txt):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines[int(m.group(1), 16)] = (cur, m.group(2).strip())
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot_s = tot_i = 0
miss = 0
for r in ins:
    off = int(r[0], 16) - base
    loc = lines.get(off, (None, ""))[0]
    if loc is None: miss += 1
    s = int(r[col["# Samples"]] or 0); n = int(r[col["Instructions Executed"]] or 0)
    lsb = int(r[col["stall_long_sb"]] or 0)
    a = agg[loc]; a[0] += s; a[1] += n; a[2] += lsb; a[3] += 1
    tot_s += s; tot_i += n
print(f"instructions {len(ins)}  unmatched {miss}  samples {tot_s}  warp-inst executed {tot_i}")
print(f"{'file:line':32s} {'samples%':>8s} {'inst%':>7s} {'long_sb%':>8s} {'sass':>5s}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    name = f"{loc[0]}:{loc[1]}" if loc else "?"
    print(f"{name:32s} {100*a[0]/tot_s:8.2f} {100*a[1]/tot_i:7.2f} {100*a[2]/max(tot_s,1):8.2f} {a[3]:5d}")
