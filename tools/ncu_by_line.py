#!/usr/bin/env python
"""Join an `ncu --page source --csv` dump (SASS view) with `nvdisasm -g -c` line info of the same
kernel and aggregate samples / executed instructions / shared wavefronts per source line and opcode.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all binary ; nvdisasm -g -c x.cubin > sass.txt
    ncu_by_line.py src.csv sass.txt <kernel-substring> [top]
"""
import collections
import csv
import re
import sys

src_csv, sass_txt, pat = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# --- nvdisasm: ordered list of (line tag, opcode) of the kernel
on = False
cur = "?"
lines = []
for line in open(sass_txt):
    m = re.match(r"\s*\.text\.(\S+):", line) or re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        on = pat in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        cur = m.group(1).split("/")[-1] + ":" + m.group(2)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        lines.append((cur, m.group(2)))

rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
print(f"sass instrs: nvdisasm {len(lines)}  ncu {len(data)}")
n = min(len(lines), len(data))
by_line = collections.defaultdict(lambda: [0, 0, 0, 0.0])  # samples, inst, wavefronts
by_op = collections.defaultdict(lambda: [0, 0, 0])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_line_stall = collections.defaultdict(lambda: collections.Counter())
tot_s = tot_i = tot_w = 0
for k in range(n):
    tag, txt = lines[k]
    r = data[k]
    s = int(r[ix["# Samples"]] or 0)
    i = int(r[ix["Instructions Executed"]] or 0)
    w = int(r[ix["L1 Wavefronts Shared"]] or 0)
    op = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip()).split()[0]
    by_line[tag][0] += s; by_line[tag][1] += i; by_line[tag][2] += w
    by_op[op][0] += s; by_op[op][1] += i; by_op[op][2] += w
    for c in stall_cols:
        v = int(r[ix[c]] or 0)
        if v:
            by_line_stall[tag][c[6:]] += v
    tot_s += s; tot_i += i; tot_w += w
print(f"total samples {tot_s}  inst {tot_i}  smem wavefronts {tot_w}")
print("\n== by source line (samples%, inst%, wavefronts%, top stalls)")
for tag, (s, i, w, _) in sorted(by_line.items(), key=lambda t: -t[1][0])[:top]:
    st = ", ".join(f"{k}:{v}" for k, v in by_line_stall[tag].most_common(3))
    print(f"{tag:28s} {100*s/tot_s:6.2f}% {100*i/tot_i:6.2f}% {100*w/max(tot_w,1):6.2f}%   {st}")
print("\n== by opcode (samples%, inst%, wavefronts%)")
for op, (s, i, w) in sorted(by_op.items(), key=lambda t: -t[1][1])[:30]:
    print(f"{op:22s} {100*s/tot_s:6.2f}% {100*i/tot_i:6.2f}% {100*w/max(tot_w,1):6.2f}%")
