#!/usr/bin/env python
"""Instruction count per source line of one kernel: nvdisasm -g -c <cubin> | sass_lines.py <kernel-substring>"""
import re, sys, collections
pat = sys.argv[1] if len(sys.argv) > 1 else ""
on = False
cur = "?"
cnt = collections.Counter()
tot = 0
for line in sys.stdin:
    m = re.match(r"\s*\.text\.(\S+):", line) or re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        on = pat in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = m.group(1).split("/")[-1] + ":" + m.group(2)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        cnt[cur] += 1
        tot += 1
print("total", tot)
for k, v in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print(v, k)
