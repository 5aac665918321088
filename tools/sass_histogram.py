#!/usr/bin/env python
"""SASS instruction histogram per kernel of the built library:  python tools/sass_histogram.py > profiles/rNN_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jitr_b200", "_lib", "libjitr_b200.so")
COLS = [("DMMA", r"^DMMA"), ("DFMA", r"^DFMA"), ("DMUL/DADD", r"^(DMUL|DADD)"), ("UBLKCP", r"^UBLKCP"), ("SYNCS", r"^SYNCS"), ("SHFL", r"^SHFL"),
        ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDL/STL (spills)", r"^(LDL|STL)"), ("REDUX", r"^(REDUX|CREDUX)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = {}
    cur = None
    hist = collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
    for k, d in zip(hist, dem):
        d = re.sub(r"\(anonymous namespace\)::|jitr::|<unnamed>::", "", d)
        names[k] = re.sub(r"\((?:jitr::)?(?:SweepArgs|GenArgs|ObsArgs).*$", "", d)
    print("# SASS instruction histogram per kernel of libjitr_b200.so (cuobjdump -sass, sm_100a; tools/sass_histogram.py)\n")
    print("Proof-of-path mnemonics: `DMMA` = FP64 tensor path (mma.sync m8n8k4 f64; Blackwell has no tcgen05 FP64 path, so no UTC*MMA is expected), "
          "`UBLKCP` = cp.async.bulk (TMA bulk copy), `SYNCS` = mbarrier, `REDUX`/`CREDUX` = warp reductions.\n")
    print("| kernel | instructions | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
        h = hist[k]
        row = [sum(v for op, v in h.items() if re.match(rx, op)) for _, rx in COLS]
        print(f"| `{names[k]}` | {sum(h.values())} | " + " | ".join(str(v) for v in row) + " |")


if __name__ == "__main__":
    sys.exit(main())
