#!/usr/bin/env python
"""Markdown summary of `ncu --set full` reports: tools/ncu_summary.py title=report.ncu-rep[:kernel-regex] ... > profiles/rNN_ncu_kernels.md

Per report: the headline counters of the first kernel whose name matches, the calibrated FP64-unit utilisation (sum of the DFMA
and DMMA pipe counters, profiles/r02_fp64_pipe_calibration.md) and -- from the source page -- the LSU wavefronts per SASS opcode
(shared-memory loads/stores; a 32-bit SHFL occupies the same pipe for one wavefront)."""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def raw_rows(rep):
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def source_rows(rep, kernel_id):
    out = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(kernel_id), "--launch-count", "1"])
    rows = list(csv.reader(io.StringIO(out)))
    for i, r in enumerate(rows):
        if r and r[0] == "Address":
            return r, rows[i + 1:]
    return None, []


def main():
    print("# ncu --set full summaries (B200, --clock-control none; each capture after the same command had exited 0 without ncu)\n")
    for arg in sys.argv[1:]:
        title, rest = arg.split("=", 1)
        rep, _, rx = rest.partition(":")
        hdr, units, rows = raw_rows(rep)
        ix = {h: i for i, h in enumerate(hdr)}
        pick = None
        for n, r in enumerate(rows):
            if re.search(rx or ".", r[ix["Kernel Name"]]):
                pick = (n, r)
                break
        if pick is None:
            print(f"## {title}\n\n(no kernel matching {rx!r} in {rep})\n")
            continue
        n, r = pick
        print(f"## {title}\n\n`{r[ix['Kernel Name']]}`  grid {r[ix['Grid Size']]} block {r[ix['Block Size']]}\n\n| metric | unit | value |\n|---|---|---|")
        vals = {}
        for m in METRICS:
            if m in ix:
                vals[m] = r[ix[m]]
                print(f"| {m} | {units[ix[m]]} | {r[ix[m]]} |")
        try:
            busy = float(vals[METRICS[1]].replace(",", "")) + float(vals[METRICS[2]].replace(",", ""))
            print(f"\nFP64 unit busy (calibrated sum of the two pipe counters): **{busy:.1f} %**")
        except Exception:
            pass
        shdr, srows = source_rows(rep, n)
        if shdr:
            sx = {h: i for i, h in enumerate(shdr)}
            ex, wf, wfi = collections.Counter(), collections.Counter(), collections.Counter()
            for s in srows:
                if len(s) < len(shdr) - 2:
                    continue
                t = s[sx["Source"]].split()
                if not t:
                    continue
                op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).rstrip(";")
                ex[op] += int(s[sx["Instructions Executed"]] or 0)
                wf[op] += int(s[sx["L1 Wavefronts Shared"]] or 0)
                wfi[op] += int(s[sx["L1 Wavefronts Shared Ideal"]] or 0)
            tot = sum(ex.values())
            print(f"\nwarp instructions executed: {tot:.4g}; top opcodes (share of executed, shared-memory wavefronts / ideal):\n")
            print("| opcode | share | wavefronts | ideal |\n|---|---|---|---|")
            for op, c in ex.most_common(16):
                print(f"| {op} | {100 * c / max(tot, 1):.1f} % | {wf[op]:.4g} | {wfi[op]:.4g} |")
            shfl = sum(c for op, c in ex.items() if op.startswith("SHFL"))
            print(f"\nLSU wavefronts: shared-memory {sum(wf.values()):.4g} (ideal {sum(wfi.values()):.4g}) + shuffles {shfl:.4g}")
        print()


if __name__ == "__main__":
    main()
