#!/usr/bin/env python
"""Benchmark of the calculable R-matrix hot path (BASELINE.json metric: partial-wave R-matrix
solves/sec on the KDUQ elastic sweep, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--samples S] [--impl reference]

One "step" = one pass of the hot path over the whole synthetic sweep: S KDUQ parameter samples
(416 real federal posterior rows + multivariate-normal draws fitted to them, seed 20241019) x
{n+40Ca, n+208Pb} x 32 energies linspace(5, 67, 32) MeV, N = 40, lmax = 30, 180 angles:
fused potential evaluation + assembly + LU + R -> S (kernel 1/2) and elastic observables (kernel 3).
For N > 1 (torchrun, one rank per GPU) the sample axis is sharded: every rank solves its own block of
S samples (weak scaling: per-GPU work fixed, the job is N x S samples), no traffic during the
solve, one NCCL gather of sigma_tot / sigma_rxn to rank 0 per step (the full d sigma/d Omega block of
a rank stays in its HBM / goes to its own pinned host buffer in the e2e leg; `--gather-dsdo` also
gathers it).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BASIS, LMAX, N_ANGLES = 40, 30, 180
TARGETS = ((40, 20, 37214.69), (208, 82, 193687.12278341982))  # (A, Z, rest mass MeV/c^2, AME2020)
MASS_N = 939.5654983938299
E_GRID = np.linspace(5.0, 67.0, 32)
FLOPS_PER_SOLVE = 8 * N_BASIS**3 / 3 + 8 * N_BASIS**2  # complex LU + one RHS = 183 467


def synthetic_kduq_samples(n, block=0):
    """rows 0..415: KDUQ federal neutron posterior; the rest: MVN fitted to them.
    ``block`` > 0 (ranks > 0 of a multi-GPU run) gives an independent block of draws only."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "kduq_params.npz"))
    real = g["federal_n"]
    if block == 0 and n <= len(real):
        return real[:n].copy()
    vary = [i for i in range(real.shape[1]) if np.std(real[:, i]) > 0]
    mean = real[:, vary].mean(axis=0)
    cov = np.cov(real[:, vary], rowvar=False)
    rng = np.random.default_rng(20241019 + block)
    out = np.tile(real[0], (n, 1))
    filled = 0
    if block == 0:
        out[: len(real)] = real
        filled = len(real)
    while filled < n:
        draw = rng.multivariate_normal(mean, cov, size=2 * (n - filled), method="cholesky", check_valid="ignore") \
            if np.all(np.linalg.eigvalsh(cov) > 0) else rng.multivariate_normal(mean, cov, size=2 * (n - filled))
        cand = np.tile(real[0], (len(draw), 1))
        cand[:, vary] = draw
        ok = np.ones(len(cand), dtype=bool)
        for A in (40, 208):  # positive radii / diffusenesses for both targets
            ok &= (cand[:, 12] - cand[:, 13] * A > 0.05) & (cand[:, 27] - cand[:, 28] * A > 0.05) & (cand[:, 36] > 0.05)
            ok &= (cand[:, 10] - cand[:, 11] * A ** (-1 / 3) > 0.3) & (cand[:, 25] - cand[:, 26] * A ** (1 / 3) > 0.3)
            ok &= cand[:, 34] - cand[:, 35] * A ** (-1 / 3) > 0.3
        cand = cand[ok][: n - filled]
        out[filled : filled + len(cand)] = cand
        filled += len(cand)
    return out


def energy_table():
    """[(target (A, Z), Elab, R_ch fm, kinematics tuple)] for the 64 (target, energy) workspaces."""
    from jitr_b200.utils import interaction_range
    from jitr_b200.utils.kinematics import semi_relativistic_kinematics

    tab = []
    for (A, Z, m) in TARGETS:
        for E in E_GRID:
            kin = semi_relativistic_kinematics(m, MASS_N, float(E), 0)
            Rch = (interaction_range(A) * kin.k + 2 * np.pi) / kin.k  # kduq_uq_demo.ipynb cell 5
            tab.append(((A, Z), float(E), float(Rch), tuple(kin)))
    return tab


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for (t, line) in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is
    pure Python and cannot travel to the GPU box) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_sweep

    rows = synthetic_kduq_samples(2048)
    tab = energy_table()
    vals = []
    procs = os.cpu_count() or 1
    ppw = args.cpu_pairs
    for i in range(args.warmup + args.steps):
        r = cpu_sweep.run(tab, rows, N_BASIS, LMAX, procs=procs, pairs_per_worker=ppw if i >= args.warmup else max(50, ppw // 10))
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["solves_per_s"] for r in vals]))
    ms = float(np.mean([r["seconds"] for r in vals]) * 1e3)
    sample = f"{procs} procs x {ppw} (sample, energy) pairs each of the same sweep (random energies round-robin), per step"
    line = {
        "impl": "reference", "metric": "partial-wave R-matrix solves/sec (KDUQ elastic sweep)", "value": v, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": workload_config(args.samples, args.gpus),
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": procs, "kind": "port", "sample": sample,
                         "samples_per_s": float(np.mean([r["samples_per_s"] for r in vals]))},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(samples, world=1):
    return {
        "workload": f"configs[1] KDUQ posterior sweep: {samples} samples per GPU x (n+40Ca, n+208Pb) x 32 energies 5-67 MeV, "
                    f"nbasis={N_BASIS}, lmax={LMAX}, {N_ANGLES} angles, elastic dsigma/dOmega + sigma_rxn, early break tol 1e-6",
        "samples_per_gpu": samples, "samples": samples * world, "pairs_per_step": samples * world * 64,
        "parallelism": f"dp{world}: sample axis sharded over ranks, no collective during the solve, one NCCL gather of the integral cross sections",
        "l2": "per-step working set (S-matrices + observables, >= 0.5 GB per 10^4 samples) exceeds the 126 MB L2; no flush needed",
    }


_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything libraries print to fd 1 (NCCL's version banner, ...) is sent
    to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=100000, help="KDUQ samples per GPU (weak scaling; configs[1] has 10^5)")
    ap.add_argument("--gather-dsdo", action="store_true", help="also gather the full dsigma/dOmega blocks to rank 0")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-pairs", type=int, default=1500, help="(sample, energy) pairs per CPU worker for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk", type=int, default=8192)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import jitr_b200
    from jitr_b200 import _cabi
    from jitr_b200.optical_potentials import kduq
    from jitr_b200.reactions import ElasticReaction

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _cabi.lib()

    # ---- setup (untimed): workspaces for the 64 (target, energy) rows, replicated on every rank
    solver = jitr_b200.Solver(N_BASIS)
    angles = np.linspace(0.1, np.pi, N_ANGLES)
    wss = []
    for (A, Z, m) in TARGETS:
        rx = ElasticReaction((A, Z), (1, 0), mass_target=m, mass_projectile=MASS_N)
        for (tz, E, Rch, kin) in [t for t in energy_table() if t[0] == (A, Z)]:
            wss.append(jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(E), Rch, solver, LMAX, angles))
    sweep = jitr_b200.ElasticSweep(wss, dev)
    model = kduq.KDUQ((1, 0))
    nE = sweep.nE

    # ---- inputs: this rank's block of the sample axis (rank 0 holds the real posterior rows)
    from jitr_b200.sharding import gather_to_root, shard_bounds

    S_total = args.samples * world
    lo, hi = shard_bounds(S_total, world, rank)
    rows_all = synthetic_kduq_samples(hi - lo, block=rank)
    rows = np.zeros((hi - lo, 40))
    rows[:, :37] = rows_all
    B = hi - lo
    params_host = torch.from_numpy(rows).pin_memory()
    params_dev = params_host.to(dev)
    Sbuf = sweep._alloc_S(B)
    dsdo = torch.empty((B, nE, N_ANGLES), dtype=torch.float64, device=dev)
    tr = torch.empty((B, nE, 2), dtype=torch.float64, device=dev)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(timed):
        """inputs resident in HBM -> S-matrices -> observables (-> gather).  Returns per-kernel ms if timed."""
        ev[0].record()
        sweep.smatrix_params(model, params_dev, out=Sbuf)
        ev[1].record()
        sweep.observables(Sbuf[0], Sbuf[1], want_Ay=False, want_Q=False, out=(dsdo, None, None, tr))
        ev[2].record()
        if world > 1:
            gather_to_root(tr, S_total, dst=0)
            if args.gather_dsdo:
                gather_to_root(dsdo, S_total, dst=0)
        ev[3].record()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    sync_all()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = L.jitr_b200_launch_count()
    fallbacks0 = L.jitr_b200_sym_fallback_count()
    t_wall0 = time.time()
    t_sweep = t_obs = 0.0
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_begin.record()
    per_step_events = []
    for _ in range(args.steps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[:] = evs
        step(True)
        per_step_events.append(evs)
    e_end.record()
    sync_all()
    t_wall1 = time.time()
    launches = L.jitr_b200_launch_count() - launches0
    fallbacks = L.jitr_b200_sym_fallback_count() - fallbacks0
    total_ms = e_begin.elapsed_time(e_end)
    for evs in per_step_events:
        t_sweep += evs[0].elapsed_time(evs[1])
        t_obs += evs[1].elapsed_time(evs[2])
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    nl = Sbuf[1]
    my_solves = int((2 * nl.to(torch.int64) - 1).sum().item())  # systems solved per step == the reference's count
    tt = torch.tensor([total_ms, float(my_solves), t_sweep, t_obs], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, t_sweep, t_obs = float(mx[0]), float(mx[2]), float(mx[3])
        all_solves = float(sm[1])
    else:
        all_solves = float(my_solves)
    ms_per_step = total_ms / args.steps
    value = all_solves / (ms_per_step * 1e-3)

    # ---- e2e: through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        # pinned host buffers of the whole block (9.2 GB per GPU at 10^5 samples)
        dsdo_host = torch.empty((B, nE, N_ANGLES), dtype=torch.float64).pin_memory()
        tr_host = torch.empty((B, nE, 2), dtype=torch.float64).pin_memory()
        sweep.xs_params_to_host(model, params_host, dsdo_host, tr_host, chunk=args.chunk)  # warm-up (buffers, streams)
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            ns = sweep.xs_params_to_host(model, params_host, dsdo_host, tr_host, chunk=args.chunk)
        b.record()
        sync_all()
        e_ms = torch.tensor([a.elapsed_time(b) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        assert torch.allclose(dsdo_host[:4].to(dev), dsdo[:4], rtol=1e-12, atol=0)
        e2e = {"value": all_solves / (float(e_ms[0]) * 1e-3), "unit": "solves/s", "ms_per_step": float(e_ms[0]),
               "h2d_bytes_per_step": int(params_host.numel() * 8 * world),
               "d2h_bytes_per_step": int((dsdo_host.numel() + tr_host.numel()) * 8 * world),
               "api": "ElasticSweep.xs_params_to_host (pinned host params -> pinned host dsdo, sigma_tot/rxn; chunked, copy/compute overlap)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused sweep): algorithmic flops / its CUDA-event time
    tf = __import__("ctypes").c_double(0)
    lat = (__import__("ctypes").c_longlong * 8)()
    peak_dfma = peak_dmma = None
    if L.jitr_b200_fp64_probe(0, 4096, 0, __import__("ctypes").byref(tf), lat) == 0:
        peak_dfma = tf.value
    if L.jitr_b200_fp64_probe(1, 4096, 0, __import__("ctypes").byref(tf), lat) == 0:
        peak_dmma = tf.value
    peak = max(peak_dfma or 0, peak_dmma or 0) or 37.2
    sweep_ms = t_sweep / args.steps
    achieved = my_solves * FLOPS_PER_SOLVE / (sweep_ms * 1e-3) / 1e12
    roofline = {
        "bound": "fp64", "kernel": "elastic_sweep_kernel<PARAMS=true, MODE=40> (fused potentials + assembly-free LU + R->S)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak, "traffic": traffic_bytes(B * nE),
        "peak_source": "measured live: FP64 pipe micro-benchmark in this library (max of DFMA and DMMA m8n8k4 issue-bound loops); "
                       "MEASURED_PEAKS.json has no FP64 entry; nominal 37.2",
        "peak_dfma_measured": peak_dfma, "peak_dmma_measured": peak_dmma, "frac_of_nominal_37.2": achieved / 37.2,
        "flops_per_solve": FLOPS_PER_SOLVE, "solves_per_launch": my_solves, "kernel_ms": sweep_ms,
        "kernel_share_of_step": sweep_ms / ms_per_step,
        "observables_kernel": {"bound": "hbm", "ms": t_obs / args.steps,
                               "achieved": (B * nE * (2 * (LMAX + 1) * 16 + N_ANGLES * 8 + 16)) / (t_obs / args.steps * 1e-3) / 1e9,
                               "unit": "GB/s", "peak": peaks().get("hbm_gbs")},
    }
    ok = roofline["observables_kernel"]
    ok["frac"] = ok["achieved"] / ok["peak"] if ok["peak"] else None
    # 8 flops per (kept partial wave, angle): at 2448 B per pair this kernel is FP64-bound, not HBM-bound
    obs_flops = 8.0 * N_ANGLES * float(nl.to(torch.float64).sum().item())
    ok["fp64_tflops"] = obs_flops / (t_obs / args.steps * 1e-3) / 1e12
    ok["fp64_frac"] = ok["fp64_tflops"] / peak

    cpu_baseline = None
    if not args.no_cpu_baseline:
        from oracle import cpu_sweep

        procs = os.cpu_count() or 1
        cpu_sweep.run(energy_table(), rows_all[:512], N_BASIS, LMAX, procs=procs, pairs_per_worker=60)  # warm
        r = cpu_sweep.run(energy_table(), rows_all[:2048], N_BASIS, LMAX, procs=procs, pairs_per_worker=args.cpu_pairs)
        cpu_baseline = {"value": r["solves_per_s"], "unit": "solves/s", "cores": procs, "kind": "port",
                        "sample": f"{procs} processes x {args.cpu_pairs} (sample, energy) pairs of the same sweep "
                                  f"({r['solves_per_pair']:.1f} solves/pair, {r['seconds']:.1f} s)",
                        "samples_per_s": r["samples_per_s"]}

    line = {
        "metric": "partial-wave R-matrix solves/sec (KDUQ elastic sweep)", "value": value, "unit": "solves/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": workload_config(args.samples, world),
        "samples_per_s": S_total * nE / (ms_per_step * 1e-3),
        "solves_per_pair": all_solves / (S_total * nE),
        "pivoting": {"scheme": "threshold pivoting on the diagonal (|a_kk| >= max_i |a_ik| / 64, symmetric LDL^T); a system that fails the test is re-solved by partial-pivoting LU",
                     "partial_pivoting_fallbacks_per_step_rank0": int(fallbacks // max(args.steps, 1)),
                     "fallback_fraction": float(fallbacks) / max(args.steps, 1) / max(my_solves, 1)},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def traffic_bytes(pairs):
    """DRAM bytes of one sweep launch, scaled per (sample, energy) pair from the committed ncu capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        return {"bytes": t["sweep_dram_bytes_per_pair"] * pairs, "per_pair": t["sweep_dram_bytes_per_pair"],
                "algorithmic_bytes_per_pair": 2 * (LMAX + 1) * 16 + 8 + 40 * 8 / 64.0, "source": t["source"]}
    except Exception:
        return None


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {"hbm_gbs": 6650.0}


if __name__ == "__main__":
    main()
