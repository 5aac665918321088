#!/usr/bin/env python
"""Benchmark of the calculable R-matrix hot path (BASELINE.json metric: partial-wave R-matrix solves/sec).

    python bench.py [--config 2] [--gpus N] [--steps K] [--warmup W] [--samples S] [--impl reference]

--config selects the BASELINE.json workload; the default, 2, is the configuration the metric is quoted on:
  1  n+208Pb 30 MeV, N = 40, lmax = 30, KD03: ONE parameter row (the reference's own CPU-runnable case; latency bound)
  2  KDUQ posterior sweep: S samples per GPU (416 federal rows + multivariate-normal draws) x {n+40Ca, n+208Pb} x 32 energies,
     N = 40, lmax = 30, 180 angles: fused potentials + LDL^T + R -> S (kernel 1), observables (kernel 2)
  3  p+48Ca 30 MeV, N = 60, lmax = 40, Coulomb + spin-orbit, KDUQ proton rows (416 federal + draws), dsigma/dOmega and A_y
  4  n+208Pb 30 MeV, N = 60, one full N x N complex Perey-Buck kernel per (sample, l), l = 0..30 (57.6 KB per system from HBM)
  5  3 coupled channels x N = 50 (examples/coupled.py), l = 0..10, S samples varying V and the couplings by 5 %
One "step" = one pass of the hot path over the whole synthetic batch.  `value`: inputs resident in HBM; `e2e`: the same pass
through the public API with pinned HOST buffers in and out.  For N > 1 (torchrun, one rank per GPU) the sample axis is sharded
(weak scaling: per-GPU work fixed), nothing is exchanged during the solve, and ONE NCCL gather per step brings the angular
distributions and integral cross sections of every rank to rank 0 (config 2: 1.15 GB per 12 500 samples).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from jitr_b200 import workloads as W  # noqa: E402  (host-side numpy only)

METRIC = "partial-wave R-matrix solves/sec"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
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
        for _ in range(20):  # a timed region shorter than the sampling period (config 1: 0.4 ms): wait for the next samples
            if any(t0 <= t <= t1 for (t, _l) in self.rows) or len([1 for (t, _l) in self.rows if t > t1]) >= 2:
                break
            time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        inside = [(t, l) for (t, l) in self.rows if t0 <= t <= t1] or self.rows[-3:]
        for (t, line) in inside:
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


_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything libraries print to fd 1 (NCCL's banner, ...) goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {"hbm_gbs": 6650.0}


def fp64_peak(L):
    """FP64 roofline denominator, measured live with this library's probes (MEASURED_PEAKS.json has no FP64 entry)."""
    tf = ctypes.c_double(0)
    lat = (ctypes.c_longlong * 8)()
    dfma = dmma = None
    if L.jitr_b200_fp64_probe(0, 4096, 0, ctypes.byref(tf), lat) == 0:
        dfma = tf.value
    if L.jitr_b200_fp64_probe(1, 4096, 0, ctypes.byref(tf), lat) == 0:
        dmma = tf.value
    return (max(dfma or 0, dmma or 0) or 37.2), dfma, dmma


# ====================================================================================================================
# workloads: each provides  setup / step (device-resident inputs) / e2e_step (host buffers) / cpu runs
# ====================================================================================================================
def _mark(torch):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


class Config2:
    """KDUQ posterior sweep (the headline).  Also configs 1 and 3 (single-workspace variants of the same fused sweep)."""

    def __init__(self, args, cfg):
        self.args, self.cfg = args, cfg
        self.N, self.lmax, self.nch = (60, 40, 1) if cfg == 3 else (40, 30, 1)
        self.projectile = (1, 1) if cfg == 3 else (1, 0)
        self.want_Ay = cfg == 3
        if args.samples is None:
            args.samples = {1: 1, 2: 100000, 3: 20000}[cfg]

    def describe(self, world):
        S = self.args.samples
        if self.cfg == 1:
            w = "configs[0] n+208Pb 30 MeV, KD03 (one parameter row), nbasis=40, lmax=30, R_ch=12 fm, 180 angles; one fused sweep call per step (latency bound)"
        elif self.cfg == 3:
            w = (f"configs[2] p+48Ca 30 MeV, nbasis=60, lmax=40, R_ch=12 fm, KDUQ proton posterior incl. Coulomb and spin-orbit: {S} rows per GPU "
                 f"(416 federal + multivariate-normal draws), dsigma/dOmega + A_y at 180 angles")
        else:
            w = (f"configs[1] KDUQ posterior sweep: {S} samples per GPU x (n+40Ca, n+208Pb) x 32 energies 5-67 MeV, nbasis=40, lmax=30, "
                 f"180 angles, elastic dsigma/dOmega + sigma_rxn, early break tol 1e-6")
        return {"workload": w, "samples_per_gpu": S, "samples": S * world, "pairs_per_step": S * world * self.nE,
                "parallelism": f"dp{world}: sample axis sharded over ranks, no collective during the solve, one NCCL gather of dsigma/dOmega + sigma_tot/rxn to rank 0 per step",
                "l2": "per-step working set (S-matrices + observables, >= 0.5 GB per 10^4 samples) exceeds the 126 MB L2; no flush needed"
                      if self.cfg != 1 else "single-sample latency case: 35 KB working set, L2 resident by definition of the workload"}

    def workspaces(self):
        if self.cfg == 1:
            return [W.config1_workspace()[1]]
        if self.cfg == 3:
            return [W.config3_workspace()[1]]
        return W.config2_workspaces()

    def setup(self, dev, rank, world):
        import torch

        import jitr_b200
        from jitr_b200.optical_potentials import kduq
        from jitr_b200.sharding import shard_bounds

        self.torch, self.dev, self.rank, self.world = torch, dev, rank, world
        self.sweep = jitr_b200.ElasticSweep(self.workspaces(), dev)
        self.model = kduq.KDUQ(self.projectile)
        self.nE = self.sweep.nE
        self.S_total = self.args.samples * world
        lo, hi = shard_bounds(self.S_total, world, rank)
        if self.cfg == 1:
            rows_all = kduq.get_kd03((1, 0))[None, :]
        else:
            rows_all = W.synthetic_kduq_samples(hi - lo, block=rank, projectile=self.projectile)
        self.rows_all = rows_all
        rows = np.zeros((hi - lo, 40))
        rows[:, : rows_all.shape[1]] = rows_all
        self.B = B = hi - lo
        self.params_host = torch.from_numpy(rows).pin_memory()
        self.params_dev = self.params_host.to(dev)
        self.Sbuf = self.sweep._alloc_S(B)
        nth = self.sweep.n_angles
        self.dsdo = torch.empty((B, self.nE, nth), dtype=torch.float64, device=dev)
        self.Ay = torch.empty_like(self.dsdo) if self.want_Ay else None
        self.tr = torch.empty((B, self.nE, 2), dtype=torch.float64, device=dev)
        self.gather_out = None
        self.windows = None
        if world > 1 and self.args.gather_mode == "p2p" and self.S_total == B * world:
            # root-resident windows written by copy-engine peer copies (sharding.PeerWindow); any failure -> the NCCL gather
            try:
                from jitr_b200.sharding import PeerWindow

                wins = [PeerWindow((self.S_total, self.nE, 2), torch.float64, dev)]
                if not self.args.no_gather_dsdo:
                    wins.append(PeerWindow((self.S_total, self.nE, nth), torch.float64, dev))
                ok = torch.ones(1, dtype=torch.int32, device=dev)
            except Exception as exc:  # noqa: BLE001
                sys.stderr.write(f"[bench] rank {rank}: peer windows unavailable ({exc!r}); using the NCCL gather\n")
                wins, ok = None, torch.zeros(1, dtype=torch.int32, device=dev)
            import torch.distributed as dist

            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            self.windows = wins if int(ok.item()) == 1 else None
        if world > 1 and self.windows is None and not self.args.no_gather_dsdo and rank == 0:
            self.gather_out = (torch.empty((self.S_total, self.nE, nth), dtype=torch.float64, device=dev),
                               torch.empty((self.S_total, self.nE, 2), dtype=torch.float64, device=dev))
        self.ev = None
        self.kernel = f"elastic_sweep_kernel<PARAMS=true, MODE={40 if self.N == 40 else 64}> (fused potentials + {'assembly-free' if self.N == 40 else 'packed lower-triangle'} LDL^T + R->S)"

    def step(self, timed_events=None):
        """One pass of the hot path over this rank's block.  N > 1: the block is processed in `--gather-chunks` chunks of samples and
        every chunk's observables start their way to rank 0 (one asynchronous NCCL gather per chunk, into the root's preallocated
        slices) while the next chunk is being solved: the 9.3 GB per rank of a 10^5-sample sweep (65 GB into the root at N = 8)
        hide behind the elimination except for the last chunk.  timed_events: a list the step appends (kind, start, end) to."""
        from jitr_b200.sharding import gather_chunk_async

        torch = self.torch
        ev = timed_events
        mark = (lambda: _mark(torch)) if ev is not None else (lambda: None)
        B = self.B
        nchunk = max(1, min(self.args.gather_chunks if self.windows is not None else self.args.nccl_gather_chunks, B)) if self.world > 1 else 1
        nchunk = min(nchunk, max(1, (B * self.nE) // 100000))  # a chunk keeps >= 1e5 (sample, energy) pairs: ~100 per resident warp
        equal = self.S_total == B * self.world
        if self.world > 1 and not equal:
            nchunk = 1
        bounds = [(c * B // nchunk, (c + 1) * B // nchunk) for c in range(nchunk)]
        go = self.gather_out
        offs = [r * B for r in range(self.world)]
        works = []
        S, nl, st = self.Sbuf
        t_last = None
        for (c0, c1) in bounds:
            t0 = mark()
            self.sweep.smatrix_params(self.model, self.params_dev[c0:c1], out=(S[c0:c1], nl[c0:c1], st[c0:c1]))
            t1 = mark()
            self.sweep.observables(S[c0:c1], nl[c0:c1], want_Ay=self.want_Ay, want_Q=False,
                                   out=(self.dsdo[c0:c1], None if self.Ay is None else self.Ay[c0:c1], None, self.tr[c0:c1]))
            t2 = mark()
            if ev is not None:
                ev.append(("k1", t0, t1))
                ev.append(("k2", t1, t2))
            t_last = t2
            if self.windows is not None:
                self.windows[0].push(self.tr[c0:c1], self.rank * B + c0)
                if len(self.windows) > 1:
                    self.windows[1].push(self.dsdo[c0:c1], self.rank * B + c0)
            elif self.world > 1 and equal:
                works.append(gather_chunk_async(self.tr[c0:c1], None if go is None else go[1], offs, c0, c1))
                if not self.args.no_gather_dsdo:
                    works.append(gather_chunk_async(self.dsdo[c0:c1], None if go is None else go[0], offs, c0, c1))
        if self.world > 1 and not equal:  # uneven blocks: the padded one-shot gather
            from jitr_b200.sharding import gather_to_root

            gather_to_root(self.tr, self.S_total, dst=0, out=None if go is None else go[1])
            if not self.args.no_gather_dsdo:
                gather_to_root(self.dsdo, self.S_total, dst=0, out=None if go is None else go[0])
        for w_ in works:
            w_.wait()
        if self.windows is not None:
            for w_ in self.windows:
                w_.flush()
        if ev is not None:
            ev.append(("coll", t_last, mark()))  # what is left of the transfers after the last chunk's kernels: the exposed part

    def collective_check(self):
        """Every rank's block sums against the sums of the rows the root received for it (all ranks call this)."""
        import torch.distributed as dist

        torch = self.torch
        mine = torch.stack([self.tr.sum(), self.dsdo.sum()]).reshape(1, 2)
        allv = torch.empty((self.world, 2), dtype=torch.float64, device=self.dev)
        dist.all_gather_into_tensor(allv, mine)
        if self.rank != 0:
            return None
        got_tr = self.windows[0].buf if self.windows is not None else (self.gather_out[1] if self.gather_out is not None else None)
        got_ds = (self.windows[1].buf if len(self.windows) > 1 else None) if self.windows is not None else (
            self.gather_out[0] if self.gather_out is not None else None)
        ok = True
        B = self.B
        for r in range(self.world):
            if got_tr is not None:
                ok &= bool(torch.allclose(got_tr[r * B:(r + 1) * B].sum(), allv[r, 0], rtol=1e-12, atol=0.0))
            if got_ds is not None:
                ok &= bool(torch.allclose(got_ds[r * B:(r + 1) * B].sum(), allv[r, 1], rtol=1e-12, atol=0.0))
        return ok

    def collective(self):
        if self.world == 1:
            return None
        per_rank = self.tr.numel() * 8 + (0 if self.args.no_gather_dsdo else self.dsdo.numel() * 8)
        what = "dsigma/dOmega [S/G, E, n_angles] + (sigma_tot, sigma_rxn) [S/G, E, 2] to rank 0"
        if self.windows is not None:
            op = (f"copy-engine peer copies (CUDA IPC window on rank 0, cudaMemcpyAsync over NVLink, no SMs) of {what}, in up to {self.args.gather_chunks} chunks of "
                  "samples per step (>= 1e5 pairs each), each queued behind its chunk's kernels and overlapped with the next chunk's solve, + one 4-byte NCCL all_reduce as "
                  "the completion barrier; ms_per_step = what stays exposed after the last chunk's kernels")
            chunks = min(int(self.args.gather_chunks), max(1, (self.B * self.nE) // 100000))
        else:
            chunks = int(self.args.nccl_gather_chunks)
            op = f"torch.distributed.gather (NCCL) of {what}, once per step" + ("" if chunks == 1 else f" in {chunks} asynchronous chunks")
        return {"op": op, "mode": "p2p" if self.windows is not None else "nccl", "chunks": chunks, "bytes_per_rank": int(per_rank),
                "bytes_into_root": int(per_rank * (self.world - 1))}

    def solves(self):
        return int((2 * self.Sbuf[1].to(self.torch.int64) - 1).sum().item())

    def status_bad(self):
        return int((self.Sbuf[2] != 0).sum().item())

    def flops_per_solve(self):
        return W.flops_per_solve(self.N)

    def observable_bytes(self):
        nth = self.sweep.n_angles
        return self.B * self.nE * (2 * (self.lmax + 1) * 16 + nth * 8 * (2 if self.want_Ay else 1) + 16)

    def e2e_setup(self):
        torch = self.torch
        nth = self.sweep.n_angles
        self.dsdo_host = torch.empty((self.B, self.nE, nth), dtype=torch.float64).pin_memory()
        self.tr_host = torch.empty((self.B, self.nE, 2), dtype=torch.float64).pin_memory()
        self.st_host = torch.empty((self.B, self.nE), dtype=torch.int32).pin_memory()

    def e2e_step(self):
        self.sweep.xs_params_to_host(self.model, self.params_host, self.dsdo_host, self.tr_host, chunk=self.args.chunk, status_host=self.st_host)

    def e2e_check(self):
        assert self.torch.allclose(self.dsdo_host[:4].to(self.dev), self.dsdo[:4], rtol=1e-12, atol=0)
        assert int((self.st_host != 0).sum()) == self.status_bad()

    def e2e_bytes(self):
        return (int(self.params_host.numel() * 8),
                int((self.dsdo_host.numel() + self.tr_host.numel()) * 8 + self.st_host.numel() * 4),
                "ElasticSweep.xs_params_to_host (pinned host parameter rows -> pinned host dsdo, sigma_tot/rxn, status; chunked, copy/compute overlap)")

    # ---- CPU arm: the oracle port of the reference's user loop on all host cores
    def cpu_run(self, pairs_per_worker, warm=False):
        from oracle import cpu_sweep

        procs = os.cpu_count() or 1
        if self.cfg == 2:
            tab, rows = W.energy_table(), W.synthetic_kduq_samples(2048)
        else:
            ws = self.workspaces()[0] if not hasattr(self, "sweep") else None
            if self.cfg == 1:
                from jitr_b200.optical_potentials import kduq
                from jitr_b200.utils.kinematics import semi_relativistic_kinematics

                kin = semi_relativistic_kinematics(W.TARGETS[1][2], W.MASS_N, 30.0, 0)
                tab, rows = [((208, 82), 30.0, 12.0, tuple(kin))], kduq.get_kd03((1, 0))[None, :]
            else:
                from jitr_b200.utils.kinematics import semi_relativistic_kinematics

                kin = semi_relativistic_kinematics(W.MASS_48CA, W.MASS_P, 30.0, 20)
                tab, rows = [((48, 20), 30.0, 12.0, tuple(kin))], W.synthetic_kduq_samples(2048, projectile=(1, 1))
            del ws
        r = cpu_sweep.run(tab, rows, self.N, self.lmax, procs=procs, pairs_per_worker=max(50, pairs_per_worker // 10) if warm else pairs_per_worker,
                          projectile=self.projectile)
        r["sample"] = (f"{procs} processes x {pairs_per_worker} (sample, energy) pairs of the same workload "
                       f"({r['solves_per_pair']:.1f} solves/pair, {r['seconds']:.1f} s)")
        return r


class Config45:
    """General batched solve: config 4 (nonlocal N = 60, kernels generated on the device) and config 5 (3 x 50 coupled)."""

    def __init__(self, args, cfg):
        self.args, self.cfg = args, cfg
        c = W.CONFIG4 if cfg == 4 else W.CONFIG5
        self.N, self.lmax = c["N"], c["lmax"]
        self.nch = 1 if cfg == 4 else 3
        if args.samples is None:
            args.samples = 10000

    def describe(self, world):
        S = self.args.samples
        if self.cfg == 4:
            w = (f"configs[3] n+208Pb 30 MeV, nbasis=60, nonlocal Perey-Buck-type kernel U_l(r,r') = -WS((r+r')/2) H_l(r,r';beta), one full 60x60 complex "
                 f"kernel (57.6 KB) per (sample, l), l = 0..30, {S} samples per GPU of (V, W, beta) ~ N(71,3), N(15,2), N(0.85,0.03)")
        else:
            w = (f"configs[4] 3 coupled channels x nbasis=50 (n = 150), examples/coupled.py geometry (real Woods-Saxon + Coulomb, surface-Gaussian couplings), "
                 f"l = 0..10, {S} samples per GPU varying V and the couplings by +-5 %")
        return {"workload": w, "samples_per_gpu": S, "samples": S * world, "systems_per_step": S * world * (self.lmax + 1),
                "parallelism": f"dp{world}: sample axis sharded over ranks, no collective during the solve, one NCCL gather of the S-matrices to rank 0 per step",
                "l2": "inputs per step (17.9 GB of kernels at 10^4 samples) exceed the 126 MB L2" if self.cfg == 4 else
                      "1.1e5 systems x 360 KB working set each stream through shared memory / L2 slabs; potentials (72 MB) + outputs exceed nothing but are touched once per step"}

    def setup(self, dev, rank, world):
        import torch

        import jitr_b200
        from jitr_b200.reactions import ProjectileTargetSystem

        self.torch, self.dev, self.rank, self.world = torch, dev, rank, world
        S = self.B = self.args.samples
        L1 = self.lmax + 1
        self.solver = solver = jitr_b200.Solver(self.N)
        t = lambda v, dt=torch.complex128: torch.tensor(np.ascontiguousarray(v), dtype=dt, device=dev)  # noqa: E731
        if self.cfg == 4:
            from jitr_b200.utils.kinematics import semi_relativistic_kinematics

            kin = semi_relativistic_kinematics(W.TARGETS[1][2], W.MASS_N, W.CONFIG4["Elab"], 0)
            self.a = a = W.CONFIG4["Rch"] * kin.k
            self.E0, self.k0 = kin.Ecm, kin.k
            sysm = ProjectileTargetSystem(a, self.lmax, Ztarget=82, Zproj=0)
            channels, asym = sysm.get_partial_wave_channels(*kin)
            self.free = [t(solver.free_matrix(a, [l])) for l in range(L1)]
            self.asym = [t(np.array([[asym[l].Hp[0]], [asym[l].Hm[0]], [asym[l].Hpp[0]], [asym[l].Hmp[0]]])) for l in range(L1)]
            smp = W.config4_samples(S, seed=W.CONFIG4["seed"] + rank)
            rows = np.concatenate([smp, np.full((S, 1), W.CONFIG4["R"]), np.full((S, 1), W.CONFIG4["a"])], axis=1)
            self.rows_host = torch.from_numpy(np.ascontiguousarray(rows)).pin_memory()
            self.rgrid = solver.radial_grid(a, kin.k)
            self.U = torch.empty((L1, S, self.N, self.N), dtype=torch.complex128, device=dev)  # 17.9 GB at 10^4 samples
            self._generate(self.rows_host.to(dev))
            self.kernel = "packed_solve_kernel<64> (bulk-async staged lower triangle + packed LDL^T, one warp per system)"
        else:
            c = W.CONFIG5
            self.a, self.E0, self.k0 = c["a"], c["E"][0], c["k"]
            cm = np.array(c["coupling"])
            sysc = ProjectileTargetSystem(channel_radius=c["a"], lmax=self.lmax, Ztarget=20, Zproj=1, channel_levels=np.array(c["levels"]),
                                          coupling=lambda l: cm)
            channels, asym = sysc.get_partial_wave_channels(c["Elab"], np.array(c["E"]), np.full(3, c["mu"]), np.full(3, c["k"]), np.full(3, c["eta"]))
            self.free = [t(solver.free_matrix(c["a"], [l] * 3, list(c["E"]), [c["mu"]] * 3)) for l in range(L1)]
            self.asym = [t(np.array([asym[l].Hp, asym[l].Hm, asym[l].Hpp, asym[l].Hmp])) for l in range(L1)]
            r = solver.radial_grid(c["a"], c["k"])
            V = W.config5_potentials(r, S, seed=c["seed"] + rank)
            self.V_host = torch.from_numpy(V).pin_memory()
            self.V = self.V_host.to(dev)
            self.kernel = "block_solve_kernel<256> (CTA per system, blocked LDL^T / partial-pivoting LU on DMMA)"
        self.Sout = torch.empty((L1, S, self.nch, self.nch), dtype=torch.complex128, device=dev)
        self.status = torch.empty((L1, S), dtype=torch.int32, device=dev)
        self.nE = L1
        self.gather_out = torch.empty((world, L1, S, self.nch, self.nch), dtype=torch.complex128, device=dev) if (world > 1 and rank == 0) else None

    def _generate(self, rows_dev):
        from jitr_b200.optical_potentials import nonlocal_forms

        nonlocal_forms.perey_buck_kernels(self.rgrid, self.lmax, rows_dev[:, 0].cpu().numpy(), rows_dev[:, 1].cpu().numpy(), rows_dev[:, 2].cpu().numpy(),
                                          R=W.CONFIG4["R"], a=W.CONFIG4["a"], l_major=True, out=self.U, device=self.dev)

    def _solve_all(self, U=None, V=None):
        """ONE batched call for all (partial wave, sample) systems: the free matrix is picked per system (free_index = l), the
        asymptotics are per system (jitr_b200_solve_batched_indexed)."""
        torch = self.torch
        L1, S = self.lmax + 1, self.B
        if not hasattr(self, "_free_stack"):
            self._free_stack = torch.stack(self.free).contiguous()
            self._free_index = torch.arange(L1, dtype=torch.int32, device=self.dev).repeat_interleave(S).contiguous()
            self._asym_sys = torch.stack(self.asym)[self._free_index.long()].contiguous()   # (L1 * S, 4, nch)
        if self.cfg == 4:
            Uall = (self.U if U is None else U).reshape(L1 * S, self.N, self.N)
            o = self.solver.solve_batch(self.a, self.E0, self.k0, 1, self._free_stack, self._asym_sys, nonlocal_potential=Uall,
                                        free_index=self._free_index)
        else:
            Vall = (self.V if V is None else V).unsqueeze(0).expand(L1, S, 3, 3, self.N).reshape(L1 * S, 3, 3, self.N)
            o = self.solver.solve_batch(self.a, self.E0, self.k0, 3, self._free_stack, self._asym_sys, local_potential=Vall,
                                        free_index=self._free_index)
        self.Sout.copy_(o["S"].reshape(L1, S, self.nch, self.nch))
        self.status.copy_(o["status"].reshape(L1, S))

    def step(self, timed_events=None):
        import torch.distributed as dist

        ev = timed_events
        mark = (lambda: _mark(self.torch)) if ev is not None else (lambda: None)
        t0 = mark()
        self._solve_all()
        t1 = mark()
        if self.world > 1:
            # (NCCL has no complex type: the S-matrices travel as interleaved float64 pairs)
            dist.gather(self.torch.view_as_real(self.Sout), [self.torch.view_as_real(g) for g in self.gather_out.unbind(0)] if self.rank == 0 else None, dst=0)
        if ev is not None:
            ev.append(("k1", t0, t1))
            ev.append(("coll", t1, mark()))

    def collective(self):
        if self.world == 1:
            return None
        b = self.Sout.numel() * 16
        return {"op": "torch.distributed.gather (NCCL) of the S-matrices [lmax+1, S/G, nch, nch] to rank 0, once per step",
                "bytes_per_rank": int(b), "bytes_into_root": int(b * (self.world - 1))}

    def solves(self):
        return self.B * (self.lmax + 1)

    def status_bad(self):
        return int((self.status != 0).sum().item())

    def flops_per_solve(self):
        return W.flops_per_solve(self.N * self.nch, self.nch)

    def e2e_setup(self):
        torch = self.torch
        self.S_host = torch.empty(tuple(self.Sout.shape), dtype=torch.complex128).pin_memory()
        self.st_host = torch.empty(tuple(self.status.shape), dtype=torch.int32).pin_memory()

    def e2e_step(self):
        if self.cfg == 4:
            # public API from parameter rows: H2D of the (V, W, beta, R, a) rows, kernels generated on the device, solves, S back
            self._generate(self.rows_host.to(self.dev, non_blocking=True))
            self._solve_all()
        else:
            self._solve_all(V=self.V_host.to(self.dev, non_blocking=True))
        self.S_host.copy_(self.Sout, non_blocking=True)
        self.st_host.copy_(self.status, non_blocking=True)

    def e2e_check(self):
        self.torch.cuda.synchronize()
        assert int((self.st_host != 0).sum()) == 0

    def e2e_bytes(self):
        if self.cfg == 4:
            return (int(self.rows_host.numel() * 8), int(self.S_host.numel() * 16 + self.st_host.numel() * 4),
                    "nonlocal_forms.perey_buck_kernels (pinned host (V, W, beta, R, a) rows -> kernels generated in HBM) + ONE Solver.solve_batch(free_index=l) -> pinned host S, status")
        return (int(self.V_host.numel() * 16), int(self.S_host.numel() * 16 + self.st_host.numel() * 4),
                "ONE Solver.solve_batch(free_index=l) (pinned host coupling potentials [S,3,3,50] -> pinned host S, status)")

    def cpu_run(self, per_worker, warm=False):
        from oracle import cpu_sweep

        procs = os.cpu_count() or 1
        n = max(20, per_worker // 10) if warm else per_worker
        r = cpu_sweep.run_general("config4" if self.cfg == 4 else "config5", procs=procs, systems_per_worker=n)
        r["sample"] = f"{procs} processes x {n} systems of the same workload (Solver.solve with a precomputed free matrix), {r['seconds']:.1f} s"
        r["samples_per_s"] = r["solves_per_s"] / (self.lmax + 1)
        return r


def make_workload(args):
    return Config2(args, args.config) if args.config in (1, 2, 3) else Config45(args, args.config)


def cpu_default(args):
    """bounded CPU samples: (sample, energy) pairs per worker for the elastic configs, systems per worker for 4 / 5"""
    if args.cpu_pairs is not None:
        return args.cpu_pairs
    return {1: 400, 2: 2000, 3: 400, 4: 6000, 5: 600}[args.config]


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is pure Python and
    cannot travel to the GPU box) on all host cores, on this arm's config / metric / unit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = make_workload(args)
    if isinstance(wl, Config2):
        wl.nE = {1: 1, 2: 64, 3: 1}[args.config]
    n = cpu_default(args)
    vals = []
    for i in range(args.warmup + args.steps):
        r = wl.cpu_run(n, warm=i < args.warmup)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["solves_per_s"] for r in vals]))
    ms = float(np.mean([r["seconds"] for r in vals]) * 1e3)
    procs = os.cpu_count() or 1
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": wl.describe(args.gpus),
        "cpu_baseline": {"value": v, "unit": "solves/s", "cores": procs, "kind": "port", "sample": vals[-1]["sample"] + ", per step",
                         "samples_per_s": float(np.mean([r["samples_per_s"] for r in vals]))},
        "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=None, help="samples per GPU (weak scaling); default: the configuration's own size")
    ap.add_argument("--no-gather-dsdo", action="store_true", help="N > 1: gather sigma_tot / sigma_rxn only")
    ap.add_argument("--gather-mode", default="p2p", choices=["nccl", "p2p"],
                    help="N > 1: how the observables reach rank 0 -- p2p: copy-engine peer copies into a CUDA-IPC window on rank 0, chunk by chunk beside "
                         "the solve (falls back to nccl if the window cannot be set up); nccl: one torch.distributed.gather at the end of the step")
    ap.add_argument("--gather-chunks", type=int, default=4, help="p2p mode: chunks of samples per step; a chunk's observables travel to rank 0 while the next is solved")
    ap.add_argument("--nccl-gather-chunks", type=int, default=1, help="nccl mode: > 1 launches one asynchronous gather per chunk (measured slower: its copy kernels wait for SMs)")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-pairs", type=int, default=None, help="CPU arm: (sample, energy) pairs (configs 1-3) or systems (4, 5) per worker")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N = 1: skip the parity record (config 2: the large-sample gate; configs 3-5: tools/parity_general.py)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling sub-record (configuration size in total)")
    ap.add_argument("--chunk", type=int, default=8192)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from jitr_b200 import _cabi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _cabi.lib()

    wl = make_workload(args)
    wl.setup(dev, rank, world)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_run(w, steps, warmup, clocks_on):
        for _ in range(warmup):
            w.step()
        sync_all()
        clocks = ClockSampler(local) if (clocks_on and rank == 0) else None
        l0, f0 = L.jitr_b200_launch_count(), L.jitr_b200_sym_fallback_count()
        t0 = time.time()
        e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        per_step = []
        e_begin.record()
        for _ in range(steps):
            w.step(per_step)
        e_end.record()
        sync_all()
        t1 = time.time()
        total = e_begin.elapsed_time(e_end)
        t_k1 = sum(a.elapsed_time(b) for k, a, b in per_step if k == "k1")
        t_k2 = sum(a.elapsed_time(b) for k, a, b in per_step if k == "k2")
        t_coll = sum(a.elapsed_time(b) for k, a, b in per_step if k == "coll")
        my_solves = w.solves()
        tt = torch.tensor([total, float(my_solves), t_k1, t_k2, t_coll], dtype=torch.float64, device=dev)
        all_solves = float(my_solves)
        if world > 1:
            mx, sm = tt.clone(), tt.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            total, t_k1, t_k2, t_coll = float(mx[0]), float(mx[2]), float(mx[3]), float(mx[4])
            all_solves = float(sm[1])
        return dict(ms_per_step=total / steps, k1_ms=t_k1 / steps, k2_ms=t_k2 / steps, coll_ms=t_coll / steps, all_solves=all_solves,
                    my_solves=my_solves, launches=L.jitr_b200_launch_count() - l0, fallbacks=L.jitr_b200_sym_fallback_count() - f0,
                    clocks=clocks.stop(t0, t1) if clocks else None)

    main_run = timed_run(wl, args.steps, args.warmup, True)
    collective_desc = wl.collective() if world > 1 else None
    if world > 1 and hasattr(wl, "collective_check") and wl.S_total == wl.B * world:
        torch.cuda.synchronize()
        collective_desc["received_block_sums_match"] = wl.collective_check()
    ms_per_step = main_run["ms_per_step"]
    value = main_run["all_solves"] / (ms_per_step * 1e-3)
    status_bad = wl.status_bad()

    # ---- e2e: through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        wl.e2e_setup()
        wl.e2e_step()  # warm-up (buffers, streams)
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            wl.e2e_step()
        b.record()
        sync_all()
        e_ms = torch.tensor([a.elapsed_time(b) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        wl.e2e_check()
        h2d, d2h, api = wl.e2e_bytes()
        e2e = {"value": main_run["all_solves"] / (float(e_ms[0]) * 1e-3), "unit": "solves/s", "ms_per_step": float(e_ms[0]),
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "api": api}

    # ---- strong scaling sub-record: the configuration's size IN TOTAL, sharded over the ranks
    strong = None
    if world > 1 and not args.no_strong and args.config in (2, 3, 4, 5):
        total_samples = args.samples
        sargs = argparse.Namespace(**vars(args))
        sargs.samples = max(1, total_samples // world)
        ws = make_workload(sargs)
        del wl.gather_out
        wl.gather_out = None
        if getattr(wl, "windows", None) is not None:
            for w_ in wl.windows:
                w_.close()
            wl.windows = None
        torch.cuda.empty_cache()
        ws.setup(dev, rank, world)
        r = timed_run(ws, args.steps, max(1, args.warmup), False)
        strong = {"samples_total": sargs.samples * world, "samples_per_gpu": sargs.samples, "ms_per_step": r["ms_per_step"],
                  "value": r["all_solves"] / (r["ms_per_step"] * 1e-3), "unit": "solves/s", "scaling": "strong",
                  "collective_ms_per_step": r["coll_ms"], "collective": ws.collective()}

    for w_obj in (wl, locals().get("ws")):
        wins = getattr(w_obj, "windows", None) if w_obj is not None else None
        if wins is not None:
            for w_ in wins:
                w_.close()
            w_obj.windows = None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: algorithmic flops / its CUDA-event time
    peak, peak_dfma, peak_dmma = fp64_peak(L)
    k_ms = main_run["k1_ms"]
    F = wl.flops_per_solve()
    achieved = main_run["my_solves"] * F / (k_ms * 1e-3) / 1e12
    n_sys = wl.N * wl.nch
    roofline = {
        "bound": "fp64", "kernel": wl.kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic_bytes(args.config, main_run["my_solves"], wl),
        "peak_source": "measured live: FP64 pipe micro-benchmark of this library (max of the DFMA and DMMA m8n8k4 issue-bound loops); MEASURED_PEAKS.json has "
                       "no FP64 entry; nominal 37.2. ncu calibration of the two pipe counters at this peak: profiles/r02_fp64_pipe_calibration.md",
        "peak_dfma_measured": peak_dfma, "peak_dmma_measured": peak_dmma, "frac_of_nominal_37.2": achieved / 37.2,
        "flops_per_solve": F, "flops_per_solve_note": "ALGORITHMIC credit 8 n^3/3 + 8 n^2 nch (complex LU + nch right-hand sides, SURVEY 8(d)); the symmetric LDL^T "
                                                      "path executes about half of it in the trailing update: see executed_flops_per_solve",
        "executed_flops_per_solve_symmetric_path": W.executed_flops_per_solve_symmetric(n_sys, wl.nch),
        "frac_executed": main_run["my_solves"] * W.executed_flops_per_solve_symmetric(n_sys, wl.nch) / (k_ms * 1e-3) / 1e12 / peak,
        "solves_per_launch": main_run["my_solves"], "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
    }
    if isinstance(wl, Config2):
        t_obs = main_run["k2_ms"]
        ob = {"bound": "hbm", "ms": t_obs, "achieved": wl.observable_bytes() / (t_obs * 1e-3) / 1e9, "unit": "GB/s", "peak": peaks().get("hbm_gbs")}
        ob["frac"] = ob["achieved"] / ob["peak"] if ob["peak"] else None
        nl_sum = float(wl.Sbuf[1].to(torch.float64).sum().item())
        ob["fp64_tflops"] = 8.0 * wl.sweep.n_angles * (2 if wl.want_Ay else 1) * nl_sum / (t_obs * 1e-3) / 1e12
        ob["fp64_frac"] = ob["fp64_tflops"] / peak
        roofline["observables_kernel"] = ob

    cpu_baseline = None
    if not args.no_cpu_baseline:
        n = cpu_default(args)
        wl.cpu_run(n, warm=True)
        r = wl.cpu_run(n)
        cpu_baseline = {"value": r["solves_per_s"], "unit": "solves/s", "cores": r["procs"], "kind": "port", "sample": r["sample"],
                        "samples_per_s": r["samples_per_s"]}

    parity = None
    if args.config == 2 and world == 1 and not args.no_parity:
        from tools import parity_config2

        rep = parity_config2.compare(n_synthetic=1000, sweep=wl.sweep, model=wl.model)
        parity = {"passed": rep["passed"], "rows": rep["rows"], "workspaces": 64, "gates": rep["tolerances"],
                  "report": "profiles/r02_parity.json (same tool: tools/parity_config2.py)"}
        for mode in ("symmetric_path", "partial_pivoting_only"):
            a_ = rep[mode]["all"]
            parity[mode] = {"systems": a_["systems"], "nl_equal": a_["nl_equal"], "S_max": a_["S"]["max"], "S_p999": a_["S"]["p999"],
                            "dsdo_max": a_["dsdo"]["max"], "dsdo_p999": a_["dsdo"]["p999"], "dsdo_over_allowed_max": a_["dsdo_over_allowed_max"],
                            "sigma_rxn_max": a_["sigma_rxn"]["max"], "fallback_systems": a_["fallback"]["systems"],
                            "fallback_S_max": a_["fallback"]["S"]["max"], "S_outliers_refereed": rep[mode]["referee"]["S_refereed"],
                            "S_outliers_explained": rep[mode]["referee"]["S_explained"]}

    if args.config in (3, 4, 5) and world == 1 and not args.no_parity:
        # the bench's own inputs (leading rows / samples of the same seeded generators), every partial wave, against the oracle
        from tools import parity_general

        parity = {3: parity_general.config3, 4: parity_general.config4, 5: parity_general.config5}[args.config]()
        parity["tool"] = "tools/parity_general.py"

    pairs = wl.B * wl.nE * world
    line = {
        "metric": METRIC + {1: " (single KD03 sweep call)", 2: " (KDUQ elastic sweep)", 3: " (p+48Ca N=60 KDUQ sweep)", 4: " (nonlocal N=60)", 5: " (3x50 coupled)"}[args.config],
        "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": wl.describe(world), "samples_per_s": pairs / (ms_per_step * 1e-3), "solves_per_pair": main_run["all_solves"] / pairs,
        "status_nonzero": status_bad,
        "pivoting": {"scheme": "threshold pivoting on the diagonal (multipliers <= 64, symmetric LDL^T); a system that fails the test is re-solved by partial-pivoting LU",
                     "partial_pivoting_fallbacks_per_step_rank0": int(main_run["fallbacks"] // max(args.steps, 1)),
                     "fallback_fraction": float(main_run["fallbacks"]) / max(args.steps, 1) / max(main_run["my_solves"], 1)},
        "e2e": e2e, "gpu_launches": int(main_run["launches"]), "clocks": main_run["clocks"], "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    if world > 1:
        c = collective_desc
        c["ms_per_step"] = main_run["coll_ms"]
        if c.get("chunks", 1) == 1:  # not overlapped: the transfer time itself
            c["GBps_into_root"] = c["bytes_into_root"] / (main_run["coll_ms"] * 1e-3) / 1e9 if main_run["coll_ms"] > 0 else None
        line["collective"] = c
        line["strong"] = strong
    if parity is not None:
        line["parity"] = parity
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def traffic_bytes(config, solves, wl):
    """DRAM bytes of one launch of the dominant kernel, scaled from the committed ncu captures (profiles/r0*_traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))[f"config{config}"]
        units = solves if t["per"] == "solve" else wl.B * wl.nE
        return {"bytes": t["dram_bytes_per_unit"] * units, "per_" + t["per"]: t["dram_bytes_per_unit"],
                "algorithmic_bytes_per_" + t["per"]: t["algorithmic_bytes_per_unit"], "source": t["source"]}
    except Exception:
        return None


if __name__ == "__main__":
    main()
