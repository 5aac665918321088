#!/usr/bin/env python
"""Throughput of the general batched solve (jitr_b200_solve_batched) on the shapes of BASELINE configs 4 and 5:
   config 4: nonlocal single channel, N = 60, full N x N complex kernel per system (57.6 KB from HBM)
   config 5: 3 coupled channels x N = 50 (n = 150), local couplings
Synthetic inputs (timing only; parity of these paths is in tests/test_gpu_parity.py).  Prints one JSON line."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import jitr_b200  # noqa: E402


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main5(B, rng, out, dev):
    """config 5 only, through the workloads generator (the real potential of examples/coupled.py)"""
    from jitr_b200 import workloads as W

    c = W.CONFIG5
    solver = jitr_b200.Solver(c["N"])
    free = torch.tensor(solver.free_matrix(c["a"], [0, 0, 0], list(c["E"]), [c["mu"]] * 3), device=dev)
    r = solver.radial_grid(c["a"], c["k"])
    Vt = torch.tensor(W.config5_potentials(r, B), device=dev)
    asym3 = torch.tensor(np.tile(np.array([[0.3 + 0.9j], [0.3 - 0.9j], [-0.8 + 0.2j], [-0.8 - 0.2j]]), (1, 3)), device=dev)
    ms = timeit(lambda: solver.solve_batch(c["a"], c["E"][0], c["k"], 3, free, asym3, local_potential=Vt))
    n = 150
    out["config5_real_potential"] = {"batch": B, "ms": ms, "solves_per_s": B / ms * 1e3, "tflops": B * (8 * n**3 / 3 + 8 * n**2 * 3) / ms / 1e9}
    print(json.dumps(out))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    import os

    from jitr_b200 import _cabi as _c

    if os.environ.get("JITR_B200_ASSUME_SYM"):
        _c.lib().jitr_b200_set_option(_c.OPT_ASSUME_SYMMETRIC_KERNEL, 1)
    only4 = bool(os.environ.get("ONLY4"))
    only5 = bool(os.environ.get("ONLY5"))
    rng = np.random.default_rng(0)
    out = {}
    dev = "cuda"
    asym1 = torch.tensor(np.array([[0.3 + 0.9j], [0.3 - 0.9j], [-0.8 + 0.2j], [-0.8 - 0.2j]]), device=dev)
    # ---- config 4
    if only5:
        return main5(B, rng, out, dev)
    N = 60
    solver = jitr_b200.Solver(N)
    a, E0, k0 = 14.5, 29.8, 1.2
    free = torch.tensor(solver.free_matrix(a, [0]), device=dev)
    r = solver.radial_grid(a, k0)
    base = -40.0 * np.exp(-((r[:, None] - r[None, :]) / 0.85) ** 2) / (1 + np.exp((0.5 * (r[:, None] + r[None, :]) - 7.0) / 0.65))
    U = torch.tensor(base[None] * (1.0 + 0.05 * rng.standard_normal((B, 1, 1))) * (1 + 0.3j), dtype=torch.complex128, device=dev)
    ms = timeit(lambda: solver.solve_batch(a, E0, k0, 1, free, asym1, nonlocal_potential=U))
    flops = 8 * N**3 / 3 + 8 * N**2
    out["config4_nonlocal_N60"] = {"batch": B, "ms": ms, "solves_per_s": B / ms * 1e3, "tflops": B * flops / ms / 1e9,
                                   "hbm_GBps": B * N * N * 16 / ms / 1e6}
    if only4:
        print(json.dumps(out))
        return
    # ---- config 5
    N, nch = 50, 3
    solver = jitr_b200.Solver(N)
    a = 6 * np.pi
    free = torch.tensor(solver.free_matrix(a, [0, 2, 2], [4.9, 2.6, 1.8], [919.0, 919.0, 919.0]), device=dev)
    r = solver.radial_grid(a, 0.48)
    V = np.zeros((B, nch, nch, N), dtype=np.complex128)
    ws = -42.0 / (1 + np.exp((r - 4.0) / 0.8))
    for i in range(nch):
        V[:, i, i] = ws[None] * (1 + 0.05 * rng.standard_normal((B, 1)))
        for j in range(i + 1, nch):
            c = 5.0 * np.exp(-((r - 4.0) / 1.2) ** 2)
            V[:, i, j] = V[:, j, i] = c[None] * (1 + 0.05 * rng.standard_normal((B, 1)))
    Vt = torch.tensor(V, device=dev)
    asym3 = torch.tensor(np.tile(np.array([[0.3 + 0.9j], [0.3 - 0.9j], [-0.8 + 0.2j], [-0.8 - 0.2j]]), (1, nch)), device=dev)
    ms = timeit(lambda: solver.solve_batch(a, 4.9, 0.48, nch, free, asym3, local_potential=Vt))
    n = N * nch
    flops = 8 * n**3 / 3 + 8 * n**2 * nch
    out["config5_coupled_3x50"] = {"batch": B, "ms": ms, "solves_per_s": B / ms * 1e3, "tflops": B * flops / ms / 1e9}
    ms = timeit(lambda: solver.solve_batch(a, 4.9, 0.48, nch, free, asym3, local_potential=Vt, wavefunction=True))
    out["config5_coupled_3x50_wavefunction"] = {"batch": B, "ms": ms, "solves_per_s": B / ms * 1e3}
    # optional per-phase clocks of the CTA-per-system kernel (library built with -DJITR_BLK_PROFILE)
    import ctypes

    from jitr_b200 import _cabi

    L = _cabi.lib()
    if hasattr(L, "jitr_b200_debug_block_phases"):
        buf = (ctypes.c_ulonglong * 16)()
        L.jitr_b200_debug_block_phases(buf, 1)
        Vc = Vt * (1 + 0.1j)  # absorptive: the symmetric path end to end
        solver.solve_batch(a, 4.9, 0.48, nch, free, asym3, local_potential=Vc)
        torch.cuda.synchronize()
        L.jitr_b200_debug_block_phases(buf, 0)
        tot = float(sum(buf)) or 1.0
        names = ["assemble", "panel", "panel_store", "upper", "trailing", "trailing_barrier", "finish", "-",
                 "col_update|sym_diag_block", "col_search_publish|sym_wait_diag", "col_barrier|sym_forward_store", "col_select|sym_barrier",
                 "sym_trailing", "sym_trailing_barrier", "-", "-"]
        out["config5_phase_share"] = {k: round(v / tot, 4) for k, v in zip(names, buf)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
