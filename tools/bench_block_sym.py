#!/usr/bin/env python
"""Config 5 (3 x 50 coupled) through the blocked kernel with the symmetric path on and off: throughput, fall-backs,
agreement of the two S-matrices."""
import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np, torch, jitr_b200
from jitr_b200 import _cabi
from bench_general import timeit
L = _cabi.lib()
B = 10000
rng = np.random.default_rng(0)
N, nch = 50, 3
solver = jitr_b200.Solver(N)
a = 6 * np.pi
free = torch.tensor(solver.free_matrix(a, [0, 2, 2], [4.9, 2.6, 1.8], [919.0, 919.0, 919.0]), device="cuda")
r = solver.radial_grid(a, 0.48)
V = np.zeros((B, nch, nch, N), dtype=np.complex128)
W = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0  # absorptive part (MeV): 0 = the real potential of examples/coupled.py
ws = -(42.0 + 1j * W) / (1 + np.exp((r - 4.0) / 0.8))
for i in range(nch):
    V[:, i, i] = ws[None] * (1 + 0.05 * rng.standard_normal((B, 1)))
    for j in range(i + 1, nch):
        c = 5.0 * np.exp(-((r - 4.0) / 1.2) ** 2)
        V[:, i, j] = V[:, j, i] = c[None] * (1 + 0.05 * rng.standard_normal((B, 1)))
Vt = torch.tensor(V, device="cuda")
asym3 = torch.tensor(np.tile(np.array([[0.3 + 0.9j], [0.3 - 0.9j], [-0.8 + 0.2j], [-0.8 - 0.2j]]), (1, nch)), device="cuda")
f0 = L.jitr_b200_sym_fallback_count()
for opt in (1, 0):
    L.jitr_b200_set_option(0, opt)
    ms = timeit(lambda: solver.solve_batch(a, 4.9, 0.48, nch, free, asym3, local_potential=Vt))
    print("symmetric option", opt, "ms", ms, "solves/s", B / ms * 1e3, "fallbacks so far", L.jitr_b200_sym_fallback_count() - f0)
L.jitr_b200_set_option(0, 1)
r1 = solver.solve_batch(a, 4.9, 0.48, nch, free, asym3, local_potential=Vt)
L.jitr_b200_set_option(0, 0)
r0 = solver.solve_batch(a, 4.9, 0.48, nch, free, asym3, local_potential=Vt)
print("max |S_sym - S_pp|", float((r1["S"] - r0["S"]).abs().max()))
