#!/usr/bin/env python
"""Single-channel general solve (jitr_b200_solve_batched, user tensors of a full interaction matrix) over nbasis:
which kernel (warp-per-system / blocked CTA-per-system) wins where.  Select the library with JITR_B200_LIB."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import jitr_b200  # noqa: E402
from bench_general import timeit  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rng = np.random.default_rng(1)
    asym1 = torch.tensor(np.array([[0.3 + 0.9j], [0.3 - 0.9j], [-0.8 + 0.2j], [-0.8 - 0.2j]]), device="cuda")
    out = {}
    for N in (16, 24, 32, 40, 48, 56, 64, 80, 100):
        solver = jitr_b200.Solver(N)
        a, E0, k0 = 14.5, 29.8, 1.2
        free = torch.tensor(solver.free_matrix(a, [0]), device="cuda")
        V = torch.tensor(rng.standard_normal((B, N)) * (1 + 0.3j) - 40.0, dtype=torch.complex128, device="cuda")
        ms = timeit(lambda: solver.solve_batch(a, E0, k0, 1, free, asym1, local_potential=V))
        flops = 8 * N**3 / 3 + 8 * N**2
        out[f"N{N}"] = {"solves_per_s": round(B / ms * 1e3), "tflops": round(B * flops / ms / 1e9, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
