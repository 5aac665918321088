#!/usr/bin/env python
"""Throughput of the quasi-elastic (p,n) DWBA workspace (jitr_b200.xs.quasielastic_pn) on the golden 48Ca(p,n) case:
B samples x (lmax + 1) partial waves x 2 j x 2 channels, every solve with wavefunction coefficients.
Synthetic depth variations of the golden potentials (timing only; parity is in tests/test_gpu_parity.py)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
sys.path.insert(0, "tests")
from bench_general import timeit  # noqa: E402
from test_gpu_parity import make_qe_ws  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    g = np.load("tests/golden/qe_48Ca_35MeV.npz")
    ws = make_qe_ws(g, tol=1e-16)  # no early break: every partial wave is solved
    rng = np.random.default_rng(3)
    U = np.repeat(g["U"][:1], B, axis=0).copy()
    U[:, 1] *= 1 + 0.05 * rng.standard_normal((B, 1))
    U[:, 3] *= 1 + 0.05 * rng.standard_normal((B, 1))
    Ut = [torch.tensor(np.ascontiguousarray(U[:, i]), device="cuda") for i in range(5)]
    ms = timeit(lambda: ws.xs_batch(*Ut))
    L1 = int(g["lmax"]) + 1
    solves = B * (2 * L1 - 1) * 2
    print(json.dumps({"workload": "48Ca(p,n)48Sc 35 MeV, N=40, lmax=18, 37 angles", "samples": B, "ms": ms,
                      "samples_per_s": B / ms * 1e3, "wavefunction_solves_per_s": solves / ms * 1e3}))


if __name__ == "__main__":
    main()
