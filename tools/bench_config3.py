#!/usr/bin/env python
"""Throughput of the fused sweep on BASELINE config 3 (p+48Ca 30 MeV, N = 60, lmax = 40, KDUQ proton rows):
the generic solver mode (40 < N <= 64).  Timing only; parity is in tests/test_gpu_parity.py.  One JSON line."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import jitr_b200  # noqa: E402
from jitr_b200.optical_potentials import kduq  # noqa: E402
from jitr_b200.reactions import ElasticReaction  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    g = np.load("tests/golden/config3_p48Ca_30MeV.npz")
    rx = ElasticReaction((48, 20), (1, 1), mass_target=float(g["masses"][0]), mass_projectile=float(g["masses"][1]))
    ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(30.0), 12.0, jitr_b200.Solver(N), 40, g["angles"])
    rows = torch.tensor(np.tile(g["params"], (B // len(g["params"]) + 1, 1))[:B], device="cuda")
    model = kduq.KDUQ((1, 1))
    iw = ws.integral_workspace
    S, nl, st = iw.smatrix_params(model, rows)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        S, nl, st = iw.smatrix_params(model, rows)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    solves = int((2 * nl.to(torch.int64) - 1).sum().item())
    flops = 8 * N**3 / 3 + 8 * N**2
    print(json.dumps({"config": f"p+48Ca 30 MeV N={N} lmax=40", "samples": B, "ms": ms, "solves_per_s": solves / ms * 1e3,
                      "solves_per_sample": solves / B, "tflops": solves * flops / ms / 1e9}))


if __name__ == "__main__":
    main()
