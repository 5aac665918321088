#!/usr/bin/env python
"""Stress the work hand-out / lock-step termination of the sweep kernel: many launches with awkward pair counts
(around multiples of the warps per CTA and of the grid), results checked against one large launch."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import jitr_b200  # noqa: E402
from jitr_b200.optical_potentials import kduq  # noqa: E402
from jitr_b200.reactions import ElasticReaction  # noqa: E402

g = np.load("tests/golden/kduq_params.npz")
rows = np.zeros((4000, 40))
rows[:, :37] = np.tile(g["federal_n"], (10, 1))[:4000]
rx = ElasticReaction((208, 82), (1, 0), mass_target=193687.12278341982, mass_projectile=939.5654983938299)
solver = jitr_b200.Solver(40)
ang = np.linspace(0.1, np.pi, 19)
wss = [jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(E), 12.0, solver, 30, ang) for E in (7.0, 30.0, 61.0)]
sweep = jitr_b200.ElasticSweep(wss)
model = kduq.KDUQ((1, 0))
full = torch.tensor(rows, device="cuda")
Sref, nlref, _ = sweep.smatrix_params(model, full)
torch.cuda.synchronize()
n = 0
for rep in range(3):
    for B in (1, 2, 3, 4, 5, 11, 12, 13, 47, 48, 49, 147, 148, 149, 591, 592, 593, 1775, 1776, 1777, 3999):
        S, nl, st = sweep.smatrix_params(model, full[:B])
        torch.cuda.synchronize()
        assert torch.equal(nl, nlref[:B]) and torch.equal(S, Sref[:B]), f"mismatch at B={B}"
        n += 1
print(f"stress ok: {n} launches, results identical to the reference launch")
