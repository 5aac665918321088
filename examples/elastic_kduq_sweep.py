#!/usr/bin/env python
"""KDUQ uncertainty sweep of n + 208Pb elastic scattering (the pattern of the reference's kduq_uq_demo notebook):
one call evaluates the optical potentials, solves every (sample, l, j) R-matrix system and forms dsigma/dOmega,
A_y and sigma_rxn for a whole batch of posterior samples on the GPU.

    python examples/elastic_kduq_sweep.py [n_samples]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import jitr_b200  # noqa: E402
from jitr_b200.optical_potentials import kduq  # noqa: E402
from jitr_b200.reactions import ElasticReaction  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rx = ElasticReaction((208, 82), (1, 0), mass_target=193687.12278341982, mass_projectile=939.5654983938299)
kin = rx.kinematics(30.0)
angles = np.linspace(0.1, np.pi, 180)
ws = jitr_b200.DifferentialWorkspace.build_from_system(rx, kin, 12.0, jitr_b200.Solver(40), 30, angles)

# synthetic posterior: the KD03 central values with a 2 % Gaussian spread on the varying parameters
rng = np.random.default_rng(0)
central = kduq.get_kd03((1, 0))
rows = np.tile(central, (B, 1)) * (1 + 0.02 * rng.standard_normal((B, central.size)))
rows[:, -2:] = central[-2:]  # Fermi-energy parameters stay fixed (kduq.py:231-237)
xs = ws.xs_params(kduq.KDUQ((1, 0)), torch.tensor(rows, device="cuda"))

# uncertainty bands on the device (= np.percentile(dsdo, [16, 50, 84], axis=0)): only 3 x 180 numbers leave the GPU
lo, med, hi = jitr_b200.uq.percentiles(xs.dsdo, [16, 50, 84]).cpu().numpy()
print(f"{B} samples; sigma_rxn = {float(xs.rxn.mean()):.1f} +- {float(xs.rxn.std()):.1f} mb")
for i in (0, 45, 90, 179):
    print(f"theta = {np.degrees(angles[i]):6.1f} deg   dsigma/dOmega = {med[i]:12.4e}  [{lo[i]:.4e}, {hi[i]:.4e}] mb/sr")
