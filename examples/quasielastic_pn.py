#!/usr/bin/env python
"""48Ca(p,n)48Sc quasi-elastic charge exchange in the DWBA (the reference's xs/quasielastic_pn.py workflow): distorted
waves of the entrance and exit channel, isovector transition potential, differential cross section -- for one
potential through the reference-shaped API and for a batch of depth variations on the device.

    python examples/quasielastic_pn.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import jitr_b200  # noqa: E402
from jitr_b200.reactions import Reaction  # noqa: E402
from jitr_b200.xs.quasielastic_pn import Workspace  # noqa: E402

rx = Reaction((48, 20), (1, 1), (1, 0), (48, 21), mass_target=44657.26581995028, mass_projectile=938.271653086152,
              mass_product=939.5654205, mass_residual=44656.47568732)
ke = rx.kinematics(35.0)
kx = rx.kinematics_exit(ke, residual_excitation_energy=6.67)  # isobaric analogue state
angles = np.linspace(0.1, np.pi - 0.1, 37)
ws = Workspace(rx, ke, kx, jitr_b200.Solver(40), angles, lmax=18, channel_radius_fm=12.0)
r = ws.radial_grid()
A, Z = 48, 20
R0 = 1.2 * A ** (1 / 3)
f = 1 / (1 + np.exp((r - R0) / 0.65))
coul = np.where(r < 1.25 * A ** (1 / 3), Z * 1.43996 * (3 - (r / (1.25 * A ** (1 / 3))) ** 2) / (2 * 1.25 * A ** (1 / 3)), Z * 1.43996 / r)
Up = -(52.0 + 2.0j) * f
Un = -(44.0 + 2.5j) * f

xs = ws.xs(coul, Up, U_n_central=Un)
print("dsigma/dOmega (mb/sr) at", np.round(np.degrees(angles[[0, 9, 18, 36]]), 1), "deg:", xs[[0, 9, 18, 36]])

B = 512
rng = np.random.default_rng(1)
scale_p, scale_n = 1 + 0.03 * rng.standard_normal((B, 1)), 1 + 0.03 * rng.standard_normal((B, 1))
t = lambda v: torch.tensor(v, dtype=torch.complex128, device="cuda")  # noqa: E731
xsb = ws.xs_batch(t(np.tile(coul, (B, 1))), t(Up[None] * scale_p), None, t(Un[None] * scale_n), None)
print(f"{B} depth variations: forward-angle dsigma/dOmega = {float(xsb[:, 0].mean()):.4f} +- {float(xsb[:, 0].std()):.4f} mb/sr")
