#!/usr/bin/env python
"""Three coupled channels (the numeric core of the reference's examples/coupled.py): p + 48Ca at 5 MeV with levels at
0 / 2.3 / 3.1 MeV, a real Woods-Saxon potential and surface-peaked couplings.  The S-matrix of a real, symmetric
interaction is unitary and symmetric; a batch of coupling strengths is solved in one call.

    python examples/coupled_channels.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import jitr_b200  # noqa: E402
from jitr_b200.reactions import ProjectileTargetSystem  # noqa: E402
from jitr_b200.utils import kinematics  # noqa: E402

N, nch = 50, 3
levels = np.array([0.0, 2.3, 3.1])
m_t, m_p = 44657.26581995028, 938.271653086152
Elab = 5.0
kin = kinematics.classical_kinematics(m_t, m_p, Elab, Zz=0)
Ecm, mu = kin.Ecm, kin.mu
E = Ecm - levels
k = np.sqrt(2 * mu * E) / 197.3269804
a = 6 * np.pi  # channel radius in units of 1/k0
solver = jitr_b200.Solver(N)
coupling = np.array([[0, 0.8, 0.1], [0.8, 0, 0.1], [0.1, 0.1, 0.0]])
system = ProjectileTargetSystem(channel_radius=a, lmax=0, mass_target=m_t, mass_projectile=m_p, Ztarget=0, Zproj=0,
                                coupling=lambda l: coupling, channel_levels=levels)
channels, asymptotics = system.get_partial_wave_channels(Elab, E, np.full(nch, mu), k, np.zeros(nch))
r = solver.radial_grid(a, k[0])
R0, a0, V0 = 4.0, 0.8, -42.0
ws = V0 / (1 + np.exp((r - R0) / a0))
spg = V0 * np.exp(-((r - R0) ** 2) / (2 * a0**2))


def potential(scale):
    V = np.zeros((nch, nch, N), dtype=np.complex128)
    for i in range(nch):
        V[i, i] = ws
        for j in range(nch):
            if i != j:
                V[i, j] = scale * coupling[i, j] * spg
    return V


# one system through the reference-shaped API ...
R, S, uext = solver.solve(channels[0], asymptotics[0], local_potential=potential(1.0))
print("S =\n", np.round(S, 6))
print("unitarity  |S^H S - 1| =", np.abs(S.conj().T @ S - np.eye(nch)).max(), "  symmetry |S - S^T| =", np.abs(S - S.T).max())

# ... and a batch of coupling strengths on the device
scales = np.linspace(0.5, 1.5, 1024)
Vb = torch.tensor(np.stack([potential(s) for s in scales]), device="cuda")
free = torch.tensor(solver.free_matrix(a, channels[0].l, channels[0].E, channels[0].mu), device="cuda")
asym = torch.tensor(np.stack([asymptotics[0].Hp, asymptotics[0].Hm, asymptotics[0].Hpp, asymptotics[0].Hmp]), device="cuda")
out = solver.solve_batch(a, float(channels[0].E[0]), float(channels[0].k[0]), nch, free, asym, local_potential=Vb)
S00 = out["S"][:, 0, 0].cpu().numpy()
print("elastic |S_00|^2 for coupling scale 0.5, 1.0, 1.5:", np.round(np.abs(S00[[0, 512, 1023]]) ** 2, 6))
