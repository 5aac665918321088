"""The BASELINE.json workloads as synthetic-input generators shared by bench.py,
the parity tools and the tests.  Host-side numpy only: nothing here touches a device.

config 1  n+208Pb 30 MeV, N = 40, lmax = 30, R_ch = 12 fm, KD03 (one parameter row)
config 2  10^5 KDUQ samples x {n+40Ca, n+208Pb} x 32 energies 5-67 MeV, N = 40, lmax = 30, 180 angles
config 3  p+48Ca 30 MeV, N = 60, lmax = 40, R_ch = 12 fm, KDUQ proton posterior (Coulomb + spin-orbit)
config 4  n+208Pb 30 MeV, N = 60, Perey-Buck-like nonlocal kernel per (sample, l), l = 0..lmax, B samples of (V, W, beta)
config 5  3 channels x N = 50 (examples/coupled.py geometry), l = 0..10, B samples varying V and the couplings by 5 %
"""
from __future__ import annotations

import numpy as np

N_BASIS, LMAX, N_ANGLES = 40, 30, 180
TARGETS = ((40, 20, 37214.69), (208, 82, 193687.12278341982))  # (A, Z, rest mass MeV/c^2, AME2020)
MASS_N = 939.5654983938299
MASS_P = 938.271653086152
MASS_48CA = 44657.272049078
E_GRID = np.linspace(5.0, 67.0, 32)


def flops_per_solve(n, nch=1):
    """complex LU with partial pivoting + nch right-hand sides, 8 n^3 / 3 + 8 n^2 nch real flops."""
    return 8.0 * n**3 / 3.0 + 8.0 * n**2 * nch


def executed_flops_per_solve_symmetric(n, nch=1):
    """What the symmetric LDL^T path actually executes (lower triangle only, no U phase): about half the LU count."""
    return 4.0 * n**3 / 3.0 + 8.0 * n**2 * nch


# ------------------------------------------------------------------------------------------------ config 2
def synthetic_kduq_samples(n, block=0, projectile=(1, 0)):
    """rows 0..415: the KDUQ federal posterior; the rest: draws from a multivariate normal fitted to them
    (seed 20241019 + block), rejecting non-positive radii / diffusenesses.
    ``block`` > 0 (ranks > 0 of a multi-GPU run) gives an independent block of draws only."""
    from .optical_potentials import kduq

    real = kduq.get_samples(projectile)
    if block == 0 and n <= len(real):
        return real[:n].copy()
    vary = [i for i in range(real.shape[1]) if np.std(real[:, i]) > 0]
    mean = real[:, vary].mean(axis=0)
    cov = np.cov(real[:, vary], rowvar=False)
    rng = np.random.default_rng(20241019 + block)
    out = np.tile(real[0], (n, 1))
    filled = 0
    if block == 0:
        out[: len(real)] = real
        filled = len(real)
    posdef = bool(np.all(np.linalg.eigvalsh(cov) > 0))
    while filled < n:
        draw = rng.multivariate_normal(mean, cov, size=2 * (n - filled), method="cholesky", check_valid="ignore") \
            if posdef else rng.multivariate_normal(mean, cov, size=2 * (n - filled))
        cand = np.tile(real[0], (len(draw), 1))
        cand[:, vary] = draw
        ok = np.ones(len(cand), dtype=bool)
        for A in (40, 48, 208):  # positive radii / diffusenesses for every target used
            ok &= (cand[:, 12] - cand[:, 13] * A > 0.05) & (cand[:, 27] - cand[:, 28] * A > 0.05) & (cand[:, 36] > 0.05)
            ok &= (cand[:, 10] - cand[:, 11] * A ** (-1 / 3) > 0.3) & (cand[:, 25] - cand[:, 26] * A ** (1 / 3) > 0.3)
            ok &= cand[:, 34] - cand[:, 35] * A ** (-1 / 3) > 0.3
        cand = cand[ok][: n - filled]
        out[filled: filled + len(cand)] = cand
        filled += len(cand)
    return out


def energy_table():
    """[(target (A, Z), Elab, R_ch fm, kinematics tuple)] for the 64 (target, energy) workspaces of config 2."""
    from .utils import interaction_range
    from .utils.kinematics import semi_relativistic_kinematics

    tab = []
    for (A, Z, m) in TARGETS:
        for E in E_GRID:
            kin = semi_relativistic_kinematics(m, MASS_N, float(E), 0)
            Rch = (interaction_range(A) * kin.k + 2 * np.pi) / kin.k  # kduq_uq_demo.ipynb cell 5
            tab.append(((A, Z), float(E), float(Rch), tuple(kin)))
    return tab


def config2_workspaces(angles=None):
    """The 64 DifferentialWorkspaces of config 2 (host setup)."""
    import jitr_b200
    from .reactions import ElasticReaction

    angles = np.linspace(0.1, np.pi, N_ANGLES) if angles is None else angles
    solver = jitr_b200.Solver(N_BASIS)
    wss = []
    tab = energy_table()
    for (A, Z, m) in TARGETS:
        rx = ElasticReaction((A, Z), (1, 0), mass_target=m, mass_projectile=MASS_N)
        for (tz, E, Rch, kin) in [t for t in tab if t[0] == (A, Z)]:
            wss.append(jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(E), Rch, solver, LMAX, angles))
    return wss


# ------------------------------------------------------------------------------------------------ config 1 / 3
def config1_workspace(angles=None):
    import jitr_b200
    from .reactions import ElasticReaction

    angles = np.linspace(0.1, np.pi, N_ANGLES) if angles is None else angles
    rx = ElasticReaction((208, 82), (1, 0), mass_target=TARGETS[1][2], mass_projectile=MASS_N)
    return rx, jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(30.0), 12.0, jitr_b200.Solver(40), 30, angles)


def config3_workspace(angles=None):
    import jitr_b200
    from .reactions import ElasticReaction

    angles = np.linspace(0.1, np.pi, N_ANGLES) if angles is None else angles
    rx = ElasticReaction((48, 20), (1, 1), mass_target=MASS_48CA, mass_projectile=MASS_P)
    return rx, jitr_b200.DifferentialWorkspace.build_from_system(rx, rx.kinematics(30.0), 12.0, jitr_b200.Solver(60), 40, angles)


# ------------------------------------------------------------------------------------------------ config 4
CONFIG4 = dict(N=60, lmax=30, Rch=12.0, Elab=30.0, R=1.25 * 208 ** (1 / 3), a=0.65, seed=4)


def config4_samples(B, seed=4):
    """(V, W, beta) ~ N(71, 3), N(15, 2), N(0.85, 0.03)   (SURVEY 8(d) config 4)."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.normal(71.0, 3.0, B), rng.normal(15.0, 2.0, B), rng.normal(0.85, 0.03, B)], axis=1)


def perey_buck_kernel(rg, rpg, l, V, W, beta, R=CONFIG4["R"], a=CONFIG4["a"]):
    """U_l(r, r') = -WS((r + r')/2; V + iW, R, a) H_l(r, r'; beta) on the meshgrid (rg, rpg): the partial-wave
    projection of a Gaussian nonlocality, built with the exponentially scaled Bessel function (the stable form of
    potential_forms.py:14-19; the definition of tests/golden/make_golden.py "config 4"):
        H_l = 2 z sqrt(pi / 2z) ive(l + 1/2, z) exp(-(r - r')^2 / beta^2) / (beta sqrt(pi)),  z = 2 r r' / beta^2."""
    from scipy.special import ive

    z = 2.0 * rg * rpg / beta**2
    H = 2.0 * z * np.sqrt(np.pi / (2.0 * z)) * ive(l + 0.5, z) * np.exp(-((rg - rpg) ** 2) / beta**2) / (beta * np.sqrt(np.pi))
    x = (0.5 * (rg + rpg) - R) / a
    ws = np.where(x <= np.log(1e16), 1.0 / (1.0 + np.exp(np.minimum(x, 40.0))), 0.0)
    return -(V + 1j * W) * ws * H


# ------------------------------------------------------------------------------------------------ config 5
CONFIG5 = dict(N=50, nch=3, lmax=10, levels=(0.0, 2.3, 3.1), a=6 * np.pi, Elab=5.0, V=-42.0, R=4.0, diff=0.8, seed=5,
               E=(4.897109281203547, 2.597109281203547, 1.797109281203547), k=0.480781357063321, mu=918.9637641236778,
               eta=1.413707698547276, coupling=((0.0, 0.8, 0.1), (0.8, 0.0, 0.1), (0.1, 0.1, 0.0)))
ALPHA, HBARC = 1.0 / 137.0359991, 197.3269804  # utils/constants.py:5-6


def config5_potentials(r, B, seed=5, absorptive=False):
    """(B, 3, 3, N) local potentials of examples/coupled.py:14-100 (real Woods-Saxon V = -42, R = 4, a = 0.8 + Coulomb
    of a charged sphere on the diagonal, coupling_matrix x surface-peaked Gaussian off it), V and the coupling
    strengths varied by +-5 % per sample.  Sample 0 is the example itself."""
    rng = np.random.default_rng(seed)
    nch, N = 3, len(r)
    V0, R0, a0 = CONFIG5["V"], CONFIG5["R"], CONFIG5["diff"]
    ws = V0 / (1.0 + np.exp((r - R0) / a0))
    coul = ALPHA * HBARC * np.where(r <= R0, (3.0 - (r / R0) ** 2) / (2.0 * R0), 1.0 / r)
    spg = V0 * np.exp(-((r - R0) ** 2) / (2 * np.pi * a0) ** 2)
    cm = np.array(CONFIG5["coupling"])
    fV = 1.0 + 0.05 * rng.uniform(-1, 1, (B, 1))
    fC = 1.0 + 0.05 * rng.uniform(-1, 1, (B, nch, nch, 1))
    fC = 0.5 * (fC + fC.transpose(0, 2, 1, 3))
    fV[0] = 1.0
    fC[0] = 1.0
    V = (cm[None, :, :, None] * spg[None, None, None, :] * fC).astype(np.complex128)
    for i in range(nch):
        V[:, i, i] += (ws[None] * fV) * ((1 + 0.1j) if absorptive else 1.0) + coul[None]
    return V
