"""Golden vectors for the quasi-elastic (p,n) DWBA workspace (xs/quasielastic_pn.py), from the REAL reference.

    python tests/golden/make_golden_qe.py      (build container only; see make_golden.py)

Two workspaces: the reference's own unit-test configuration (tests/test_xs_optional_spin_orbit.py:56-103) and a
48Ca(p,n)48Sc case at 35 MeV with Woods-Saxon + Thomas spin-orbit + charged-sphere Coulomb potentials for a few
depth variations, with the default early break and with it disabled.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_env  # noqa: E402

_ref_env.setup()

import numpy as np  # noqa: E402
from jitr.optical_potentials import potential_forms as pf  # noqa: E402
from jitr.reactions import Reaction  # noqa: E402
from jitr.rmatrix import Solver  # noqa: E402
from jitr.xs.quasielastic_pn import Workspace  # noqa: E402

from make_golden import kin_array, save  # noqa: E402


def case(reaction, Elab, Ex, N, lmax, Rch, angles, pots, tol):
    ke = reaction.kinematics(Elab)
    kx = reaction.kinematics_exit(ke, Ex)
    ws = Workspace(reaction=reaction, kinematics_entrance=ke, kinematics_exit=kx, solver=Solver(N), angles=angles,
                   lmax=lmax, channel_radius_fm=Rch, tmatrix_abs_tol=tol)
    r = ws.radial_grid()
    T, Sn, Sp, xs, U = [], [], [], [], []
    for make in pots:
        u = make(r)  # (U_p_coulomb, U_p_central, U_p_so or None, U_n_central, U_n_so or None)
        t, sn, sp = ws.tmatrix(*u)
        T.append(t); Sn.append(sn); Sp.append(sp)
        xs.append(ws.xs(*u))
        U.append([np.zeros(N, dtype=np.complex128) if v is None else np.asarray(v, dtype=np.complex128) * np.ones(N) for v in u])
    return dict(
        masses=np.array([reaction.target.m0, reaction.projectile.m0, reaction.product.m0, reaction.residual.m0]),
        AZ=np.array([reaction.target.A, reaction.target.Z, reaction.projectile.A, reaction.projectile.Z,
                     reaction.product.A, reaction.product.Z, reaction.residual.A, reaction.residual.Z]),
        Q=np.float64(reaction.Q), Elab=np.float64(Elab), Ex=np.float64(Ex), N=N, lmax=lmax, Rch=np.float64(Rch),
        tol=np.float64(tol), angles=angles, kin_entrance=kin_array(ke), kin_exit=kin_array(kx), rgrid=r,
        isovector_factor=np.float64(ws.isovector_factor), xs_factor=np.float64(ws.xs_factor),
        geometric_factor=ws.geometric_factor, U=np.array(U), Tpn=np.array(T), Sn=np.array(Sn), Sp=np.array(Sp),
        xs=np.array(xs), has_so=np.array([u is not None for u in pots[0](r)])[[2, 4]],
    )


def gaussian(r, strength, radius):
    return np.asarray(strength * np.exp(-((r / radius) ** 2)), dtype=np.complex128)


def main():
    # --- the reference's unit test
    rx = Reaction((48, 20), (1, 1), (1, 0), (48, 21))
    pots = [lambda r: (gaussian(r, 2.0, 4.0), gaussian(r, -30.0 - 1.0j, 3.2), None, gaussian(r, -22.0 - 1.5j, 2.8), None)]
    save("qe_reftest", **case(rx, 20.0, 2.0, 14, 3, 8.0, np.linspace(0.15, np.pi - 0.15, 5), pots, 1e-6))
    # --- 48Ca(p,n)48Sc to the isobaric analogue state region, 35 MeV
    A, Z = 48, 20
    R0, RC = 1.2 * A ** (1 / 3), 1.25 * A ** (1 / 3)

    def ws_pots(Vp, Wp, Vn, Wn, Vso):
        def make(r):
            coul = pf.coulomb_charged_sphere(r, Z * 1, RC)
            up = -(Vp + 1j * Wp) * pf.woods_saxon_safe(r, R0, 0.65) + 4j * 3.0 * pf.woods_saxon_prime_safe(r, 1.05 * R0, 0.6)
            un = -(Vn + 1j * Wn) * pf.woods_saxon_safe(r, R0, 0.65) + 4j * 4.0 * pf.woods_saxon_prime_safe(r, 1.05 * R0, 0.6)
            so_p = Vso * pf.thomas_safe(r, 0.95 * R0, 0.6) * 2.0
            so_n = (Vso - 0.4) * pf.thomas_safe(r, 0.95 * R0, 0.6) * 2.0
            return coul, up, so_p, un, so_n
        return make

    samples = [(52.0, 2.0, 44.0, 2.5, 5.8), (55.0, 1.0, 41.0, 3.5, 6.2), (49.0, 3.0, 46.0, 1.5, 5.0),
               (53.5, 2.2, 42.5, 2.8, 6.6), (50.5, 2.7, 45.0, 2.1, 5.4)]
    pots = [ws_pots(*s) for s in samples]
    angles = np.linspace(0.1, np.pi - 0.1, 37)
    save("qe_48Ca_35MeV", **case(rx, 35.0, 6.67, 40, 18, 12.0, angles, pots, 1e-6))
    save("qe_48Ca_35MeV_nobreak", **case(rx, 35.0, 6.67, 40, 18, 12.0, angles, pots[:2], 1e-16))


if __name__ == "__main__":
    main()
