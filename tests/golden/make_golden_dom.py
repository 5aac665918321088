"""Golden vectors for the dispersive optical model (optical_potentials/dom.py), from the REAL reference.

    python tests/golden/make_golden_dom.py      (build container only; see make_golden.py)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_env  # noqa: E402

_ref_env.setup()

import numpy as np  # noqa: E402
from jitr import rmatrix  # noqa: E402
from jitr.optical_potentials import dom  # noqa: E402
from jitr.reactions import ElasticReaction  # noqa: E402
from jitr.xs.elastic import DifferentialWorkspace  # noqa: E402

from make_golden import asym_table, kin_array, save  # noqa: E402

BASE = dict(v0=52.0, v1=0.006, rv=1.25, av=0.65, wv0=12.0, wv1=24.0, rw=1.28, aw=0.62, ws0=15.0, ws1=12.5, ws2=0.021, rd=1.28,
            ad=0.55, vso0=6.0, wso0=0.35, wso1=25.0, rso=1.05, aso=0.59, rC=1.25)  # tests/test_dom_model.py:24-45


def rows(n, seed):
    rng = np.random.default_rng(seed)
    base = np.array([BASE[k] for k in dom.get_param_names()])
    p = base[None] * (1 + 0.05 * rng.standard_normal((n, base.size)))
    p[0] = base
    p[1, dom.get_param_names().index("ws2")] = 0.0  # the undamped branch of the surface dispersion integral
    return p


def main():
    names = dom.get_param_names()
    # ---- parameter map on a grid of (projectile, target, Ecm, Ef)
    cases = [((1, 0), (40, 20), 9.3, -9.1), ((1, 0), (208, 82), 29.8, -5.6), ((1, 1), (24, 12), 18.0, -6.0),
             ((1, 1), (208, 82), 7.0, -5.9), ((1, 0), (90, 40), -3.0 + 0.0, -8.0), ((1, 1), (48, 20), 44.0, -8.4)]
    P = rows(4, 1)
    shp = []
    for prj, tgt, Ecm, Ef in cases:
        for p in P:
            c, s, co = dom.calculate_params(prj, tgt, Ecm, Ef, *p)
            shp.append(list(c) + list(s) + list(co))
    r = np.linspace(0.05, 14.0, 57)
    c, s, co = dom.calculate_params((1, 1), (48, 20), 44.0, -8.4, *P[2])
    save("dom_params", names=np.array(names), params=P, cases=np.array([[*prj, *tgt, E, Ef] for prj, tgt, E, Ef in cases]),
         shapes=np.array(shp, dtype=np.float64), r=r, central=dom.central(r, *c), spin_orbit=dom.spin_orbit(r, *s),
         dVs=np.array([dom.delta_Vs_analytic(x, 15.0, 12.5, w2) for x in (-40.0, -3.0, 0.0, 0.7, 12.5, 80.0) for w2 in (0.0, 0.021)]))
    # ---- elastic workspaces (the reference computes the effective Fermi energy from its mass tables: recorded)
    angles = np.linspace(0.1, np.pi, 90)
    for name, tgt, prj, Elab, N, lmax, Rch in [("dom_n40Ca_14MeV", (40, 20), (1, 0), 14.0, 40, 20, 11.0),
                                                 ("dom_p208Pb_30MeV", (208, 82), (1, 1), 30.0, 40, 30, 13.0)]:
        rx = ElasticReaction(target=tgt, projectile=prj)
        kn = rx.kinematics(Elab)
        ws = DifferentialWorkspace.build_from_system(rx, kn, Rch, rmatrix.Solver(N), lmax, angles)
        rg = ws.radial_grid()
        model = dom.DOM()
        Pw = rows(5, 2)
        out = dict(dsdo=[], Ay=[], tot=[], rxn=[], nl=[], central=[], so=[], coul=[])
        splus = np.zeros((len(Pw), lmax + 1), dtype=np.complex128)
        sminus = np.zeros_like(splus)
        for i, p in enumerate(Pw):
            cc, so, coul = model(rg, rx, kn, *p)
            sp, sm = ws.integral_workspace.smatrix(cc, so, coul)
            x = ws.xs(cc, so, coul)
            splus[i, : len(sp)] = sp
            sminus[i, : len(sm)] = sm
            out["nl"].append(len(sp)); out["dsdo"].append(x.dsdo); out["Ay"].append(x.Ay); out["tot"].append(x.t); out["rxn"].append(x.rxn)
            out["central"].append(cc); out["so"].append(so); out["coul"].append(np.asarray(coul, dtype=np.float64) * np.ones(N))
        save(name, target=np.array(tgt), projectile=np.array(prj), masses=np.array([rx.target.m0, rx.projectile.m0]), Ef=np.float64(rx.Ef),
             kin=kin_array(kn), N=N, lmax=lmax, Rch=np.float64(Rch), angles=angles, rgrid=rg, asym=asym_table(ws), params=Pw,
             Splus=splus, Sminus=sminus, **{k: np.array(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
