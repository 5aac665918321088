"""Golden vectors for the Whitehead-Lim-Holt (WLH) optical model, from the REAL reference.

    python tests/golden/make_golden_wlh.py      (build container only; see make_golden.py)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_env  # noqa: E402

_ref_env.setup()

import numpy as np  # noqa: E402
from jitr.optical_potentials import wlh  # noqa: E402

from make_golden import elastic_case, save  # noqa: E402


def main():
    angles = np.linspace(0.1, np.pi, 180)
    mean_n = np.array(list(wlh.Global((1, 0)).params.values()))
    mean_p = np.array(list(wlh.Global((1, 1)).params.values()))
    sn = wlh.get_samples((1, 0))[:16]
    sp = wlh.get_samples((1, 1))[:16]
    save("wlh_params", mean_n=mean_n, mean_p=mean_p, samples_n=sn, samples_p=sp,
         names_n=np.array(wlh.get_param_names((1, 0))), names_p=np.array(wlh.get_param_names((1, 1))))
    # shape parameters for a grid of (projectile, target, Elab) incl. both branches of the surface term
    cases = [((1, 0), (208, 82), 14.0), ((1, 0), (40, 20), 55.0), ((1, 1), (120, 50), 12.0), ((1, 1), (48, 20), 45.0),
             ((1, 1), (208, 82), 30.0)]
    shp = []
    for prj, tgt, E in cases:
        for p in ([mean_n] + list(sn[:3]) if prj == (1, 0) else [mean_p] + list(sp[:3])):
            c, s, co = wlh.calculate_params(prj, tgt, E, *p)
            shp.append(list(c) + list(s) + list(co))
    save("wlh_shape_params", cases=np.array([[*prj, *tgt, E] for prj, tgt, E in cases]), shapes=np.array(shp, dtype=np.float64))
    # elastic workspaces: n+208Pb 14 MeV (surface term on), p+48Ca 45 MeV (surface term off, Coulomb)
    save("wlh_n208Pb_14MeV", **elastic_case((208, 82), (1, 0), 14.0, 40, 30, 13.0, angles, wlh.WLH((1, 0)), [mean_n] + list(sn[:6])))
    save("wlh_p48Ca_45MeV", **elastic_case((48, 20), (1, 1), 45.0, 40, 30, 12.0, angles, wlh.WLH((1, 1)), [mean_p] + list(sp[:6])))


if __name__ == "__main__":
    main()
