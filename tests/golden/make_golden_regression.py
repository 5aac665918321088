#!/usr/bin/env python
"""Goldens for the reference's external-code regression (tests/regression/test_regression.py:9-20, cases F1..F8:
Frescox proton / neutron elastic scattering on 78Ni, nbasis 50-100, lmax up to 100).  For every case the REAL reference
builds its workspace and mesh potentials with its own tests/regression/_builders.py; saved per case: kinematics, masses,
matching parameters, the mesh potentials, the Frescox angular distribution with its tolerance, and the reference's own
dsigma/dOmega, S-matrices and partial-wave count.  Build container only (imports /root/reference)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_env  # noqa: E402

_ref_env.setup()
sys.path.insert(0, "/root/reference")
from tests.regression._builders import build_case  # noqa: E402
from tests.regression._readers import load_case  # noqa: E402

manifest = json.load(open("/root/reference/tests/regression/manifest.json"))
out = {}
names = []
for case in manifest:
    if case["reference_code"] != "frescox":
        continue
    ref = load_case(case)
    built = build_case(ref)
    ws = built.workspace
    x = ws.xs(**built.xs_kwargs)
    sp, sm = ws.integral_workspace.smatrix(built.xs_kwargs["central_potential"], built.xs_kwargs["spin_orbit_potential"],
                                           built.xs_kwargs["coulomb_potential"])
    cid = ref.case_id.split("_")[0]
    names.append(ref.case_id)
    kin = ws.kinematics
    m = ref.metadata["matching"]
    rx = ws.reaction
    L1 = int(m["lmax"]) + 1
    pad = lambda v: np.concatenate([v, np.zeros(L1 - len(v), dtype=np.complex128)])  # noqa: E731
    co = built.xs_kwargs["coulomb_potential"]
    out.update({
        f"{cid}_kin": np.array([kin.Elab, kin.Ecm, kin.mu, kin.k, kin.eta]), f"{cid}_masses": np.array([rx.target.m0, rx.projectile.m0]),
        f"{cid}_AZ": np.array([rx.target.A, rx.target.Z, rx.projectile.A, rx.projectile.Z]),
        f"{cid}_match": np.array([float(m["channel_radius_fm"]), float(m["lmax"]), float(m["nbasis"])]),
        f"{cid}_angles": ref.theta_cm_rad, f"{cid}_frescox": ref.dsdo, f"{cid}_tol": np.array([ref.tolerance["rtol"], ref.tolerance["atol"]]),
        f"{cid}_central": built.xs_kwargs["central_potential"], f"{cid}_so": built.xs_kwargs["spin_orbit_potential"],
        f"{cid}_coul": np.zeros(0, dtype=np.complex128) if co is None else co,
        f"{cid}_dsdo": x.dsdo, f"{cid}_rxn": np.array(x.rxn), f"{cid}_nl": np.array(len(sp)), f"{cid}_splus": pad(sp), f"{cid}_sminus": pad(sm),
    })
    err = np.max(np.abs(x.dsdo - ref.dsdo) / (ref.tolerance["atol"] + ref.tolerance["rtol"] * np.abs(ref.dsdo)))
    print(ref.case_id, "N", m["nbasis"], "lmax", m["lmax"], "kept", len(sp), "angles", len(ref.theta_cm_rad), "reference vs frescox (units of tol):", err)
out["cases"] = np.array(names)
np.savez_compressed(os.path.join(HERE, "regression_frescox.npz"), **out)
