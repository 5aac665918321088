#!/usr/bin/env python
"""Parity of the benchmarked workloads of BASELINE configs 3, 4 and 5 against the CPU oracle, on the bench's OWN inputs
(`bench.py --config N` calls these for its ``parity`` record; tests/test_gpu_parity.py runs them at a smaller size).

    config 3  p+48Ca 30 MeV, N = 60, lmax = 40: `rows` of the bench's KDUQ proton block -> nl exact, S 1e-10, dsigma/dOmega and A_y 1e-8
              (dsigma/dOmega: or the S tolerance propagated through the partial-wave sum, tools/parity_config2.dsdo_condition)
    config 4  nonlocal N = 60: kernels generated on the device from (V, W, beta) rows, solved by the packed bulk-async kernel, against
              scipy's exponentially scaled Bessel function + the oracle's explicit-inverse solve, every l = 0..lmax
    config 5  3 x 50 coupled, the real potential of examples/coupled.py, every l = 0..lmax
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RTOL_S, RTOL_XS = 1e-10, 1e-8


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300)))


def config3(n_rows=96, procs=None):
    import torch

    import jitr_b200
    from jitr_b200 import workloads as W
    from jitr_b200.optical_potentials import kduq
    from jitr_b200.utils.kinematics import semi_relativistic_kinematics
    from oracle import parity_sweep
    from tools import parity_config2 as pc

    rows = W.synthetic_kduq_samples(416 + n_rows, projectile=(1, 1))
    pick = np.concatenate([np.arange(min(48, n_rows)), 416 + np.arange(max(0, n_rows - 48))])  # federal rows + synthetic rows
    rows = rows[pick]
    rx, ws = W.config3_workspace()
    sweep = jitr_b200.ElasticSweep([ws])
    p = np.zeros((len(rows), 40))
    p[:, : rows.shape[1]] = rows
    S, nl, status = sweep.smatrix_params(kduq.KDUQ((1, 1)), torch.tensor(p, device="cuda"))
    x = sweep.observables(S, nl)
    kin = semi_relativistic_kinematics(W.MASS_48CA, W.MASS_P, 30.0, 20)
    tab = [((48, 20), 30.0, 12.0, tuple(kin))]
    o = parity_sweep.run(tab, rows, 60, 40, projectile=(1, 1), procs=procs)
    g = dict(S=S.cpu().numpy(), nl=nl.cpu().numpy(), status=status.cpu().numpy(), dsdo=x.dsdo.cpu().numpy(), rxn=x.rxn.cpu().numpy(),
             tot=x.t.cpu().numpy(), mask=np.zeros(nl.shape, dtype=np.uint64))
    kappa = pc.dsdo_condition(o, pc.sweep_tables(sweep))
    r = pc.compare_one(g, o, kappa)
    ok = (r["nl_equal"] and r["zeros_beyond_nl"] and r["status_nonzero"] == 0 and r["S"]["max"] < RTOL_S and r["dsdo_over_allowed_max"] <= 1.0
          and r["sigma_rxn"]["max"] < RTOL_XS)
    return {"passed": bool(ok), "rows": int(len(rows)), "systems": r["systems"], "nl_equal": r["nl_equal"], "S_max": r["S"]["max"],
            "dsdo_max": r["dsdo"]["max"], "dsdo_over_allowed_max": r["dsdo_over_allowed_max"], "sigma_rxn_max": r["sigma_rxn"]["max"]}


def config4(n_samples=24):
    import torch

    import jitr_b200
    from jitr_b200 import workloads as W
    from jitr_b200.optical_potentials import nonlocal_forms
    from jitr_b200.reactions import ProjectileTargetSystem
    from jitr_b200.utils.kinematics import semi_relativistic_kinematics
    from oracle import jitr_oracle as orc

    c = W.CONFIG4
    N, lmax = c["N"], c["lmax"]
    kin = semi_relativistic_kinematics(W.TARGETS[1][2], W.MASS_N, c["Elab"], 0)
    a = c["Rch"] * kin.k
    solver = jitr_b200.Solver(N)
    sysm = ProjectileTargetSystem(a, lmax, Ztarget=82, Zproj=0)
    channels, asym = sysm.get_partial_wave_channels(*kin)
    smp = W.config4_samples(n_samples, seed=c["seed"])
    r = solver.radial_grid(a, kin.k)
    U = nonlocal_forms.perey_buck_kernels(r, lmax, smp[:, 0], smp[:, 1], smp[:, 2], R=c["R"], a=c["a"], l_major=True)
    L1 = lmax + 1
    free = torch.tensor(np.stack([solver.free_matrix(a, [l]) for l in range(L1)]), device="cuda")
    asym_t = torch.tensor(np.stack([np.array([[asym[l].Hp[0]], [asym[l].Hm[0]], [asym[l].Hpp[0]], [asym[l].Hmp[0]]]) for l in range(L1)]), device="cuda")
    idx = torch.arange(L1, dtype=torch.int32, device="cuda").repeat_interleave(n_samples)
    out = solver.solve_batch(a, kin.Ecm, kin.k, 1, free, asym_t[idx.long()].contiguous(), nonlocal_potential=U.reshape(L1 * n_samples, N, N), free_index=idx)
    Sg = out["S"].reshape(L1, n_samples).cpu().numpy()
    rg, rpg = solver.nonlocal_radial_grids(a, kin.k)
    worst = 0.0
    b = orc.boundary_vector(N)
    for l in range(L1):
        F = orc.free_matrix(N, a, [l])
        H = [np.array([complex(v)]) for v in (asym[l].Hp[0], asym[l].Hm[0], asym[l].Hpp[0], asym[l].Hmp[0])]
        for s in range(n_samples):
            Uo = W.perey_buck_kernel(rg, rpg, l, *smp[s])
            V = orc.interaction_matrix(N, kin.k, kin.Ecm, a, 1, nonlocal_potential=Uo)
            _, S, _, _ = orc.solve_smatrix(F + V, b, *H, np.ones(1), a, 1, N)
            worst = max(worst, abs(Sg[l, s] - S[0, 0]) / abs(S[0, 0]))
    return {"passed": bool(worst < RTOL_S and int(out["status"].abs().sum()) == 0), "samples": n_samples, "systems": n_samples * L1, "S_max": float(worst)}


def config5(n_samples=12):
    import torch

    import jitr_b200
    from jitr_b200 import workloads as W
    from jitr_b200.reactions import ProjectileTargetSystem
    from oracle import jitr_oracle as orc

    c = W.CONFIG5
    N, lmax, nch = c["N"], c["lmax"], 3
    solver = jitr_b200.Solver(N)
    cm = np.array(c["coupling"])
    sysc = ProjectileTargetSystem(channel_radius=c["a"], lmax=lmax, Ztarget=20, Zproj=1, channel_levels=np.array(c["levels"]), coupling=lambda l: cm)
    channels, asym = sysc.get_partial_wave_channels(c["Elab"], np.array(c["E"]), np.full(3, c["mu"]), np.full(3, c["k"]), np.full(3, c["eta"]))
    L1 = lmax + 1
    r = solver.radial_grid(c["a"], c["k"])
    V = W.config5_potentials(r, n_samples, seed=c["seed"])
    free = torch.tensor(np.stack([solver.free_matrix(c["a"], [l] * 3, list(c["E"]), [c["mu"]] * 3) for l in range(L1)]), device="cuda")
    asym_t = torch.tensor(np.stack([np.array([asym[l].Hp, asym[l].Hm, asym[l].Hpp, asym[l].Hmp]) for l in range(L1)]), device="cuda")
    idx = torch.arange(L1, dtype=torch.int32, device="cuda").repeat_interleave(n_samples)
    Vt = torch.tensor(V, device="cuda").unsqueeze(0).expand(L1, n_samples, 3, 3, N).reshape(L1 * n_samples, 3, 3, N)
    out = solver.solve_batch(c["a"], c["E"][0], c["k"], 3, free, asym_t[idx.long()].contiguous(), local_potential=Vt, free_index=idx)
    Sg = out["S"].reshape(L1, n_samples, 3, 3).cpu().numpy()
    worst_abs, worst_unit = 0.0, 0.0
    b = orc.boundary_vector(N)
    w = np.zeros(nch)
    w[0] = 1
    for l in range(L1):
        F = orc.free_matrix(N, c["a"], [l] * 3, list(c["E"]), [c["mu"]] * 3)
        H = [np.ascontiguousarray(np.asarray(v, dtype=np.complex128)) for v in (asym[l].Hp, asym[l].Hm, asym[l].Hpp, asym[l].Hmp)]
        for s in range(n_samples):
            Vm = orc.interaction_matrix(N, c["k"], c["E"][0], c["a"], nch, local_potential=V[s])
            _, S, _, _ = orc.solve_smatrix(F + Vm, b, *H, w, c["a"], nch, N)
            worst_abs = max(worst_abs, float(np.max(np.abs(Sg[l, s] - S))))
            worst_unit = max(worst_unit, float(np.max(np.abs(Sg[l, s].conj().T @ Sg[l, s] - np.eye(3)))))
    # elements of a unitary 3 x 3 S-matrix: |S_ij| <= 1, compared in absolute terms (= relative to the matrix norm)
    return {"passed": bool(worst_abs < RTOL_S and int(out["status"].abs().sum()) == 0), "samples": n_samples, "systems": n_samples * L1,
            "S_max_abs": worst_abs, "unitarity_defect": worst_unit}


if __name__ == "__main__":
    import json

    which = [int(v) for v in sys.argv[1:]] or [3, 4, 5]
    print(json.dumps({f"config{c}": {3: config3, 4: config4, 5: config5}[c]() for c in which}, indent=1))
