"""Local dispersive optical model (DOM) -- host mirror of the reference's ``optical_potentials/dom.py``.

The energy dependence of the DOM depths (dom.py:151-283) includes the analytic surface dispersion integral
(:399-455, exponential integrals of real and complex argument), so the map *parameters -> shape numbers* stays on
the host (vectorised numpy, one row per (sample, energy)); the radial forms (:63-128: Woods-Saxon volume terms at two
geometries, surface derivative term, Thomas spin-orbit, charged-sphere Coulomb) are evaluated on the mesh by the same
device code as every other built-in model, through the per-pair canonical rows ``JITR_B200_MODEL_SHAPE_PAIR``:

    Vv Rv av Wv Rw aw Vd Wd Rd ad Vso Rso aso Wso Rwso awso RC Vw
    V  R0 av Wv Rw aw dVs Ws Rd ad 2(Vso+dVso) Rso aso 2 Wso Rso aso RC dVv
"""
from __future__ import annotations

import numpy as np
from scipy.special import exp1, expi

from .. import _cabi
from ..constants import ALPHA, HBARC
from .omp import SingleChannelOpticalModel

PARAM_NAMES = ("v0", "v1", "rv", "av", "wv0", "wv1", "rw", "aw", "ws0", "ws1", "ws2", "rd", "ad", "vso0", "wso0", "wso1",
               "rso", "aso", "rC")


def get_param_names():
    """dom.py:53-60."""
    return list(PARAM_NAMES)


# ---- energy dependence of the depths (dom.py:151-283); all accept arrays
def V_depth(dE, V0, v01):
    return V0 * np.exp(-v01 * dE)


def Vso_depth(dE, Vso, v01):
    return Vso * np.exp(-v01 * dE)


def Wv_depth(dE, Wv0, wv1):
    return Wv0 * dE**2 / (dE**2 + wv1**2)


def Wso_depth(dE, Wso, wso1):
    return Wso * dE**2 / (dE**2 + wso1**2)


def Ws_depth(dE, Ws0, ws1, ws2):
    return Ws0 * dE**4 / (dE**4 + ws1**4) * np.exp(-ws2 * np.abs(dE))


def Delta_Vv_depth(dE, Wv0, wv1):
    return Wv0 * wv1 * dE / (dE**2 + wv1**2)


def Delta_Vso_depth(dE, Wso, wso1):
    return Wso * wso1 * dE / (dE**2 + wso1**2)


def delta_Vs_analytic(delta_E, Ws0, ws1, ws2):
    """Analytic surface-dispersion correction (dom.py:399-455), broadcasting over arrays of all four arguments."""
    dE, Ws0, ws1, ws2 = np.broadcast_arrays(*(np.asarray(v, dtype=np.float64) for v in (delta_E, Ws0, ws1, ws2)))
    scalar = dE.ndim == 0
    dE, Ws0, ws1, ws2 = (np.atleast_1d(v).astype(np.float64) for v in (dE, Ws0, ws1, ws2))
    out = np.zeros_like(dE)
    nz = dE != 0.0
    if np.any(nz):
        x, w0, s1, s2 = dE[nz], Ws0[nz], ws1[nz], ws2[nz]
        ax = np.abs(x)
        damp = s2 != 0.0
        J1 = np.zeros_like(x)
        z1 = s2[damp] * ax[damp]
        J1[damp] = (-np.exp(-z1) * expi(z1) - np.exp(z1) * exp1(z1)) / (2.0 * ax[damp])
        J2 = np.pi * (s1**2 + x**2) / (2.0 * np.sqrt(2.0) * s1**3)  # ws2 = 0 branch
        if np.any(damp):
            roots = s1[damp][:, None] * np.array([np.exp(1j * np.pi / 4), np.exp(1j * 3 * np.pi / 4)])[None, :]
            F = np.exp(-s2[damp][:, None] * roots) * exp1(-s2[damp][:, None] * roots)
            contrib = (roots**2 + x[damp][:, None] ** 2) / (4.0 * roots**3) * F
            J2[damp] = 2.0 * np.real(np.sum(contrib, axis=1))
        pre1 = x**4 / (x**4 + s1**4)
        pre2 = s1**4 / (x**4 + s1**4)
        out[nz] = (2.0 * w0 * x / np.pi) * (pre1 * J1 + pre2 * J2)
    return out.item() if scalar else out


def Delta_Vs_depth(dE, Ws0, ws1, ws2):
    return delta_Vs_analytic(dE, Ws0, ws1, ws2)


def coulomb_correction(Zz, RC):
    """dom.py:385-396."""
    return 6.0 * Zz * ALPHA * HBARC / (5 * RC)


def calculate_params(projectile, target, Ecm, Ef, v0, v1, rv, av, wv0, wv1, rw, aw, ws0, ws1, ws2, rd, ad, vso0, wso0, wso1,
                     rso, aso, rC):
    """(central_params, spin_orbit_params, coulomb_params) of one parameter set (dom.py:287-382)."""
    A, Z = target
    Ap, Zp = projectile
    assert Ap == 1 and Zp in (0, 1)
    A13 = A ** (1 / 3)
    R0, Rw, Rd, Rso, RC = rv * A13, rw * A13, rd * A13, rso * A13, rC * A13
    Ec = coulomb_correction(Z * Zp, RC) if Zp == 1 else 0.0
    dE = Ecm - Ec - Ef
    central_params = (float(V_depth(dE, v0, v1)), R0, av, float(Wv_depth(dE, wv0, wv1)), float(Delta_Vv_depth(dE, wv0, wv1)), Rw, aw,
                      float(Ws_depth(dE, ws0, ws1, ws2)), float(Delta_Vs_depth(dE, ws0, ws1, ws2)), Rd, ad)
    spin_orbit_params = (float(Vso_depth(dE, vso0, v1)) + float(Delta_Vso_depth(dE, wso0, wso1)), float(Wso_depth(dE, wso0, wso1)), Rso, aso)
    return central_params, spin_orbit_params, (float(Z * Zp), RC)


def shape_rows(projectile, target, Ecm, Ef, params):
    """Canonical per-pair rows: params (B, 19), Ecm (E,) -> (B, E, 18) float64 (vectorised ``calculate_params``)."""
    p = np.atleast_2d(np.asarray(params, dtype=np.float64))
    if p.shape[1] != len(PARAM_NAMES):
        raise ValueError(f"DOM expects {len(PARAM_NAMES)} parameters in get_param_names() order, got {p.shape[1]}.")
    A, Z = target
    Ap, Zp = projectile
    assert Ap == 1 and Zp in (0, 1)
    (v0, v1, rv, av, wv0, wv1, rw, aw, ws0, ws1, ws2, rd, ad, vso0, wso0, wso1, rso, aso, rC) = (p[:, i][:, None] for i in range(19))
    A13 = A ** (1 / 3)
    RC = rC * A13
    Ec = coulomb_correction(Z * Zp, RC) if Zp == 1 else 0.0
    dE = np.asarray(Ecm, dtype=np.float64)[None, :] - Ec - Ef  # (B, E)
    one = np.ones_like(dE)
    rows = np.stack([
        V_depth(dE, v0, v1), rv * A13 * one, av * one, Wv_depth(dE, wv0, wv1), rw * A13 * one, aw * one,
        Delta_Vs_depth(dE, ws0 * one, ws1 * one, ws2 * one), Ws_depth(dE, ws0, ws1, ws2), rd * A13 * one, ad * one,
        2.0 * (Vso_depth(dE, vso0, v1) + Delta_Vso_depth(dE, wso0, wso1)), rso * A13 * one, aso * one,
        2.0 * Wso_depth(dE, wso0, wso1), rso * A13 * one, aso * one, RC * one, Delta_Vv_depth(dE, wv0, wv1),
    ], axis=-1)
    return np.ascontiguousarray(rows)


class DOM(SingleChannelOpticalModel):
    """dom.py:483-529.  ``evaluate`` keeps the reference's single-sample contract.  In a sweep the 19 global parameters go to
    the device as they are: the energy-dependent map -- including the analytic surface-dispersion integral with its exponential
    integrals of real and complex argument -- runs inside the fused kernel (``JITR_B200_MODEL_DOM``, csrc/potentials.cuh), the
    effective Fermi energy of each (projectile, target) pair comes with the per-energy table (``Reaction(..., Ef=...)``).
    ``DOM(host_map=True)`` keeps the round-1 route: the map on the host, one canonical row per (sample, energy)."""

    def __init__(self, host_map=False):
        super().__init__(params=get_param_names())
        self.host_map = bool(host_map)
        self.model_id = _cabi.MODEL_SHAPE_PAIR if self.host_map else _cabi.MODEL_DOM

    @property
    def pair_rows(self):
        """The sweep asks for this: a callable when the map runs on the host, None when it runs in the kernel."""
        return self.host_pair_rows if self.host_map else None

    @staticmethod
    def _fermi(reaction):
        Ef = getattr(reaction, "Ef", None)
        return 0.0 if Ef is None else float(Ef)

    def host_pair_rows(self, reaction_kinematics, params, device=None):
        """[(reaction, kinematics)] per energy + params (B, 19) -> CUDA tensor (B, E, 18)."""
        import torch

        if isinstance(params, torch.Tensor):
            params = params.detach().cpu().numpy()
        # a sweep may mix targets (and projectiles) over its energy rows: A, Z, the Fermi energy, the Coulomb correction and
        # the radii are those of EACH row's own reaction -- rows are grouped by (projectile, target, Ef) and mapped per group
        groups = {}
        for e, (rx, kin) in enumerate(reaction_kinematics):
            key = ((rx.projectile.A, rx.projectile.Z), (rx.target.A, rx.target.Z), self._fermi(rx))
            groups.setdefault(key, []).append(e)
        rows = np.empty((np.asarray(params).shape[0], len(reaction_kinematics), 18), dtype=np.float64)
        for (proj, tgt, Ef), idx in groups.items():
            Ecm = np.array([reaction_kinematics[e][1].Ecm for e in idx])
            rows[:, idx, :] = shape_rows(proj, tgt, Ecm, Ef, params)
        return torch.tensor(np.ascontiguousarray(rows), dtype=torch.float64, device=device or "cuda")

    def pack(self, params, device=None):
        if self.host_map:
            raise TypeError("DOM(host_map=True) rows depend on the energy: use pair_rows (the workspaces do)")
        return super().pack(params, device)

    def evaluate(self, rgrid, reaction, kinematics, *params):
        if len(params) != self.n_params:
            raise ValueError(f"DOM expects {self.n_params} parameters in get_param_names() order, got {len(params)}.")
        import torch

        from ..xs.elastic import energy_table_row

        _cabi.require_device()
        r = np.atleast_1d(np.asarray(rgrid, dtype=np.float64))
        etab = energy_table_row(reaction, kinematics, kinematics.k, None)
        etab[_cabi.ET_RCH] = 1.0
        dev = torch.device("cuda", torch.cuda.current_device())
        if self.host_map:
            p = self.host_pair_rows([(reaction, kinematics)], np.asarray(params, dtype=np.float64)[None, :], dev)
        else:
            p = self.pack(np.asarray(params, dtype=np.float64)[None, :], dev)
        rt = torch.tensor(r, dtype=torch.float64, device=dev)
        et = torch.tensor(etab[None, :], dtype=torch.float64, device=dev)
        c = torch.empty(r.shape[0], dtype=torch.complex128, device=dev)
        so = torch.empty_like(c)
        co = torch.empty(r.shape[0], dtype=torch.float64, device=dev)
        rc = _cabi.lib().jitr_b200_eval_potentials(
            r.shape[0], self.model_id, 1, 1, _cabi.ptr(rt), _cabi.ptr(et), _cabi.ptr(p), p.shape[-1],
            _cabi.ptr(c), _cabi.ptr(so), _cabi.ptr(co), _cabi.stream_ptr(),
        )
        _cabi.check(rc, "jitr_b200_eval_potentials")
        return c.cpu().numpy(), so.cpu().numpy(), co.cpu().numpy()
