"""Single-channel optical-model plug-in interface (optical_potentials/omp.py:19-149).

``SingleChannelOpticalModel.evaluate(rgrid, reaction, kinematics, *params)`` keeps the reference
contract -- ``(central, spin_orbit, coulomb)`` arrays on ``rgrid`` -- but the built-in models
evaluate their Woods-Saxon / Thomas / charged-sphere forms with the SAME device code the fused
sweep kernel inlines (``jitr_b200_eval_potentials``).  In a sweep nobody calls ``evaluate``: the
workspace's ``*_params`` methods take ``(B, n_params)`` rows and the forms are evaluated on the
mesh inside the solve kernel.
"""
from __future__ import annotations

import numpy as np

from .. import _cabi


class SingleChannelOpticalModel:
    """Base class: subclasses either implement ``evaluate`` on the host (user models; their mesh
    values then go through the ``*_batch`` tensor path) or set ``model_id`` (built-ins)."""

    model_id = None

    def __init__(self, params):
        self.params = params
        self.n_params = len(params)

    # -- built-in device evaluation
    def pack(self, params, device=None):
        """(B, n_params) rows -> contiguous CUDA float64 tensor padded to the kernel's row length."""
        import torch

        if isinstance(params, torch.Tensor):
            p = params.to(device=device or "cuda", dtype=torch.float64)
        else:
            p = torch.tensor(np.atleast_2d(np.asarray(params, dtype=np.float64)), device=device or "cuda")
        if p.dim() == 1:
            p = p.unsqueeze(0)
        need = _cabi.MODEL_NPARAMS[self.model_id]
        if p.shape[1] == need and p.is_contiguous():
            return p  # already in the kernel's padded row layout
        if p.shape[1] != self.n_params:
            raise ValueError(f"expected {self.n_params} parameters per row, got {p.shape[1]}")
        if p.shape[1] < need:
            p = torch.nn.functional.pad(p, (0, need - p.shape[1]))
        return p.contiguous()

    def evaluate(self, rgrid, reaction, kinematics, *params):
        if self.model_id is None:
            raise NotImplementedError("Subclasses must implement the evaluate method.")
        import torch

        from ..xs.elastic import energy_table_row

        _cabi.require_device()
        r = np.atleast_1d(np.asarray(rgrid, dtype=np.float64))
        etab = energy_table_row(reaction, kinematics, kinematics.k, getattr(self, "radius_factor", lambda rx: None)(reaction))
        etab[_cabi.ET_RCH] = 1.0  # r = mesh_x * 1
        dev = torch.device("cuda", torch.cuda.current_device())
        p = self.pack(np.asarray(params, dtype=np.float64)[None, :], dev)
        rt = torch.tensor(r, dtype=torch.float64, device=dev)
        et = torch.tensor(etab[None, :], dtype=torch.float64, device=dev)
        c = torch.empty(r.shape[0], dtype=torch.complex128, device=dev)
        so = torch.empty_like(c)
        co = torch.empty(r.shape[0], dtype=torch.float64, device=dev)
        rc = _cabi.lib().jitr_b200_eval_potentials(
            r.shape[0], self.model_id, 1, 1, _cabi.ptr(rt), _cabi.ptr(et), _cabi.ptr(p), p.shape[1],
            _cabi.ptr(c), _cabi.ptr(so), _cabi.ptr(co), _cabi.stream_ptr(),
        )
        _cabi.check(rc, "jitr_b200_eval_potentials")
        return c.cpu().numpy(), so.cpu().numpy(), co.cpu().numpy()

    def __call__(self, rgrid, reaction, kinematics, *params):
        return self.evaluate(rgrid, reaction, kinematics, *params)


class ShapeModel(SingleChannelOpticalModel):
    """The canonical 17 shape numbers every built-in reduces to:
    Vv Rv av Wv Rw aw Vd Wd Rd ad Vso Rso aso Wso Rwso awso RC, with Vso/Wso the factors multiplying the
    Thomas form (i.e. including 1/kappa^2) and Zz taken from the reaction."""

    model_id = _cabi.MODEL_SHAPE

    def __init__(self):
        super().__init__(["Vv", "Rv", "av", "Wv", "Rw", "aw", "Vd", "Wd", "Rd", "ad", "Vso", "Rso", "aso", "Wso", "Rwso", "awso", "RC"])


class LocalOpticalPotential(SingleChannelOpticalModel):
    """Simple local optical potential with optional nucleus-nucleus radius scaling (omp.py:85-149)."""

    model_id = _cabi.MODEL_LOCAL

    def __init__(self, scale_radii_by_At_and_Ap=False):
        self.central_params = ["Vv", "rv", "av", "Wv", "rw", "aw", "Wd", "Vd", "rd", "ad"]
        self.spin_orbit_params = ["Vso", "Wso", "rso", "aso"]
        self.coulomb_params = ["rC"]
        super().__init__(params=self.central_params + self.spin_orbit_params + self.coulomb_params)
        self.scale_radii_by_At_and_Ap = scale_radii_by_At_and_Ap

    def radius_factor(self, reaction_model):
        """omp.py:106-113."""
        At, Ap = reaction_model.target.A, reaction_model.projectile.A
        if self.scale_radii_by_At_and_Ap:
            return At ** (1 / 3) + Ap ** (1 / 3)
        return At ** (1 / 3)
