"""Koning-Delaroche (KD03 / KDUQ) global optical potential (optical_potentials/kduq.py).

Parameter ordering follows ``kduq.get_param_names`` of the reference (kduq.py:32-37,200-367):
37 numbers for neutrons, 40 for protons (three Coulomb-radius parameters appended).  The global ->
shape map (kduq.py:370-528) and the forms (kduq.py:134-197) run on the device
(csrc/potentials.cuh); nothing here evaluates a potential on the CPU.
"""
from __future__ import annotations

import numpy as np

from .. import _cabi
from .omp import SingleChannelOpticalModel

NUM_POSTERIOR_SAMPLES = 416

_NAMES_N = [
    "Ef_0", "Ef_A", "v1_0", "v1_asymm", "v1_A", "v2_0", "v2_A", "v3_0", "v3_A", "v4_0", "rv_0", "rv_A", "av_0", "av_A",
    "w1_0", "w1_A", "w2_0", "w2_A", "d1_0", "d1_asymm", "d2_0", "d2_A", "d2_A2", "d2_A3", "d3_0", "rd_0", "rd_A", "ad_0",
    "ad_A", "Vso1_0", "Vso1_A", "Vso2_0", "Wso1_0", "Wso2_0", "rso_0", "rso_A", "aso_0",
]
_NAMES_P = _NAMES_N + ["rc_0", "rc_A", "rc_A2"]

# KD03 frequentist values (Koning & Delaroche 2003), as loaded by the reference from data/KD_default.json
_KD03_N = [
    -11.2814, 0.02646, 59.3, 21.0, 0.024, 0.007228, 1.48e-06, 1.994e-05, 2e-08, 7e-09, 1.3039, 0.4054, 0.6778, 0.0001487,
    12.195, 0.0167, 73.55, 0.0795, 16.0, 16.0, 0.018, 0.003802, 8.0, 156.0, 11.5, 1.3424, 0.01585, 0.5446, 0.0001656,
    5.922, 0.003, 0.004, -3.1, 160.0, 1.1854, 0.647, 0.59,
]
_KD03_P = [
    -8.4075, 0.01378, 59.3, 21.0, 0.024, 0.007067, 4.23e-06, 1.729e-05, 1.136e-08, 7e-09, 1.3039, 0.4054, 0.6778, 0.0001487,
    14.667, 0.009629, 73.55, 0.0795, 16.0, 16.0, 0.018, 0.003802, 8.0, 156.0, 11.5, 1.3424, 0.01585, 0.5187, 0.0005205,
    5.922, 0.003, 0.004, -3.1, 160.0, 1.1854, 0.647, 0.59, 1.198, 0.697, 12.994,
]


def _check(projectile):
    projectile = tuple(projectile)
    if projectile not in ((1, 0), (1, 1)):
        raise RuntimeError("kduq.Global is defined only for neutron and proton projectiles")
    return projectile


def get_param_names(projectile):
    """kduq.py:32-37."""
    return list(_NAMES_N if _check(projectile) == (1, 0) else _NAMES_P)


def get_kd03(projectile):
    """Original KD03 parameter vector in ``get_param_names`` order (kduq.py:40-56)."""
    return np.asarray(_KD03_N if _check(projectile) == (1, 0) else _KD03_P, dtype=np.float64)


def get_samples(projectile, posterior="federal"):
    """Posterior samples (NUM_POSTERIOR_SAMPLES, n_params) of the KDUQ Federal or Democratic posterior
    (kduq.py:59-97), from the packed copy of the reference's data files (jitr_b200/data)."""
    if posterior not in ("federal", "democratic"):
        raise ValueError("posterior must be either 'federal' or 'democratic'")
    from ..data import posterior as _load

    return _load(f"kduq_{posterior}_{'n' if _check(projectile) == (1, 0) else 'p'}")


class KDUQ(SingleChannelOpticalModel):
    """KDUQ optical model (kduq.py:531-559); ``params`` are the global parameters."""

    model_id = _cabi.MODEL_KDUQ

    def __init__(self, projectile):
        self.projectile = _check(projectile)
        super().__init__(params=get_param_names(self.projectile))
