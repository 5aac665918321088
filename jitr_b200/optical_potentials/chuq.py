"""Chapel-Hill (CH89 / CHUQ) global optical potential (optical_potentials/chuq.py); 22 parameters
in ``chuq.get_param_names`` order.  Map chuq.py:140-227 and forms chuq.py:92-137 run on the device."""
from __future__ import annotations

import numpy as np

from .. import _cabi
from .omp import SingleChannelOpticalModel

NUM_POSTERIOR_SAMPLES = 208
NUM_PARAMS = 22

_NAMES = [
    "V0", "Ve", "Vt", "r0", "r0_0", "a0", "Wv0", "Wve0", "Wvew", "rw", "rw_0", "aw", "Ws0", "Wst", "Wse0", "Wsew",
    "Vso", "rso", "rso_0", "aso", "rc", "rc_0",
]
# CH89 (Varner et al. 1991), the defaults of chuq.calculate_params (chuq.py:144-165)
_CH89 = [52.9, -0.299, 13.1, 1.25, -0.225, 0.69, 7.8, 35.0, 16.0, 1.33, -0.42, 0.69, 10.0, 18.0, 36.0, 37.0, 5.9, 1.34, -1.2, 0.63, 1.24, 0.12]


def get_param_names():
    return list(_NAMES)


def get_ch89():
    return np.asarray(_CH89, dtype=np.float64)


def get_samples(posterior="federal"):
    """Posterior samples (NUM_POSTERIOR_SAMPLES, 22) of the CHUQ Federal or Democratic posterior (chuq.py:57-89)."""
    if posterior not in ("federal", "democratic"):
        raise ValueError("posterior must be either 'federal' or 'democratic'")
    from ..data import posterior as _load

    return _load(f"chuq_{posterior}")


class CHUQ(SingleChannelOpticalModel):
    model_id = _cabi.MODEL_CHUQ

    def __init__(self):
        super().__init__(params=get_param_names())
