"""Whitehead-Lim-Holt (WLH) global microscopic optical potential (optical_potentials/wlh.py); 46 parameters
in ``wlh.get_param_names`` order (the same for neutrons and protons).  The global -> shape map
(wlh.py:251-383) and the forms (wlh.py:61-111) run on the device (csrc/potentials.cuh).
"""
from __future__ import annotations

import numpy as np

from .. import _cabi
from .omp import SingleChannelOpticalModel

NUM_POSTERIOR_SAMPLES = 1000
NUM_PARAMS = 46

_NAMES = ["uv0", "uv1", "uv2", "uv3", "uv4", "uv5", "uv6", "rv0", "rv1", "rv2", "rv3", "av0", "av1", "av2", "av3", "av4", "uw0", "uw1", "uw2", "uw3", "uw4", "rw0", "rw1", "rw2", "rw3", "rw4", "rw5", "aw0", "aw1", "aw2", "aw3", "aw4", "ud0", "ud1", "ud3", "ud4", "rd0", "rd1", "rd2", "ad0", "uso0", "uso1", "rso0", "rso1", "aso0", "aso1"]
# data/WLH_mean.json, as loaded by wlh.Global(projectile) of the reference (wlh.py:124-178)
_MEAN_N = [52.6912521913, 0.2849592984, -0.0002654968, 2.6234e-06, 21.0895801061, 0.2847889774, 0.0010745253, 1.2978610209, 0.312492324, 0.0008999731, 8.6727e-06, 0.741279032, 0.0008329878, 1.02475e-05, 0.2792868326, 0.8764603676, 2.8722965665, 0.2480075276, 0.0004858014, 8.7072302176, 0.0246183387, 0.5629863841, 79.8221560535, 0.7264601343, 86.5142396423, 0.9175065953, 3.6419e-06, 0.2391198977, 0.5849807386, 6.7877182692, 0.0529392097, 0.0005803378, 1.6408748071, 0.0395001751, 2.355450576, 0.1418364594, 1.2882811754, 1.4774733457, 0.0030997663, 0.8506076645, 9.6220335001, 0.0085454732, 1.2794000707, 0.8734769907, 0.8060570111, 0.0003509748]
_MEAN_P = [53.6540319643, 0.300042293, -0.0001192337, 2.1471e-06, 12.918410608, 0.1330858829, 0.0003328253, 1.3041479419, 0.2938537968, 0.0008039876, 6.8659e-06, 0.779602661, 0.0001182833, 5.4588e-06, 0.0445486378, 0.5571679352, 4.3701953072, 0.2267714208, 0.000338132, 11.4340928806, 0.077068475, 0.6541271675, 46.0846969827, 0.6099691135, 58.2530624354, 0.6418349689, 1.5491e-06, 0.5616542778, 0.2708915659, 14.3691362537, 0.4184493049, 0.0024314838, 0.8982409214, 0.0351390393, 0.8135662888, 0.1214099645, 0.9505407272, 0.4391797103, 0.003947669, 0.7649106373, 9.6195980947, 0.0085112416, 1.276250854, 0.8685769478, 0.8032951951, 0.0003508356]


def _check(projectile):
    projectile = tuple(projectile)
    if projectile not in ((1, 0), (1, 1)):
        raise RuntimeError("wlh.Global is defined only for neutron and proton projectiles")
    return projectile


def get_param_names(projectile):
    """wlh.py:31-36."""
    _check(projectile)
    return list(_NAMES)


def get_wlh_mean(projectile):
    """Mean WLH parameter vector (data/WLH_mean.json) in ``get_param_names`` order."""
    return np.asarray(_MEAN_N if _check(projectile) == (1, 0) else _MEAN_P, dtype=np.float64)


def get_samples(projectile):
    """The 1000 WLH parameter samples (wlh.py:40-58), shape (NUM_POSTERIOR_SAMPLES, 46)."""
    from ..data import posterior as _load

    return _load(f"wlh_{'n' if _check(projectile) == (1, 0) else 'p'}")


class WLH(SingleChannelOpticalModel):
    """WLH optical model (wlh.py:386-421); ``params`` are the 46 global parameters."""

    model_id = _cabi.MODEL_WLH

    def __init__(self, projectile):
        self.projectile = _check(projectile)
        super().__init__(params=get_param_names(self.projectile))
