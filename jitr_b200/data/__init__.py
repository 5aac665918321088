"""Posterior parameter samples of the built-in global optical potentials (input data of the reference,
data/KDUQ*, data/CHUQ*, data/WLHSamples, packed into one array file by tools/make_package_data.py)."""
from __future__ import annotations

import os

import numpy as np

_CACHE = {}


def posterior(key):
    if "npz" not in _CACHE:
        _CACHE["npz"] = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "posteriors.npz"))
    if key not in _CACHE:
        _CACHE[key] = np.ascontiguousarray(_CACHE["npz"][key], dtype=np.float64)
    return _CACHE[key].copy()
