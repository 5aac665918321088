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


def rest_mass(A, Z):
    """Default rest mass (MeV/c^2) of the nuclide (A, Z): what the reference's ``mass.mass(A, Z)[0]`` returns with its default
    mass model (utils/mass.py, used by reactions/reaction.py:167); None when the nuclide is not in the table."""
    if "masses" not in _CACHE:
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "masses.npz"))
        _CACHE["masses"] = {(int(a), int(z)): float(m) for (a, z), m in zip(g["AZ"], g["m0"])}
    return _CACHE["masses"].get((int(A), int(Z)))
