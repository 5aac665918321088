#!/usr/bin/env python
"""Pack the reference's posterior parameter files (READ-ONLY INPUT DATA: data/KDUQ{Federal,Democratic},
data/CHUQ{Federal,Democratic}, data/WLHSamples of the reference repository, thousands of small JSON
files) into ONE array file, jitr_b200/data/posteriors.npz, read by kduq.get_samples / chuq.get_samples /
wlh.get_samples (optical_potentials/kduq.py:59-97, chuq.py:57-89, wlh.py:40-58).  Run in the build
container only (it imports the reference through tests/golden/_ref_env.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import _ref_env  # noqa: E402

_ref_env.setup()
from jitr.optical_potentials import chuq, kduq, wlh  # noqa: E402

out = {}
for post in ("federal", "democratic"):
    out[f"kduq_{post}_n"] = kduq.get_samples((1, 0), post)
    out[f"kduq_{post}_p"] = kduq.get_samples((1, 1), post)
    out[f"chuq_{post}"] = chuq.get_samples(post)
out["wlh_n"] = wlh.get_samples((1, 0))
out["wlh_p"] = wlh.get_samples((1, 1))
for k, v in out.items():
    print(k, v.shape, v.dtype)
np.savez_compressed(os.path.join(ROOT, "jitr_b200", "data", "posteriors.npz"), **out)

# ---- default rest masses (utils/mass.py: mass.mass(A, Z)[0] with the default model, as reactions/reaction.py:167 uses it)
from jitr.utils import mass as _mass  # noqa: E402

_mass.init_mass_db()
db = _mass.__MASS_DB__
AZ, m0 = [], []
for A, Z in sorted(set(zip(db["A"].astype(int), db["Z"].astype(int)))):
    if A < 2:
        continue
    try:
        v = _mass.mass(int(A), int(Z))[0]
    except Exception:
        continue
    if v is not None and np.isfinite(v):
        AZ.append((A, Z))
        m0.append(float(v))
np.savez_compressed(os.path.join(ROOT, "jitr_b200", "data", "masses.npz"), AZ=np.array(AZ, dtype=np.int32), m0=np.array(m0))
print("masses", len(m0))
