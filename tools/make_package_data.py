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
