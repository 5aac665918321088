"""Device-side reductions for uncertainty quantification over the sample axis of a sweep.

``percentiles(values, q)`` is ``numpy.percentile(values, q, axis=0)`` (default linear interpolation) on a CUDA tensor
-- the band computation of the reference's ``kduq_uq_demo`` notebook (cell 18) -- so that only
``(len(q), energies, angles)`` numbers leave the GPU after a sweep.  The weighted misfit against data lives in
``ElasticSweep.misfit`` / ``misfit_params``.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi


def percentiles(values, q):
    """values: CUDA float64 tensor (S, ...); q: percentages in [0, 100] (at most 8).  Returns (len(q), ...)."""
    import torch

    _cabi.require_device()
    qs = np.atleast_1d(np.asarray(q, dtype=np.float64))
    if qs.ndim != 1 or not 1 <= qs.size <= 8:
        raise ValueError("q must hold between 1 and 8 percentages")
    if np.any(qs < 0) or np.any(qs > 100):
        raise ValueError("Percentiles must be in the range [0, 100]")
    if values.dim() < 1 or values.shape[0] < 1:
        raise ValueError("values needs a non-empty leading sample axis")
    v = values.to(dtype=torch.float64).contiguous()
    S = v.shape[0]
    M = int(np.prod(v.shape[1:])) if v.dim() > 1 else 1
    out = torch.empty((qs.size,) + tuple(v.shape[1:]), dtype=torch.float64, device=v.device)
    qc = (ctypes.c_double * qs.size)(*qs.tolist())
    with torch.cuda.device(v.device):
        rc = _cabi.lib().jitr_b200_percentiles(_cabi.ptr(v), S, M, ctypes.cast(qc, ctypes.c_void_p), int(qs.size), _cabi.ptr(out),
                                               _cabi.stream_ptr())
    _cabi.check(rc, "jitr_b200_percentiles")
    return out
