"""Nonlocal potential forms generated on the device (optical_potentials/potential_forms.py:14-19).

``perey_buck_nonlocal(r, rp, beta, ell)`` of the reference evaluates ``spherical_jn`` at an imaginary argument and overflows
beyond z ~ 700; the device routine uses the exponentially scaled modified spherical Bessel function instead
(csrc/nonlocal_gen.cu) and produces the kernels of all partial waves of a batch of samples in one launch, ready for
``Solver.solve_batch(nonlocal_potential=...)``.
"""
from __future__ import annotations

import numpy as np

from .. import _cabi

TWO_PI = 2.0 * np.pi


def perey_buck_kernels(rgrid, lmax, V, W, beta, R=None, a=None, zfac=2.0, device=None, l_major=False, out=None):
    """(B, lmax + 1, N, N) complex128 CUDA tensor
    ``U_l(r, r') = -(V + iW) f_WS((r + r')/2; R, a) H_l(r, r'; beta)``, ``H_l = 2 z i_l(z) exp(-(r^2 + r'^2)/beta^2) / (beta sqrt(pi))``,
    ``z = zfac r r'/beta^2``.  ``V``, ``W``, ``beta`` (and ``R``, ``a``): scalars or (B,) arrays; without ``R``/``a`` there is no
    Woods-Saxon factor.  ``zfac = 2`` is the Gaussian nonlocality; ``zfac = 2 pi`` is the literal ``z`` of the reference's
    ``perey_buck_nonlocal`` (with ``V = -1, W = 0`` its values).  ``l_major``: return (lmax + 1, B, N, N) instead, one
    contiguous batch per partial wave."""
    import torch

    _cabi.require_device()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    cols = [np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in (V, W, beta, 0.0 if R is None else R, -1.0 if a is None else a)]
    B = max(len(c) for c in cols)
    p = np.stack([np.broadcast_to(c, (B,)) for c in cols], axis=1)
    rg = torch.as_tensor(np.ascontiguousarray(rgrid, dtype=np.float64), device=dev)
    N = rg.shape[0]
    pt = torch.as_tensor(np.ascontiguousarray(p), device=dev)
    shape = (lmax + 1, B, N, N) if l_major else (B, lmax + 1, N, N)
    if out is None:
        out = torch.empty(shape, dtype=torch.complex128, device=dev)
    assert tuple(out.shape) == shape and out.is_contiguous()
    ss, ls = (N * N, B * N * N) if l_major else (0, 0)
    with torch.cuda.device(dev):
        rc = _cabi.lib().jitr_b200_perey_buck_kernels(N, int(lmax), B, _cabi.ptr(rg), _cabi.ptr(pt), float(zfac), _cabi.ptr(out), ss, ls,
                                                      _cabi.stream_ptr())
    _cabi.check(rc, "jitr_b200_perey_buck_kernels")
    return out


def perey_buck_nonlocal(r, rp, beta, ell, device=None):
    """Device counterpart of ``potential_forms.perey_buck_nonlocal(r, rp, beta, ell)`` on the mesh ``r`` x ``rp`` (1-D arrays of the
    same mesh, as produced by ``Solver.nonlocal_radial_grids``): an (N, N) real CUDA tensor, finite where the reference overflows."""
    r = np.asarray(r, dtype=np.float64)
    if r.ndim == 2:
        r = r[:, 0] if np.ptp(r[:, 0]) > 0 else r[0]
    out = perey_buck_kernels(r, int(ell), -1.0, 0.0, beta, zfac=TWO_PI, device=device)
    return out[0, int(ell)].real
