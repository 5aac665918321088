"""Multi-precision referee for the parity gate (TEST INFRASTRUCTURE ONLY, see oracle/jitr_oracle.py).

The reference obtains R through the explicit LAPACK inverse of A (rmatrix/core.py:7-22); cond(A) reaches 1e7, so the
reference's own S carries rounding noise of order cond * eps ~ 1e-10 .. 1e-9 on a few systems of a large sweep.  Where
the CUDA path and the float64 oracle differ by more than the 1e-10 gate, this module decides who is right: it solves
the SAME float64 matrix the oracle builds (xs/elastic.py:145-175) exactly -- mpmath, 40 digits -- and returns that S.
"""
from __future__ import annotations

import numpy as np


def smatrix_element_mp(ws, central, spin_orbit, coulomb, l, jj, dps=40):
    """S_{l, j = l + 1/2 (jj = 0) or l - 1/2 (jj = 1)} of the oracle ElasticWorkspace `ws` for the given mesh potentials,
    from a `dps`-digit solve of the float64 matrix  free[l] + (central + coulomb + (l.s) spin_orbit) / Ecm."""
    import mpmath as mp

    from oracle import jitr_oracle as orc

    N = ws.N
    so = np.zeros(N, dtype=np.complex128) if spin_orbit is None else np.asarray(spin_orbit, dtype=np.complex128)
    im_c = np.diag(np.asarray(central, dtype=np.complex128)) / ws.Ecm
    im_so = np.diag(so) / ws.Ecm
    if coulomb is not None:
        im_c = im_c + np.diag(np.asarray(coulomb, dtype=np.complex128)) / ws.Ecm
    if l == 0:
        A = ws.free[0] + im_c
    else:
        cp, cm = orc.spin_half_orbit_couplings(l)
        A = ws.free[l] + (im_c + (cp if jj == 0 else cm) * im_so)
    with mp.workdps(dps):
        Am = mp.matrix(N, N)
        for i in range(N):
            for j in range(N):
                Am[i, j] = mp.mpc(A[i, j].real, A[i, j].imag)
        bm = mp.matrix([mp.mpf(float(v.real)) for v in ws.b])
        x = mp.lu_solve(Am, bm)
        R = sum(bm[i] * x[i] for i in range(N)) / mp.mpf(ws.a) ** 2
        Hp, Hm, Hpp, Hmp = (mp.mpc(complex(v).real, complex(v).imag) for v in ws.asym[l])
        a = mp.mpf(ws.a)
        S = (Hm - a * R * Hmp) / (Hp - a * R * Hpp)
        return complex(S)


def _oracle_matrix(ws, central, spin_orbit, coulomb, l, jj):
    from oracle import jitr_oracle as orc

    N = ws.N
    so = np.zeros(N, dtype=np.complex128) if spin_orbit is None else np.asarray(spin_orbit, dtype=np.complex128)
    im_c = np.diag(np.asarray(central, dtype=np.complex128)) / ws.Ecm
    im_so = np.diag(so) / ws.Ecm
    if coulomb is not None:
        im_c = im_c + np.diag(np.asarray(coulomb, dtype=np.complex128)) / ws.Ecm
    if l == 0:
        return ws.free[0] + im_c
    cp, cm = orc.spin_half_orbit_couplings(l)
    return ws.free[l] + (im_c + (cp if jj == 0 else cm) * im_so)


def smatrix_element_refined(ws, central, spin_orbit, coulomb, l, jj, steps=3):
    """The same quantity as smatrix_element_mp, ~1000x faster: float64 LU solve of the oracle's matrix followed by
    `steps` rounds of iterative refinement with the residual accumulated in 80-bit extended precision
    (np.clongdouble; eps 1e-19), which converges to cond(A) * 1e-19 ~ 1e-12 -- two orders below the 1e-10 gate
    (checked against the 40-digit solve in tests/test_oracle_vs_golden.py)."""
    import scipy.linalg as sla

    if np.finfo(np.longdouble).eps > 1e-18:
        return smatrix_element_mp(ws, central, spin_orbit, coulomb, l, jj)
    return _refined_from_matrix(ws, _oracle_matrix(ws, central, spin_orbit, coulomb, l, jj), l, steps)


def _refined_from_matrix(ws, A, l, steps=3):
    import scipy.linalg as sla

    b = np.asarray(ws.b, dtype=np.complex128)
    lu = sla.lu_factor(A)
    x = sla.lu_solve(lu, b).astype(np.clongdouble)
    Al, bl = A.astype(np.clongdouble), b.astype(np.clongdouble)
    for _ in range(steps):
        r = bl - Al @ x
        x = x + sla.lu_solve(lu, r.astype(np.complex128)).astype(np.clongdouble)
    R = (bl @ x) / np.longdouble(ws.a) ** 2
    Hp, Hm, Hpp, Hmp = (np.clongdouble(complex(v)) for v in ws.asym[l])
    a = np.longdouble(ws.a)
    return complex((Hm - a * R * Hmp) / (Hp - a * R * Hpp))


def smatrix_input_noise(ws, central, spin_orbit, coulomb, l, jj, trials=4, seed=1):
    """What ONE ULP of uncertainty in the float64 matrix does to this S-matrix element: max over `trials` random sign patterns E of
    |S(A o (1 + 2^-53 E)) - S(A)| / |S(A)|, every solve exact to ~1e-12 (smatrix_element_refined).  Any float64 algorithm --
    LAPACK's explicit inverse in the reference as much as an LU or LDL^T on the GPU -- is backward stable at best: its result is the
    exact one of a matrix perturbed by a modest multiple p(n) of that, so p(n) x this number is the accuracy to which the
    reference's value is DEFINED.  The parity gate uses it to judge the handful of elements of a large sweep that differ by more
    than 1e-10 (tools/parity_config2.referee_outliers)."""
    if np.finfo(np.longdouble).eps > 1e-18:
        return float("nan")
    A = _oracle_matrix(ws, central, spin_orbit, coulomb, l, jj)
    s0 = _refined_from_matrix(ws, A, l)
    rng = np.random.default_rng(seed)
    u = 2.0 ** -53
    worst = 0.0
    for _ in range(trials):
        E = rng.choice([-1.0, 1.0], size=A.shape)
        E = np.triu(E) + np.triu(E, 1).T  # a symmetric perturbation keeps the matrix complex symmetric
        s1 = _refined_from_matrix(ws, A * (1.0 + u * E), l)
        worst = max(worst, abs(s1 - s0) / abs(s0))
    return float(worst)


def _referee_chunk(args):
    return referee(*args, procs=1)


def referee(energy_table, rows, items, N=40, lmax=30, projectile=(1, 0), dps=40, procs=None):
    """items: list of (row index, energy index, l, jj).  Returns the list of exact-to-1e-12 S values (complex128):
    dps > 0 -> mpmath at `dps` digits, dps = 0 -> float64 LU + 80-bit iterative refinement; dps < 0 -> not S but its sensitivity to
    one ulp of input uncertainty (smatrix_input_noise)."""
    from oracle import jitr_oracle as orc

    import os

    procs = procs or os.cpu_count() or 1
    if procs > 1 and len(items) > 1:
        import multiprocessing as mp_

        procs = min(procs, len(items))
        chunks = [items[i::procs] for i in range(procs)]
        with mp_.get_context("fork").Pool(procs) as pool:
            res = pool.map(_referee_chunk, [(energy_table, rows, ch, N, lmax, projectile, dps) for ch in chunks])
        out = [None] * len(items)
        for i, r in enumerate(res):
            out[i::procs] = r
        return out

    out = []
    cache = {}
    for (i, e, l, jj) in items:
        if e not in cache:
            tgt, Elab, Rch, kin = energy_table[e]
            cache[e] = (orc.ElasticWorkspace(N, lmax, Rch, kin, np.linspace(0.1, np.pi, 4), tgt[1] * projectile[1]), tgt, Elab)
        ws, tgt, Elab = cache[e]
        with np.errstate(over="ignore"):
            c, so, co = orc.kduq_potentials(ws.rgrid, projectile, tgt, Elab, rows[i])
        if dps < 0:
            out.append(smatrix_input_noise(ws, c, so, co, l, jj))
        else:
            out.append(smatrix_element_mp(ws, c, so, co, l, jj, dps) if dps > 0 else smatrix_element_refined(ws, c, so, co, l, jj))
    return out
