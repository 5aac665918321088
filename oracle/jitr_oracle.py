"""CPU oracle: a numpy restatement of jitr's calculable R-matrix hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` leg may import it; nothing under ``jitr_b200/`` does.

Every function cites the reference file:line it restates (paths relative to
``/root/reference/src/jitr``).  The restatement is pinned against golden
vectors produced by importing the *real* reference in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``); see
``tests/test_oracle_vs_golden.py``.  Parity status: PINNED (every function
below is compared with reference outputs at <= 1e-12 relative).

The algorithm is deliberately the reference's own: explicit inverse of the
system matrix per (l, j) (``np.linalg.inv`` -> LAPACK zgetrf+zgetri), Python
loop over partial waves with the data-dependent early ``break``; this is what
the CPU baseline times.
"""
from __future__ import annotations

import numpy as np

# utils/constants.py:5-7
HBARC = 197.3269804
ALPHA = 1.0 / 137.0359991
WAVENUMBER_PION = np.sqrt(1 / 2)
# optical_potentials/potential_forms.py:11
MAX_ARG = np.log(1 / 1e-16)


# ----------------------------------------------------------------------------
# mesh / kinetic operator
# ----------------------------------------------------------------------------
_MESH_CACHE = {}


def legendre_mesh(nbasis):
    """Gauss-Legendre nodes/weights mapped to (0,1).  quadrature/quadrature.py:66-74.  Computed once per nbasis, as the
    reference does in Kernel.__init__ (quadrature/kernel.py:30-55); callers get copies."""
    if nbasis not in _MESH_CACHE:
        t, w = np.polynomial.legendre.leggauss(nbasis)
        _MESH_CACHE[nbasis] = (0.5 * (t + 1.0), 0.5 * w)
    x, w = _MESH_CACHE[nbasis]
    return x.copy(), w.copy()


def kinetic_bloch_matrix(nbasis, a, l):
    """Kinetic + Bloch operator T(l, a) on the Lagrange-Legendre mesh (Baye 2015
    eqs 3.128/3.129 scaled by 1/E with r -> s = k r).
    quadrature/quadrature.py:180-229 (element formulas :197-218, symmetrisation :224-229)."""
    x, _ = legendre_mesh(nbasis)
    N = nbasis
    T = np.zeros((N, N), dtype=np.complex128)
    for n in range(N):
        xn = x[n]
        T[n, n] = ((4 * N**2 + 4 * N + 3) * xn * (1 - xn) - 6 * xn + 1) / (
            3 * xn**2 * (1 - xn) ** 2
        ) / a**2 + l * (l + 1) / (a * xn) ** 2
        for m in range(n + 1, N):
            xm = x[m]
            # (n, m) are 0-based here; (-1)**(n+m) has the same parity as the 1-based form
            val = (
                (-1.0) ** (n + m)
                * (
                    (N**2 + N + 1.0)
                    + (xn + xm - 2 * xn * xm) / (xn - xm) ** 2
                    - 1.0 / (1.0 - xn)
                    - 1.0 / (1.0 - xm)
                )
                / np.sqrt(xn * xm * (1.0 - xn) * (1.0 - xm))
                / a**2
            )
            T[n, m] = val
            T[m, n] = val
    return T


def boundary_vector(nbasis):
    """Lagrange-Legendre basis functions at the channel radius, f_n(a).
    rmatrix/rmatrix.py:36-42 with quadrature/quadrature.py:37-55 at s = a
    (x = 1, P_N(1) = 1)  ->  (-1)^(N-n) sqrt((1-x_n)/x_n) / (1 - x_n), n = 1..N."""
    from scipy.special import eval_legendre

    x, _ = legendre_mesh(nbasis)
    N = nbasis
    n = np.arange(1, N + 1)
    xx = 1.0
    vals = (
        (-1.0) ** (N - n)
        * np.sqrt((1 - x) / x)
        * eval_legendre(N, 2.0 * xx - 1.0)
        * xx
        / (xx - x)
    )
    return vals.astype(np.complex128)


def basis_function(nbasis, n, a, s):
    """n-th (1-based) Lagrange-Legendre function at s.  quadrature/quadrature.py:37-55."""
    from scipy.special import eval_legendre

    x, _ = legendre_mesh(nbasis)
    N = nbasis
    xx = np.asarray(s) / a
    xn = x[n - 1]
    return (
        (-1.0) ** (N - n)
        * np.sqrt((1 - xn) / xn)
        * eval_legendre(N, 2.0 * xx - 1.0)
        * xx
        / (xx - xn)
    )


def free_matrix(nbasis, a, l, E=None, mu=None):
    """blockdiag(T(l_i) mu_0/mu_i) - blockdiag(I E_i/E_0).  rmatrix/rmatrix.py:56-124."""
    l = np.atleast_1d(np.asarray(l))
    nch = l.size
    mu = np.ones(nch) if mu is None else np.asarray(mu, dtype=np.float64)
    E = np.ones(nch) if E is None else np.asarray(E, dtype=np.float64)
    F = np.zeros((nch * nbasis, nch * nbasis), dtype=np.complex128)
    for i in range(nch):
        sl = slice(i * nbasis, (i + 1) * nbasis)
        F[sl, sl] += kinetic_bloch_matrix(nbasis, a, int(l[i])) * mu[0] / mu[i]
        F[sl, sl] -= np.eye(nbasis) * E[i] / E[0]
    return F


def interaction_matrix(nbasis, k0, E0, a, nch, local_potential=None, nonlocal_potential=None):
    """Interaction matrix from mesh values.  rmatrix/rmatrix.py:126-178,
    quadrature/kernel.py:180-214 (nonlocal weighting sqrt(w_n w_m) * V * radius)."""
    N = nbasis
    _, w = legendre_mesh(N)
    V = np.zeros((N * nch, N * nch), dtype=np.complex128)
    radius = a / k0
    if local_potential is not None:
        loc = np.asarray(local_potential, dtype=np.complex128)
        if loc.ndim == 1:
            loc = loc.reshape(1, 1, N)
        for i in range(nch):
            for j in range(nch):
                V[i * N : (i + 1) * N, j * N : (j + 1) * N] = np.diag(loc[i, j])
    if nonlocal_potential is not None:
        nl = np.asarray(nonlocal_potential, dtype=np.complex128)
        if nl.ndim == 2:
            nl = nl.reshape(1, 1, N, N)
        sw = np.sqrt(np.outer(w, w))
        nl = sw[None, None] * nl * radius
        V += nl.swapaxes(1, 2).reshape(N * nch, N * nch) / k0
    return V / E0


# ----------------------------------------------------------------------------
# the solve  (rmatrix/core.py)
# ----------------------------------------------------------------------------
def _rmatrix_with_inverse_py(A, b, nch, nbasis, a):
    """R_ij = b^T (A^-1)_ij b / a^2 using the explicit inverse.  rmatrix/core.py:7-22."""
    C = np.linalg.inv(A)
    R = np.zeros((nch, nch), dtype=np.complex128)
    for i in range(nch):
        for j in range(nch):
            Cblock = np.ascontiguousarray(C[i * nbasis : (i + 1) * nbasis, j * nbasis : (j + 1) * nbasis])
            R[i, j] = b.T @ Cblock @ b
    return R / a**2, C


def _solve_smatrix_py(A, b, Hp, Hm, Hpp, Hmp, weights, a, nch, nbasis):
    """R, S, A^-1 and u'_ext(a).  rmatrix/core.py:25-60."""
    R, Ainv = rmatrix_with_inverse(A, b, nch, nbasis, a)
    Zp = np.diag(Hp) - R @ np.diag(Hpp) * a
    Zm = np.diag(Hm) - R @ np.diag(Hmp) * a
    S = np.linalg.solve(Zp, Zm)
    uext_prime = 1j / 2 * (Hmp * weights - S @ np.copy(Hpp))
    return R, S, Ainv, uext_prime


# The reference compiles these two functions with numba @njit (np.linalg.inv -> LAPACK zgetrf+zgetri,
# about 2x faster than numpy's inv).  The CPU baseline must not be slower than the reference for
# a reason that is not the reference's, so the same decorator is applied when numba is importable.
try:  # pragma: no cover - depends on the image
    import numba as _numba

    rmatrix_with_inverse = _numba.njit(_rmatrix_with_inverse_py)
    solve_smatrix = _numba.njit(_solve_smatrix_py)
    HAVE_NUMBA = True
except Exception:  # numba missing: plain numpy (same numbers to rounding)
    rmatrix_with_inverse = _rmatrix_with_inverse_py
    solve_smatrix = _solve_smatrix_py
    HAVE_NUMBA = False


def solution_coeffs(Ainv, b, uext_prime, nch, nbasis):
    """x = A^-1 vec_i(b u'_ext,i) -> (nch, nbasis).  rmatrix/core.py:63-82."""
    x = (b * uext_prime[:, None]).reshape(nch * nbasis)
    return (Ainv @ x).reshape(nch, nbasis)


def solver_solve(
    nbasis, a, l, E, k, mu, asym, local_potential=None, nonlocal_potential=None,
    weights=None, wavefunction=False,
):
    """End-to-end ``Solver.solve`` for one coupled set.  rmatrix/rmatrix.py:180-246.
    ``asym`` = (Hp, Hm, Hpp, Hmp) arrays of length nch."""
    l = np.atleast_1d(l)
    nch = l.size
    E = np.atleast_1d(np.asarray(E, dtype=np.float64))
    k = np.atleast_1d(np.asarray(k, dtype=np.float64))
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    F = free_matrix(nbasis, a, l, E, mu)
    b = boundary_vector(nbasis)
    if weights is None:
        weights = np.zeros(nch)
        weights[0] = 1
    V = interaction_matrix(nbasis, k[0], E[0], a, nch, local_potential, nonlocal_potential)
    R, S, Ainv, up = solve_smatrix(F + V, b, *asym, weights, a, nch, nbasis)
    if not wavefunction:
        return R, S, up
    return R, S, solution_coeffs(Ainv, b, up, nch, nbasis), up


# ----------------------------------------------------------------------------
# asymptotics  (utils/free_solutions.py, reactions/system.py)
# ----------------------------------------------------------------------------
def coulomb_hankel(l, eta, s):
    """(H+, H-, H+', H-') at s.  utils/free_solutions.py:39-109: H+- = G +- iF from
    mpmath.coulombf/g; derivative by the l -> l+1 recurrence (:73-84)."""
    from mpmath import coulombf, coulombg

    def FG(ll):
        return np.complex128(coulombf(ll, eta, s)), np.complex128(coulombg(ll, eta, s))

    F0, G0 = FG(l)
    F1, G1 = FG(l + 1)
    Hp0, Hm0 = G0 + 1j * F0, G0 - 1j * F0
    Hp1, Hm1 = G1 + 1j * F1, G1 - 1j * F1
    rec = np.sqrt(1 + eta**2 / (l + 1) ** 2)
    shift = (l + 1) / s + eta / (l + 1)
    return Hp0, Hm0, shift * Hp0 - rec * Hp1, shift * Hm0 - rec * Hm1


def spin_half_orbit_couplings(l):
    """(l.s) expectation values for j = l+1/2 [, l-1/2].  reactions/system.py:23-33."""
    js = [l + 0.5] if l == 0 else [l + 0.5, l - 0.5]
    return np.array([j * (j + 1) - l * (l + 1) - 0.75 for j in js])


# ----------------------------------------------------------------------------
# potentials  (optical_potentials/potential_forms.py, kduq.py, chuq.py, omp.py)
# ----------------------------------------------------------------------------
def woods_saxon_safe(r, R, a):
    """potential_forms.py:40-50."""
    x = (r - R) / a
    out = np.zeros_like(r, dtype=np.float64)
    m = x <= MAX_ARG
    out[m] = 1.0 / (1.0 + np.exp(x[m]))
    return out


def woods_saxon_prime_safe(r, R, a):
    """potential_forms.py:53-67."""
    x = (r - R) / a
    out = np.zeros_like(r, dtype=np.float64)
    m = x <= MAX_ARG
    out[m] = -np.exp(x[m]) / (a * (1 + np.exp(x[m])) ** 2)
    return out


def thomas_safe(r, R, a):
    """potential_forms.py:70-88."""
    x = (r - R) / a
    y = 1.0 / r
    out = np.zeros_like(r, dtype=np.float64)
    m = x <= MAX_ARG
    out[m] = y[m] * (-np.exp(x[m]) / (a * (1 + np.exp(x[m])) ** 2))
    return out


def coulomb_charged_sphere(r, zz, r_c):
    """potential_forms.py:134-153."""
    if zz == 0:
        return np.zeros_like(r, dtype=np.float64)
    m = r <= r_c
    out = np.zeros_like(r, dtype=np.float64)
    out[m] = 1.0 / (2.0 * r_c) * (3.0 - (r[m] / r_c) ** 2)
    out[~m] = 1.0 / r[~m]
    return zz * ALPHA * HBARC * out


def kduq_shape_params(projectile, target, Elab, p):
    """KDUQ global parameters -> (central, spin-orbit, coulomb) shape parameters.
    optical_potentials/kduq.py:370-528 (energy dependence :100-131).
    ``p`` is the 37- (neutron) or 40-vector (proton) in get_param_names order."""
    p = list(p) + [0.0] * (40 - len(p))
    (Ef_0, Ef_A, v1_0, v1_asymm, v1_A, v2_0, v2_A, v3_0, v3_A, v4_0, rv_0, rv_A, av_0, av_A,
     w1_0, w1_A, w2_0, w2_A, d1_0, d1_asymm, d2_0, d2_A, d2_A2, d2_A3, d3_0, rd_0, rd_A,
     ad_0, ad_A, Vso1_0, Vso1_A, Vso2_0, Wso1_0, Wso2_0, rso_0, rso_A, aso_0,
     rc_0, rc_A, rc_A2) = p
    A, Z = target
    Ap, Zp = projectile
    asym = (A - 2 * Z) / A
    factor = (-1) ** (Zp + 1)
    asym *= factor
    Ef = Ef_0 + Ef_A * A
    v1 = v1_0 + v1_asymm * asym - v1_A * A
    v2 = v2_0 + v2_A * A * factor
    v3 = v3_0 + v3_A * A * factor
    v4 = v4_0
    dE = Elab - Ef
    vv = v1 * (1 - v2 * dE + v3 * dE**2 - v4 * dE**3)
    rv = rv_0 - rv_A * A ** (-1.0 / 3.0)
    av = av_0 - av_A * A
    w1 = w1_0 + w1_A * A
    w2 = w2_0 + w2_A * A
    wv = w1 * dE**2 / (dE**2 + w2**2)
    d1 = d1_0 + d1_asymm * asym
    with np.errstate(over="ignore"):
        d2 = d2_0 + d2_A / (1 + np.exp((A - d2_A3) / d2_A2))
    d3 = d3_0
    wd = d1 * dE**2 / (dE**2 + d3**2) * np.exp(-d2 * dE)
    rd = rd_0 - rd_A * A ** (1.0 / 3.0)
    ad = ad_0 + ad_A * A * factor
    vso = (Vso1_0 + Vso1_A * A) * np.exp(-Vso2_0 * dE)
    rso = rso_0 - rso_A * A ** (-1.0 / 3.0)
    aso = aso_0
    wso = Wso1_0 * dE**2 / (dE**2 + Wso2_0**2)
    R_C = rv * A ** (1.0 / 3.0)
    if Zp == 1:
        rc0 = rc_0 + rc_A * A ** (-2.0 / 3.0) + rc_A2 * A ** (-5.0 / 3.0)
        R_C = rc0 * A ** (1.0 / 3.0)
        Vcbar = 1.73 / rc0 * Z * A ** (-1.0 / 3.0)
        vv += v1 * Vcbar * (v2 - 2 * v3 * dE + 3 * v4 * dE**2)
    A13 = A ** (1.0 / 3.0)
    central = (vv, rv * A13, av, wv, rv * A13, av, wd, rd * A13, ad)
    so = (vso, rso * A13, aso, wso, rso * A13, aso)
    return central, so, (Z * Zp, R_C)


def kduq_potentials(r, projectile, target, Elab, p):
    """(central, spin_orbit, coulomb) on grid r.  kduq.py:134-197, :541-559."""
    (vv, Rv, av, wv, Rw, aw, wd, Rd, ad), (vso, Rso, aso, wso, Rwso, awso), (zz, RC) = (
        kduq_shape_params(projectile, target, Elab, p)
    )
    central = (
        -vv * woods_saxon_safe(r, Rv, av)
        - 1j * wv * woods_saxon_safe(r, Rw, aw)
        - 1j * (-4 * ad) * wd * woods_saxon_prime_safe(r, Rd, ad)
    )
    so = vso / WAVENUMBER_PION**2 * thomas_safe(r, Rso, aso) + 1j * wso / WAVENUMBER_PION**2 * thomas_safe(
        r, Rwso, awso
    )
    return central.astype(np.complex128), so.astype(np.complex128), coulomb_charged_sphere(r, zz, RC)


def chuq_shape_params(projectile, target, Elab, p):
    """CHUQ global parameters (22) -> shape parameters.  chuq.py:140-227."""
    (V0, Ve, Vt, r0, r0_0, a0, Wv0, Wve0, Wvew, rw, rw_0, aw, Ws0, Wst, Wse0, Wsew,
     Vso, rso, rso_0, aso, rc, rc_0) = p
    A, Z = target
    Ap, Zp = projectile
    is_p = Zp == 1
    alpha = (A - Z - Z) / A
    asym = (1 if is_p else -1) * alpha
    R0 = r0 * A ** (1 / 3) + r0_0
    Rw = rw * A ** (1 / 3) + rw_0
    Rso = rso * A ** (1 / 3) + rso_0
    RC = rc * A ** (1 / 3) + rc_0
    Ec = 6.0 * Z * ALPHA * HBARC / (5 * RC) if is_p else 0.0
    dE = Elab - Ec
    V = V0 + Ve * dE + asym * Vt
    Wv = Wv0 / (1 + np.exp((Wve0 - dE) / Wvew))
    Ws = (Ws0 + asym * Wst) / (1 + np.exp((dE - Wse0) / Wsew))
    return (V, Wv, Ws, R0, a0, Rw, aw), (Vso, Rso, aso), (Z * Zp, RC)


def chuq_potentials(r, projectile, target, Elab, p):
    """chuq.py:92-137."""
    (V, W, Wd, Rv, av, Rd, ad), (Vso, Rso, aso), (zz, RC) = chuq_shape_params(
        projectile, target, Elab, p
    )
    central = -(V * woods_saxon_safe(r, Rv, av)) - 1j * W * woods_saxon_safe(r, Rd, ad) - (
        -(4j * ad * Wd) * woods_saxon_prime_safe(r, Rd, ad)
    )
    so = 2 * Vso * thomas_safe(r, Rso, aso)
    return central.astype(np.complex128), so.astype(np.complex128), coulomb_charged_sphere(r, zz, RC)


def wlh_shape_params(projectile, target, Elab, p):
    """WLH global parameters (46) -> shape parameters.  wlh.py:251-383."""
    (uv0, uv1, uv2, uv3, uv4, uv5, uv6, rv0, rv1, rv2, rv3, av0, av1, av2, av3, av4, uw0, uw1, uw2, uw3, uw4,
     rw0, rw1, rw2, rw3, rw4, rw5, aw0, aw1, aw2, aw3, aw4, ud0, ud1, ud3, ud4, rd0, rd1, rd2, ad0,
     uso0, uso1, rso0, rso1, aso0, aso1) = p
    A, Z = target
    Ap, Zp = projectile
    asym_factor = (A - 2 * Z) / A
    factor = (-1) ** (Zp + 1)
    uv = uv0 - uv1 * Elab + uv2 * Elab**2 + uv3 * Elab**3 + factor * (uv4 - uv5 * Elab + uv6 * Elab**2) * asym_factor
    rv = rv0 - rv1 * A ** (-1.0 / 3) - rv2 * Elab + rv3 * Elab**2
    av = av0 - factor * av1 * Elab - av2 * Elab**2 - (av3 - av4 * asym_factor) * asym_factor
    uw = uw0 + uw1 * Elab - uw2 * Elab**2 + (factor * uw3 - uw4 * Elab) * asym_factor
    rw = rw0 + (rw1 + rw2 * A) / (rw3 + A + rw4 * Elab) + rw5 * Elab**2
    aw = aw0 - (aw1 * Elab) / (-aw2 - Elab) + (aw3 - aw4 * Elab) * asym_factor
    if (tuple(projectile) == (1, 0) and Elab < 40) or (tuple(projectile) == (1, 1) and Elab < 20 and A > 100):
        ud = -(ud0 - ud1 * Elab - (ud3 - ud4 * Elab) * asym_factor)
    else:
        ud = 0
    rd = rd0 - rd2 * Elab - rd1 * A ** (-1.0 / 3)
    ad = ad0
    uso = uso0 - uso1 * A
    rso = rso0 - rso1 * A ** (-1.0 / 3.0)
    aso = aso0 - aso1 * A
    A13 = A ** (1.0 / 3.0)
    return (uv, rv * A13, av, uw, rw * A13, aw, ud, rd * A13, ad), (uso, rso * A13, aso), (Z * Zp, rv * A13)


def wlh_potentials(r, projectile, target, Elab, p):
    """(central, spin_orbit, coulomb) on grid r.  wlh.py:61-111, :396-421."""
    (uv, Rv, av, uw, Rw, aw, ud, Rd, ad), (uso, Rso, aso), (zz, RC) = wlh_shape_params(projectile, target, Elab, p)
    central = (
        -uv * woods_saxon_safe(r, Rv, av)
        - 1j * uw * woods_saxon_safe(r, Rw, aw)
        - 1j * (-4 * ad) * ud * woods_saxon_prime_safe(r, Rd, ad)
    )
    so = (uso / WAVENUMBER_PION**2) * thomas_safe(r, Rso, aso)
    return central.astype(np.complex128), so.astype(np.complex128), coulomb_charged_sphere(r, zz, RC)


def dom_delta_Vs(dE, Ws0, ws1, ws2):
    """Analytic surface-dispersion correction of the DOM for a scalar energy offset.  optical_potentials/dom.py:399-455."""
    from scipy.special import exp1, expi

    if dE == 0.0:
        return 0.0
    x = float(dE)
    ax = abs(x)
    if ws2 == 0.0:
        J1 = 0.0
        J2 = np.pi * (ws1**2 + x**2) / (2.0 * np.sqrt(2.0) * ws1**3)
    else:
        z1 = ws2 * ax
        J1 = (-np.exp(-z1) * expi(z1) - np.exp(z1) * exp1(z1)) / (2.0 * ax)
        roots = ws1 * np.array([np.exp(1j * np.pi / 4), np.exp(1j * 3 * np.pi / 4)])
        F = np.exp(-ws2 * roots) * exp1(-ws2 * roots)
        J2 = 2.0 * np.real(np.sum((roots**2 + x**2) / (4.0 * roots**3) * F))
    pre1 = x**4 / (x**4 + ws1**4)
    pre2 = ws1**4 / (x**4 + ws1**4)
    return (2.0 * Ws0 * x / np.pi) * (pre1 * J1 + pre2 * J2)


def dom_shape_params(projectile, target, Ecm, Ef, p):
    """(central 11, spin-orbit 4, coulomb 2) numbers of the dispersive model.  optical_potentials/dom.py:287-382 with
    the depth functions :151-283."""
    (v0, v1, rv, av, wv0, wv1, rw, aw, ws0, ws1, ws2, rd, ad, vso0, wso0, wso1, rso, aso, rC) = p
    A, Z = target
    Ap, Zp = projectile
    A13 = A ** (1 / 3)
    R0, Rw, Rd, Rso, RC = rv * A13, rw * A13, rd * A13, rso * A13, rC * A13
    Ec = 6.0 * Z * Zp * ALPHA * HBARC / (5 * RC) if Zp == 1 else 0.0
    dE = Ecm - Ec - Ef
    V0 = v0 * np.exp(-v1 * dE)
    Vso = vso0 * np.exp(-v1 * dE)
    Wv = wv0 * dE**2 / (dE**2 + wv1**2)
    Ws = ws0 * dE**4 / (dE**4 + ws1**4) * np.exp(-ws2 * np.abs(dE))
    Wso = wso0 * dE**2 / (dE**2 + wso1**2)
    dVv = wv0 * wv1 * dE / (dE**2 + wv1**2)
    dVso = wso0 * wso1 * dE / (dE**2 + wso1**2)
    dVs = dom_delta_Vs(dE, ws0, ws1, ws2)
    return (V0, R0, av, Wv, dVv, Rw, aw, Ws, dVs, Rd, ad), (Vso + dVso, Wso, Rso, aso), (float(Z * Zp), RC)


def dom_potentials(r, projectile, target, Ecm, Ef, p):
    """(central, spin_orbit, coulomb) on r.  optical_potentials/dom.py:63-128, 492-529."""
    (V, R, a, Wv, dVv, Rw, aw, Ws, dVs, Rd, ad), (Vso, Wso, Rso, aso), (zz, RC) = dom_shape_params(projectile, target, Ecm, Ef, p)
    central = -V * woods_saxon_safe(r, R, a) - (dVv + 1j * Wv) * woods_saxon_safe(r, Rw, aw) \
        + (4 * ad * (dVs + 1j * Ws)) * woods_saxon_prime_safe(r, Rd, ad)
    spin_orbit = 2 * (Vso + 1j * Wso) * thomas_safe(r, Rso, aso)
    return central, spin_orbit, coulomb_charged_sphere(r, zz, RC)


def local_omp_potentials(r, A_target, zz, p, A_proj=None):
    """LocalOpticalPotential.  omp.py:54-149 (15 parameters)."""
    Vv, rv, av, Wv, rw, aw, Wd, Vd, rd, ad, Vso, Wso, rso, aso, rC = p
    rf = A_target ** (1 / 3) + (A_proj ** (1 / 3) if A_proj else 0.0)
    central = (
        -Vv * woods_saxon_safe(r, rv * rf, av)
        - 1j * Wv * woods_saxon_safe(r, rw * rf, aw)
        - (-4 * ad) * Vd * woods_saxon_prime_safe(r, rd * rf, ad)
        - 1j * (-4 * ad) * Wd * woods_saxon_prime_safe(r, rd * rf, ad)
    )
    so = (Vso + 1j * Wso) / WAVENUMBER_PION**2 * thomas_safe(r, rso * rf, aso)
    return central.astype(np.complex128), so.astype(np.complex128), coulomb_charged_sphere(r, zz, rC * rf)


# ----------------------------------------------------------------------------
# kinematics  (utils/kinematics.py:38-73)
# ----------------------------------------------------------------------------
def semi_relativistic_kinematics(m_t, m_p, Elab, Zz=0):
    Ecm = m_t / (m_t + m_p) * Elab
    Ep = Ecm + m_p
    k = m_t * np.sqrt(Elab * (Elab + 2 * m_p)) / np.sqrt((m_t + m_p) ** 2 + 2 * m_t * Elab) / HBARC
    mu = k**2 * Ep / (Ep**2 - m_p * m_p) * HBARC**2
    eta = (ALPHA * Zz * mu / HBARC) / k
    return Elab, Ecm, mu, k, eta


# ----------------------------------------------------------------------------
# spin-1/2 elastic workspace  (xs/elastic.py)
# ----------------------------------------------------------------------------
class ElasticWorkspace:
    """IntegralWorkspace + DifferentialWorkspace of the reference, one energy.
    xs/elastic.py:42-82 (setup), :104-183 (smatrix), :234-281 (angular tables),
    :332-381 (observables)."""

    def __init__(self, nbasis, lmax, channel_radius_fm, kin, angles, Zz, tol=1e-6, asym_table=None):
        from scipy.special import eval_legendre, gamma, lpmv

        Elab, Ecm, mu, k, eta = kin
        self.N, self.lmax, self.tol = nbasis, lmax, tol
        self.k, self.Ecm, self.eta, self.Elab = k, Ecm, eta, Elab
        self.a = channel_radius_fm * k
        x, _ = legendre_mesh(nbasis)
        self.rgrid = x * self.a / k
        self.free = [kinetic_bloch_matrix(nbasis, self.a, l) - np.eye(nbasis) for l in range(lmax + 1)]
        self.b = boundary_vector(nbasis)
        if asym_table is None:
            asym_table = np.array([coulomb_hankel(l, eta, self.a) for l in range(lmax + 1)])
        self.asym = asym_table  # (lmax+1, 4): Hp, Hm, Hpp, Hmp
        self._H = [[np.array([v], dtype=np.complex128) for v in row] for row in self.asym]
        self._one = np.ones(1, dtype=np.float64)
        self.angles = angles
        ls = np.arange(lmax + 1)[:, None]
        self.P = eval_legendre(ls, np.cos(angles))
        self.P1 = lpmv(1, ls, np.cos(angles))
        self.sigma_l = np.angle(gamma(1 + ls.astype(np.float64) + 1j * eta))
        if Zz > 0:
            s2 = np.sin(angles / 2.0)
            self.rutherford = 10 * eta**2 / (4 * k**2 * (s2**2) ** 2)
            self.f_c = -eta / (2 * k * s2**2) * np.exp(2j * self.sigma_l[0] - 2j * eta * np.log(s2))
        else:
            self.rutherford = None
            self.f_c = np.zeros_like(angles)

    def _solve(self, A, l):
        H = self._H[l]
        _, S, _, _ = solve_smatrix(A, self.b, H[0], H[1], H[2], H[3], self._one, self.a, 1, self.N)
        return S[0, 0]

    def smatrix(self, central, spin_orbit=None, coulomb=None, count=None):
        """xs/elastic.py:104-183, including the data-dependent early break (:178-181)."""
        N = self.N
        sp = np.zeros(self.lmax + 1, dtype=np.complex128)
        sm = np.zeros(self.lmax + 1, dtype=np.complex128)
        so = np.zeros(N, dtype=np.complex128) if spin_orbit is None else np.asarray(spin_orbit, dtype=np.complex128)
        im_c = np.diag(np.asarray(central, dtype=np.complex128)) / self.Ecm
        im_so = np.diag(so) / self.Ecm
        if coulomb is not None:
            im_c = im_c + np.diag(np.asarray(coulomb, dtype=np.complex128)) / self.Ecm
        sp[0] = self._solve(self.free[0] + im_c, 0)
        nsolve = 1
        last = 0
        for l in range(1, self.lmax + 1):
            cp, cm = spin_half_orbit_couplings(l)
            sp[l] = self._solve(self.free[l] + (im_c + cp * im_so), l)
            sm[l] = self._solve(self.free[l] + (im_c + cm * im_so), l)
            nsolve += 2
            last = l
            if abs(1 - sp[l]) < self.tol and abs(1 - sm[l]) < self.tol:
                break
        if count is not None:
            count[0] += nsolve
        return sp[: last + 1], sm[: last + 1]

    def xs(self, central, spin_orbit=None, coulomb=None, count=None):
        """DifferentialWorkspace.xs -> (dsdo, Ay, Q, sigma_tot, sigma_rxn).  xs/elastic.py:283-381."""
        sp, sm = self.smatrix(central, spin_orbit, coulomb, count)
        return differential_elastic_xs(self.k, self.angles, sp, sm, self.P, self.P1, self.f_c, self.sigma_l)


def differential_elastic_xs(k, angles, sp, sm, P, P1, f_c, sigma_l, eps=1e-30):
    """xs/elastic.py:332-381."""
    a = np.zeros_like(angles, dtype=np.complex128) + f_c
    b = np.zeros_like(angles, dtype=np.complex128)
    rxn = 0.0
    tot = 0.0
    for l in range(sp.shape[0]):
        phase = np.exp(2j * sigma_l[l]) / (2j * k)
        a += P[l, :] * phase * ((l + 1) * (sp[l] - 1) + l * (sm[l] - 1))
        b += P1[l, :] * phase * (sp[l] - sm[l])
        rxn += (l + 1) * (1 - np.abs(sp[l]) ** 2) + l * (1 - np.abs(sm[l]) ** 2)
        tot += (l + 1) * (1 - np.real(sp[l])) + l * (1 - np.real(sm[l]))
    d0 = np.abs(a) ** 2 + np.abs(b) ** 2
    den = np.maximum(d0, eps)
    Ay = 2.0 * np.imag(np.conjugate(a) * b) / den
    Q = 2.0 * np.real(np.conjugate(a) * b) / den
    return d0 * 10.0, Ay, Q, tot * 10.0 * 2.0 * np.pi / k**2, rxn * 10.0 * np.pi / k**2


# ----------------------------------------------------------------------------
# quasi-elastic (p,n) in the DWBA  (xs/quasielastic_pn.py)
# ----------------------------------------------------------------------------
def exit_kinematics(kin_entrance, masses, Q, Ex_residual=0.0, Ex_product=0.0, Zz_exit=0):
    """reactions/reaction.py:580-623.  masses = (target, projectile, product, residual) rest masses."""
    _, Ecm_in, _, _, _ = kin_entrance
    _, _, m_prod, m_res = masses
    Ecm = Ecm_in + Q - Ex_residual - Ex_product
    Elab = (m_res + m_prod) / m_res * Ecm
    return semi_relativistic_kinematics(m_res + Ex_residual, m_prod, Elab, Zz=Zz_exit)


class QuasielasticWorkspace:
    """xs/quasielastic_pn.py:83-421: distorted waves of the entrance (p) and exit (n) channel with
    wavefunction=True, T_lj = sum_n x_p U_1 x_n / (R k_p k_n) (:317-322), early break on |T| (:336-340),
    angular sum with the precomputed geometric factors (:183-217, :375-392)."""

    def __init__(self, nbasis, lmax, channel_radius_fm, kin_entrance, kin_exit, angles, A_target, Z_target,
                 Zz_entrance, Zz_exit=0, tol=1e-6):
        from scipy.special import gamma, sph_harm_y
        from sympy.physics.wigner import clebsch_gordan

        self.N, self.lmax, self.R, self.tol = nbasis, lmax, channel_radius_fm, tol
        self.kin_p, self.kin_n = tuple(kin_entrance), tuple(kin_exit)
        _, self.Ep, self.mup, self.kp, self.etap = self.kin_p
        _, self.En, self.mun, self.kn, self.etan = self.kin_n
        self.ap, self.an = channel_radius_fm * self.kp, channel_radius_fm * self.kn
        self.angles = np.asarray(angles, dtype=np.float64)
        Nn = A_target - Z_target
        self.isovector_factor = np.sqrt(np.fabs(Nn - Z_target)) / (Nn - Z_target - 1)  # :144-147
        eye = np.eye(nbasis)
        self.free_p = [kinetic_bloch_matrix(nbasis, self.ap, l) - eye for l in range(lmax + 1)]  # :150-163
        self.free_n = [kinetic_bloch_matrix(nbasis, self.an, l) - eye for l in range(lmax + 1)]
        self.b = boundary_vector(nbasis)
        self.asym_p = [coulomb_hankel(l, self.etap, self.ap) for l in range(lmax + 1)]
        self.asym_n = [coulomb_hankel(l, self.etan, self.an) for l in range(lmax + 1)]
        x, _ = legendre_mesh(nbasis)
        self.rgrid = x * self.ap / self.kp  # :219-223
        # :186-217
        self.xs_factor = (self.kn / self.kp) * self.mup * self.mun / (4 * np.pi**2 * HBARC**4 * (2 * 1.0 / 2 + 1))
        ls = np.arange(lmax + 1)
        self.sigma_c = np.angle(gamma(1 + ls + 1j * self.etap))
        G = np.zeros((2, 2, lmax + 1, 2, self.angles.shape[0]), dtype=np.complex128)
        for im, m in enumerate([-0.5, 0.5]):
            for imp, mp in enumerate([-0.5, 0.5]):
                for l in range(lmax + 1):
                    for ijp, jp in enumerate([l + 1 / 2, l - 1 / 2] if l > 0 else [l + 1 / 2]):
                        if abs(m - mp) <= l and jp >= 0:
                            ylm = sph_harm_y(l, int(m - mp), self.angles, 0)
                            cg0 = clebsch_gordan(l, 1 / 2, jp, m - mp, m, mp)
                            cg1 = clebsch_gordan(l, 1 / 2, jp, 0, m, m)
                            G[im, imp, l, ijp, :] = (
                                (4 * np.pi) ** (3.0 / 2.0) / (self.kp * self.kn) * np.exp(1j * self.sigma_c[l])
                                * complex(cg1) * complex(cg0) * np.sqrt(2 * l + 1) * (-1) ** (2 * jp + 1) * ylm
                            )
        self.geometric_factor = G

    def _wave(self, A, asym, a):
        one = np.ones(1, dtype=np.float64)
        H = [np.array([v], dtype=np.complex128) for v in asym]
        _, S, Ainv, up = solve_smatrix(A, self.b, H[0], H[1], H[2], H[3], one, a, 1, self.N)
        return S[0, 0], solution_coeffs(Ainv, self.b, up, 1, self.N)[0]

    def tmatrix(self, U_p_coulomb, U_p_central, U_p_spin_orbit=None, U_n_central=None, U_n_spin_orbit=None):
        """:225-342.  Returns (Tpn, Sn, Sp), each (lmax + 1, 2)."""
        N, L1 = self.N, self.lmax + 1
        z = np.zeros(N, dtype=np.complex128)
        pc, pcoul = np.asarray(U_p_central, dtype=np.complex128), np.asarray(U_p_coulomb, dtype=np.complex128)
        pso = z if U_p_spin_orbit is None else np.asarray(U_p_spin_orbit, dtype=np.complex128)
        nc = np.asarray(U_n_central, dtype=np.complex128)
        nso = z if U_n_spin_orbit is None else np.asarray(U_n_spin_orbit, dtype=np.complex128)
        U1c = -(nc - pc) * self.isovector_factor
        U1so = -(nso - pso) * self.isovector_factor
        T = np.zeros((L1, 2), dtype=np.complex128)
        Sn = np.zeros((L1, 2), dtype=np.complex128)
        Sp = np.zeros((L1, 2), dtype=np.complex128)

        def element(l, lds):
            sn, xn = self._wave(self.free_n[l] + np.diag(nc + lds * nso) / self.En, self.asym_n[l], self.an)
            sp, xp = self._wave(self.free_p[l] + np.diag(pc + pcoul + lds * pso) / self.Ep, self.asym_p[l], self.ap)
            t = np.sum(xp * (U1c + lds * U1so) * xn) / self.R / self.kp / self.kn
            return t, sn, sp

        T[0, 0], Sn[0, 0], Sp[0, 0] = element(0, 0.0)
        for l in range(1, L1):
            lds = spin_half_orbit_couplings(l)
            T[l, 0], Sn[l, 0], Sp[l, 0] = element(l, lds[0])
            T[l, 1], Sn[l, 1], Sp[l, 1] = element(l, lds[1])
            if abs(T[l, 0]) < self.tol and abs(T[l, 1]) < self.tol:
                break
        return T, Sn, Sp

    def xs(self, *potentials):
        """:344-392 (the angular sum stops at lmax - 1, as the reference's does)."""
        T, _, _ = self.tmatrix(*potentials)
        Tmmp = np.zeros((2, 2, self.angles.shape[0]), dtype=np.complex128)
        for im, m in enumerate([-0.5, 0.5]):
            for imp, mp in enumerate([-0.5, 0.5]):
                for l in range(0, self.lmax):
                    for ijp, jp in enumerate([l + 0.5, l - 0.5]):
                        if abs(m - mp) <= l and jp >= 0:
                            Tmmp[im, imp, :] += self.geometric_factor[im, imp, l, ijp, :] * T[l, ijp]
        return self.xs_factor * 10 * np.sum(np.absolute(Tmmp) ** 2, axis=(0, 1))


def interaction_range(A, rA=1.2, r0=0.0):
    """utils/utils.py:118-120."""
    return r0 + rA * A ** (1 / 3)


def suggested_channel_radius_fm(A, k):
    """(interaction_range(A) k + 2 pi) / k, as kduq_uq_demo.ipynb cell 5 (utils/utils.py:123-129)."""
    return (interaction_range(A) * k + 2 * np.pi) / k
