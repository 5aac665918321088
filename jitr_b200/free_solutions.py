"""Coulomb-Hankel asymptotics H+-(l, eta, s) and their derivatives at the channel radius.

Same definitions as the reference (utils/free_solutions.py:39-109): H+- = G +- iF and
H' = ((l+1)/s + eta/(l+1)) H_l - sqrt(1 + eta^2/(l+1)^2) H_{l+1}.  Setup-side, host code.

Differences in HOW (not what): a whole table l = 0..lmax is produced from F, G at l = 0..lmax+1
(2(lmax+2) special-function evaluations instead of the reference's 12 per l); for eta = 0 the
Riccati-Bessel closed form s j_l(s), -s y_l(s) replaces the mpmath calls; for eta > 0 the table comes from
Steed's continued fractions (``coulomb_fg_steed``, ~0.3 ms instead of ~1 s per energy) wherever they hold
1e-13 -- at or above the l = 0 turning point, judged a posteriori by |G_0 / F_0| -- and from mpmath below it.
"""
from __future__ import annotations

import numpy as np
import scipy.special as sc


class FreeAsymptotics:
    """Riccati-Bessel functions (neutral particles).  utils/free_solutions.py:25-36."""

    @staticmethod
    def F(s, l, _eta=None):
        return s * sc.spherical_jn(l, s)

    @staticmethod
    def G(s, l, _eta=None):
        return -s * sc.spherical_yn(l, s)


class CoulombAsymptotics:
    """Coulomb wave functions through mpmath (53-bit), utils/free_solutions.py:39-50."""

    @staticmethod
    def F(s, l, eta):
        from mpmath import coulombf

        return np.complex128(coulombf(l, eta, s))

    @staticmethod
    def G(s, l, eta):
        from mpmath import coulombg

        return np.complex128(coulombg(l, eta, s))


def H_plus(s, l, eta, asym=CoulombAsymptotics):
    return asym.G(s, l, eta) + 1j * asym.F(s, l, eta)


def H_minus(s, l, eta, asym=CoulombAsymptotics):
    return asym.G(s, l, eta) - 1j * asym.F(s, l, eta)


def coulomb_func_deriv(func, s, l, eta):
    recurrence_factor = np.sqrt(1 + eta**2 / (l + 1) ** 2)
    shift_term = (l + 1) / s + eta / (l + 1)
    return shift_term * func(s, l, eta) - recurrence_factor * func(s, l + 1, eta)


def H_plus_prime(s, l, eta, asym=CoulombAsymptotics):
    return coulomb_func_deriv(lambda ss, ll, ee: H_plus(ss, ll, ee, asym=asym), s, l, eta)


def H_minus_prime(s, l, eta, dx=1e-6, asym=CoulombAsymptotics):
    return coulomb_func_deriv(lambda ss, ll, ee: H_minus(ss, ll, ee, asym=asym), s, l, eta)


def coulomb_fg_steed(lmax, eta, rho, acc=1e-16, maxit=100000):
    """Regular and irregular Coulomb functions F_l, G_l and their rho-derivatives for l = 0..lmax at rho > 0.

    Steed's method (Barnett, Comput. Phys. Commun. 27 (1982) 147): F'/F at l = lmax from the continued fraction
    CF1 (modified Lentz, sign of F from the denominators), downward recurrence to l = 0, p + iq = H+'/H+ at
    l = 0 from CF2, normalisation by the Wronskian, upward recurrence for G.  Relative accuracy ~1e-14 at and
    above the l = 0 turning point rho = 2 eta; below it the normalisation loses about |G_0 / F_0| ulps.
    Returns (F, G, F', G') or None when a continued fraction does not converge."""
    eta, rho, L = float(eta), float(rho), int(lmax)
    tiny = 1e-300
    xi = 1.0 / rho

    def S(k):
        return k * xi + eta / k

    def R2(k):
        return 1.0 + (eta / k) ** 2

    k = L + 1
    f = S(k) or tiny
    C, D, sign = f, 0.0, 1.0
    for _ in range(maxit):
        a, b = -R2(k), S(k) + S(k + 1)
        D = (b + a * D) or tiny
        C = (b + a / C) or tiny
        D = 1.0 / D
        if D < 0.0:
            sign = -sign
        delta = C * D
        f *= delta
        k += 1
        if abs(delta - 1.0) < acc:
            break
    else:
        return None
    F, Fp = np.zeros(L + 1), np.zeros(L + 1)
    F[L] = sign * 1e-60
    Fp[L] = f * F[L]
    for l in range(L, 0, -1):
        Sl, Rl = S(l), np.sqrt(R2(l))
        F[l - 1] = (Sl * F[l] + Fp[l]) / Rl
        Fp[l - 1] = Sl * F[l - 1] - Rl * F[l]
        if abs(F[l - 1]) > 1e250:
            F[l - 1 :] *= 1e-250
            Fp[l - 1 :] *= 1e-250
    f0 = Fp[0] / F[0]
    # CF2 at l = 0 in real arithmetic (a b = (i eta)(i eta + 1))
    P, Q = 0.0, 1.0 - eta * xi
    ar, ai = -(eta * eta), eta
    br, bi = 2.0 * (rho - eta), 2.0
    den = br * br + bi * bi
    dr, di = br / den, -bi / den
    dp, dq = -xi * (ar * di + ai * dr), xi * (ar * dr - ai * di)
    pk = 0.0
    for _ in range(maxit):
        P += dp
        Q += dq
        pk += 2.0
        ar += pk
        ai += 2.0 * eta
        bi += 2.0
        d_r = ar * dr - ai * di + br
        d_i = ai * dr + ar * di + bi
        c = 1.0 / (d_r * d_r + d_i * d_i)
        dr, di = c * d_r, -c * d_i
        a_, b_ = br * dr - bi * di - 1.0, bi * dr + br * di
        dp, dq = dp * a_ - dq * b_, dp * b_ + dq * a_
        if abs(dp) + abs(dq) < (abs(P) + abs(Q)) * acc:
            break
    else:
        return None
    P += dp
    Q += dq
    gam = (f0 - P) / Q
    w2 = (f0 - P) * gam + Q
    if not (w2 > 0.0 and np.isfinite(w2)):
        return None
    F0 = np.copysign(1.0 / np.sqrt(w2), F[0])
    scale = F0 / F[0]
    F *= scale
    Fp *= scale
    G, Gp = np.zeros(L + 1), np.zeros(L + 1)
    G[0] = gam * F0
    Gp[0] = (P * gam - Q) * F0
    for l in range(L):
        Sl, Rl = S(l + 1), np.sqrt(R2(l + 1))
        G[l + 1] = (Sl * G[l] - Gp[l]) / Rl
        Gp[l + 1] = Rl * G[l] - Sl * G[l + 1]
    return F, G, Fp, Gp


def _taylor_step(l, eta, rho0, u0, up0, h, nterms=48):
    """One step of the Coulomb equation  u'' = (l(l+1)/rho^2 + 2 eta/rho - 1) u  from rho0 to rho0 + h by its Taylor series at
    rho0 (|h| <= rho0/4: the terms fall off like 4^-n).  Returns (u, u') at rho0 + h."""
    L = l * (l + 1.0)
    c = np.zeros(nterms + 2)
    c[0], c[1] = u0, up0
    k0, k1 = L + 2.0 * eta * rho0 - rho0 * rho0, 2.0 * eta - 2.0 * rho0
    r2 = rho0 * rho0
    for m in range(nterms):
        acc = (k0 - m * (m - 1.0)) * c[m] - 2.0 * rho0 * (m + 1.0) * m * c[m + 1]
        if m >= 1:
            acc += k1 * c[m - 1]
        if m >= 2:
            acc -= c[m - 2]
        c[m + 2] = acc / (r2 * (m + 2.0) * (m + 1.0))
    u = up = 0.0
    for n in range(nterms + 1, -1, -1):  # Horner
        u = u * h + c[n]
    for n in range(nterms + 1, 0, -1):
        up = up * h + n * c[n]
    return u, up


def coulomb_fg_below_turning_point(lmax, eta, rho):
    """F_l, G_l, F_l', G_l' for l = 0..lmax at 0 < rho < 2 eta (inside the l = 0 turning point), without multi-precision
    arithmetic.  Steed's normalisation cancels there (it loses |G_0/F_0| ulps), so:
      1. Steed's method at rho_1 = 4 eta + 2, comfortably outside, gives G_0, G_0' to ~1e-15;
      2. G_0 is carried INWARD to rho by Taylor-series steps of the Coulomb equation (steps of min(rho/4, 1), 48 terms): the
         irregular solution grows inward, so the integration is stable and keeps the relative accuracy;
      3. F_0'/F_0 at rho from the continued fraction CF1 + downward recurrence (stable for the regular solution), and the
         Wronskian F'G - FG' = 1 gives  F_0 = 1 / (f_0 G_0 - G_0');
      4. G_l by upward recurrence (stable for the irregular solution), F_l from the downward-recurrence ratios.
    Returns (F, G, F', G') or None."""
    eta, rho, Lm = float(eta), float(rho), int(lmax)
    rho1 = 4.0 * eta + 2.0
    if not (0.0 < rho < rho1):
        return None
    st = coulomb_fg_steed(0, eta, rho1)
    if st is None:
        return None
    g, gp = float(st[1][0]), float(st[3][0])
    r = rho1
    while r > rho:
        h = max(rho - r, -min(0.25 * r, 1.0))  # |h| <= 1: no cancellation between the terms where the solution oscillates
        g, gp = _taylor_step(0, eta, r, g, gp, h)
        r += h
    if not (np.isfinite(g) and np.isfinite(gp)):
        return None
    # CF1 at lmax, downward recurrence: F up to a common factor
    xi = 1.0 / rho
    S = lambda k: k * xi + eta / k  # noqa: E731
    R2 = lambda k: 1.0 + (eta / k) ** 2  # noqa: E731
    tiny = 1e-300
    k = Lm + 1
    f = S(k) or tiny
    C, D = f, 0.0
    for _ in range(200000):
        a_, b_ = -R2(k), S(k) + S(k + 1)
        D = (b_ + a_ * D) or tiny
        C = (b_ + a_ / C) or tiny
        D = 1.0 / D
        delta = C * D
        f *= delta
        k += 1
        if abs(delta - 1.0) < 1e-16:
            break
    else:
        return None
    F, Fp = np.zeros(Lm + 1), np.zeros(Lm + 1)
    F[Lm] = 1e-60
    Fp[Lm] = f * F[Lm]
    for l in range(Lm, 0, -1):
        Sl, Rl = S(l), np.sqrt(R2(l))
        F[l - 1] = (Sl * F[l] + Fp[l]) / Rl
        Fp[l - 1] = Sl * F[l - 1] - Rl * F[l]
        if abs(F[l - 1]) > 1e250:
            F[l - 1:] *= 1e-250
            Fp[l - 1:] *= 1e-250
    f0 = Fp[0] / F[0]
    F0 = 1.0 / (f0 * g - gp)
    scale = F0 / F[0]
    F *= scale
    Fp *= scale
    G, Gp = np.zeros(Lm + 1), np.zeros(Lm + 1)
    G[0], Gp[0] = g, gp
    for l in range(Lm):
        Sl, Rl = S(l + 1), np.sqrt(R2(l + 1))
        G[l + 1] = (Sl * G[l] - Gp[l]) / Rl
        Gp[l + 1] = Rl * G[l] - Sl * G[l + 1]
    if not (np.all(np.isfinite(F)) and np.all(np.isfinite(G))):
        return None
    return F, G, Fp, Gp


# Steed's normalisation loses about |G_0 / F_0| ulps INSIDE the turning point: there the inward integration of G_0 + Wronskian
# route (coulomb_fg_below_turning_point) is used; mpmath only if a routine fails (overflow far inside the barrier)

_TABLE_CACHE: dict = {}


def coulomb_hankel_table(lmax, eta, s):
    """(lmax+1, 4) complex128 table [H+, H-, H+', H-'] for l = 0..lmax at s (cached)."""
    key = (int(lmax), float(eta), float(s))
    if key in _TABLE_CACHE:
        return _TABLE_CACHE[key]
    ls = np.arange(lmax + 2)
    if eta == 0.0:
        F = s * sc.spherical_jn(ls, s)
        G = -s * sc.spherical_yn(ls, s)
    else:
        # at or above the l = 0 turning point rho = 2 eta Steed's method holds ~1e-14 of |H| everywhere (also at the nodes of
        # F_0, where |G_0/F_0| is large for a harmless reason); inside it the inward integration of G_0 + Wronskian takes over
        fg = coulomb_fg_steed(lmax + 1, eta, s) if (eta > 0.0 and s >= 2.0 * eta) else None
        if fg is not None and np.all(np.isfinite(fg[0])) and np.all(np.isfinite(fg[1])):
            F, G = fg[0], fg[1]
        elif eta > 0.0 and s < 2.0 * eta and (fb := coulomb_fg_below_turning_point(lmax + 1, eta, s)) is not None:
            F, G = fb[0], fb[1]
        else:
            F = np.array([float(np.real(CoulombAsymptotics.F(s, int(l), eta))) for l in ls])
            G = np.array([float(np.real(CoulombAsymptotics.G(s, int(l), eta))) for l in ls])
    Hp = G + 1j * F
    Hm = G - 1j * F
    l = ls[:-1]
    rec = np.sqrt(1 + eta**2 / (l + 1) ** 2)
    shift = (l + 1) / s + eta / (l + 1)
    out = np.stack([Hp[:-1], Hm[:-1], shift * Hp[:-1] - rec * Hp[1:], shift * Hm[:-1] - rec * Hm[1:]], axis=1)
    out = np.ascontiguousarray(out.astype(np.complex128))
    _TABLE_CACHE[key] = out
    return out
