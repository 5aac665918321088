"""Coulomb-Hankel asymptotics H+-(l, eta, s) and their derivatives at the channel radius.

Same definitions as the reference (utils/free_solutions.py:39-109): H+- = G +- iF and
H' = ((l+1)/s + eta/(l+1)) H_l - sqrt(1 + eta^2/(l+1)^2) H_{l+1}.  Setup-side, host code.

Differences in HOW (not what): a whole table l = 0..lmax is produced from F, G at l = 0..lmax+1
(2(lmax+2) special-function evaluations instead of the reference's 12 per l), and for eta = 0 the
Riccati-Bessel closed form s j_l(s), -s y_l(s) replaces the mpmath calls.
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
