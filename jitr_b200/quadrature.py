"""Lagrange-Legendre mesh: nodes, basis functions, kinetic+Bloch operator, quadrature Kernel.

Host-side (float64 numpy) setup data of the R-matrix path; mirrors the Legendre part of the
reference's ``quadrature`` package (quadrature/quadrature.py:37-55,66-74,151-229;
quadrature/kernel.py:27-67,180-214) with the same names and argument meaning.  These arrays are
computed once per (nbasis, a) and uploaded to the device by ``Solver`` / the workspaces.
"""
from __future__ import annotations

import numpy as np
from scipy.special import eval_legendre


def generate_legendre_quadrature(nbasis):
    """Gauss-Legendre zeros/weights shifted to (0, 1).  quadrature/quadrature.py:66-74."""
    x, w = np.polynomial.legendre.leggauss(nbasis)
    return 0.5 * (x + 1), 0.5 * w


def legendre(n, a, s, quadrature):
    """n-th (1-based) regularised Lagrange-Legendre function at s.  quadrature/quadrature.py:37-55."""
    assert 1 <= n <= quadrature.nbasis
    N = quadrature.nbasis
    x = s / a
    xn = quadrature.abscissa[n - 1]
    return (-1.0) ** (N - n) * np.sqrt((1 - xn) / xn) * eval_legendre(N, 2.0 * x - 1.0) * x / (x - xn)


class LagrangeLegendreQuadrature:
    """quadrature/quadrature.py:151-229."""

    def __init__(self, abscissa, weights, overlap=None):
        self.nbasis = len(abscissa)
        assert len(abscissa) == len(weights)
        self.abscissa = abscissa
        self.weights = weights
        self.overlap = np.diag(np.ones(self.nbasis)) if overlap is None else overlap
        self._that = None

    def kinetic_operator_element(self, n, m, a, l):
        """(n, m) element (1-based) of kinetic + Bloch operator (Baye 2015 eq. 3.128/3.129)."""
        assert 1 <= n <= self.nbasis and 1 <= m <= self.nbasis
        xn, xm = self.abscissa[n - 1], self.abscissa[m - 1]
        N = self.nbasis
        if n == m:
            centrifugal = l * (l + 1) / (a * xn) ** 2
            radial = ((4 * N**2 + 4 * N + 3) * xn * (1 - xn) - 6 * xn + 1) / (3 * xn**2 * (1 - xn) ** 2) / a**2
            return radial + centrifugal
        return (
            (-1.0) ** (n + m)
            * ((N**2 + N + 1.0) + (xn + xm - 2 * xn * xm) / (xn - xm) ** 2 - 1.0 / (1.0 - xn) - 1.0 / (1.0 - xm))
            / np.sqrt(xn * xm * (1.0 - xn) * (1.0 - xm))
            / a**2
        )

    def kinetic_hat(self):
        """a-independent, l = 0 operator That (real N x N): T(l, a) = That / a^2 + diag(l(l+1)/(a x)^2).

        Vectorised evaluation of the same element formulas; the upper triangle is mirrored like
        quadrature/quadrature.py:224-229."""
        if self._that is None:
            x = self.abscissa
            N = self.nbasis
            xn, xm = x[:, None], x[None, :]
            sign = (-1.0) ** (np.add.outer(np.arange(1, N + 1), np.arange(1, N + 1)))
            with np.errstate(divide="ignore", invalid="ignore"):
                off = (
                    sign
                    * ((N**2 + N + 1.0) + (xn + xm - 2 * xn * xm) / (xn - xm) ** 2 - 1.0 / (1.0 - xn) - 1.0 / (1.0 - xm))
                    / np.sqrt(xn * xm * (1.0 - xn) * (1.0 - xm))
                )
            T = np.triu(off, k=1)
            T = T + T.T
            T[np.diag_indices(N)] = ((4 * N**2 + 4 * N + 3) * x * (1 - x) - 6 * x + 1) / (3 * x**2 * (1 - x) ** 2)
            self._that = T
        return self._that

    def kinetic_matrix(self, a, l):
        """Kinetic operator matrix T(l, a), complex128 like the reference (:220-229)."""
        F = (self.kinetic_hat() / a**2).astype(np.complex128)
        F[np.diag_indices(self.nbasis)] += l * (l + 1) / (a * self.abscissa) ** 2
        return F


class Kernel:
    """Quadrature rule + basis functions (Legendre mesh).  quadrature/kernel.py:27-67,180-214."""

    def __init__(self, nbasis, basis="Legendre"):
        self.overlap = np.diag(np.ones(nbasis))
        if basis != "Legendre":
            raise NotImplementedError("jitr_b200 implements the Lagrange-Legendre mesh (the R-matrix hot path) only")
        x, w = generate_legendre_quadrature(nbasis)
        self.quadrature = LagrangeLegendreQuadrature(x, w)
        self.basis_function = legendre
        self.weight_matrix = np.outer(w, w)
        self.Xn, self.Xm = np.meshgrid(x, x)

    def f(self, n, a, s):
        return self.basis_function(n, a, s, self.quadrature)

    def radial_grid(self, radius):
        return self.quadrature.abscissa * radius

    def nonlocal_radial_grids(self, radius):
        return self.Xn * radius, self.Xm * radius

    def matrix_local(self, values):
        values_array = np.asarray(values, dtype=np.complex128)
        nbasis = self.quadrature.nbasis
        if values_array.ndim == 1 and values_array.shape == (nbasis,):
            return values_array
        if values_array.ndim == 3 and values_array.shape[0] == values_array.shape[1] and values_array.shape[2] == nbasis:
            return values_array
        raise ValueError(f"local potential values must have shape ({nbasis},) or (nchannels, nchannels, {nbasis})")

    def matrix_nonlocal(self, values, radius):
        values_array = np.asarray(values, dtype=np.complex128)
        nbasis = self.quadrature.nbasis
        weight_sqrt = np.sqrt(self.weight_matrix)
        if values_array.ndim == 2 and values_array.shape == (nbasis, nbasis):
            return weight_sqrt * values_array * radius
        if values_array.ndim == 4 and values_array.shape[0] == values_array.shape[1] and values_array.shape[2:] == (nbasis, nbasis):
            return weight_sqrt[np.newaxis, np.newaxis, :, :] * values_array * radius
        raise ValueError(
            f"nonlocal potential values must have shape ({nbasis}, {nbasis}) or (nchannels, nchannels, {nbasis}, {nbasis})"
        )
