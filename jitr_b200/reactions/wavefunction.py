"""Internal/external wavefunctions rebuilt from the solve outputs (reactions/wavefunction.py:15-82)."""
from __future__ import annotations

import numpy as np

from ..free_solutions import CoulombAsymptotics, H_minus, H_plus


class Wavefunctions:
    def __init__(self, solver, coeffs, S, uext_prime_boundary, channels, incoming_weights=None, asym=CoulombAsymptotics):
        self.solver = solver
        self.coeffs = coeffs
        self.S = S
        self.uext_prime_boundary = uext_prime_boundary
        self.channels = channels
        if incoming_weights is None:
            incoming_weights = np.zeros(channels.size, dtype=np.float64)
            incoming_weights[0] = 1
        self.incoming_weights = incoming_weights
        self.asym = asym

    def uext(self):
        """External wavefunctions beyond the boundary.  (The reference's version calls
        ``len(channels)`` on an object without ``__len__`` -- reactions/wavefunction.py:56,66 --
        and raises; ``channels.size`` is what it means.)"""
        nch = self.channels.size

        def uext_channel(i):
            l = self.channels.l[i]
            eta = self.channels.eta[i]

            def one(s):
                inc = self.incoming_weights[i] * H_minus(s, l, eta, asym=self.asym)
                out = sum(self.incoming_weights[j] * self.S[i, j] * H_plus(s, l, eta, asym=self.asym) for j in range(nch))
                return 1j / 2 * (inc - out)

            return lambda s_mesh: np.array([one(s) for s in s_mesh], dtype=np.complex128)

        return [uext_channel(i) for i in range(nch)]

    def uint(self):
        """Internal wavefunctions in the Lagrange basis (reactions/wavefunction.py:68-82)."""
        nb = self.solver.kernel.quadrature.nbasis

        def uint_channel(i):
            return lambda s: np.sum(
                [self.coeffs[i, n] / self.channels.a * self.solver.kernel.f(n + 1, self.channels.a, s) for n in range(nb)],
                axis=0,
            )

        return [uint_channel(i) for i in range(self.channels.size)]
