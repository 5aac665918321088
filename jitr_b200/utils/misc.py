"""utils/utils.py: block slicing and scalar helpers used around the path."""
from __future__ import annotations

import numpy as np

from ..free_solutions import CoulombAsymptotics, H_minus, H_minus_prime, H_plus, H_plus_prime


def block(matrix, block_index, block_size):
    """Sub-block view of a block-structured matrix (utils/utils.py:31-49)."""
    i, j = block_index
    n, m = block_size
    return matrix[i * n : i * n + n, j * m : j * m + m]


def smatrix(Rl, a, l, eta, asym=CoulombAsymptotics):
    """S-matrix element from a channel R-matrix value (utils/utils.py:78-88)."""
    return (H_minus(a, l, eta, asym=asym) - a * Rl * H_minus_prime(a, l, eta, asym=asym)) / (
        H_plus(a, l, eta, asym=asym) - a * Rl * H_plus_prime(a, l, eta, asym=asym)
    )


def delta(Sl):
    """Phase shift and attenuation in degrees (utils/utils.py:91-94)."""
    phase = np.log(Sl) / 2.0j
    return np.rad2deg(np.real(phase)), np.rad2deg(np.imag(phase))


def interaction_range(A, rA=1.2, r0=0.0):
    """utils/utils.py:118-120."""
    return r0 + rA * A ** (1 / 3)


def suggested_dimensionless_channel_radius(interaction_range, k):
    """utils/utils.py:123-129."""
    return interaction_range * k + 2 * np.pi


def suggested_basis_size(a, zeros_per_node=5):
    """utils/utils.py:132-134."""
    return zeros_per_node * int(np.ceil(a / np.pi))
