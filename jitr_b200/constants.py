"""Physical constants (values of the reference, utils/constants.py:5-12)."""
import numpy as np

HBARC = 197.3269804
ALPHA = 1.0 / 137.0359991
WAVENUMBER_PION = np.sqrt(1 / 2)
AMU = 931.494102
MASS_N = 1.008665 * AMU
MASS_P = 1.007276 * AMU
MASS_E = 0.5109989461
C = 2.99792458e23
