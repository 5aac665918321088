"""jitr_b200: B200-native calculable R-matrix hot path behind jitr's Solver / workspace API.

    from jitr_b200 import rmatrix, reactions, xs, optical_potentials, utils

Host code is Python; every solve runs in hand-written sm_100a CUDA kernels behind the C ABI in
``include/jitr_b200.h`` (``jitr_b200/_lib/libjitr_b200.so``).  There is no CPU fallback.
"""
from . import constants, free_solutions, optical_potentials, quadrature, reactions, rmatrix, uq, utils, xs
from .rmatrix import Solver
from .xs.elastic import DifferentialWorkspace, ElasticSweep, ElasticXS, IntegralWorkspace

__version__ = "0.1.0"
__all__ = [
    "DifferentialWorkspace", "ElasticSweep", "ElasticXS", "IntegralWorkspace", "Solver", "constants",
    "free_solutions", "optical_potentials", "quadrature", "reactions", "rmatrix", "uq", "utils", "xs",
]
