"""R-matrix solver on a Lagrange-Legendre mesh (device-resident batched solve behind the reference API)."""
from .rmatrix import Solver

__all__ = ["Solver"]
