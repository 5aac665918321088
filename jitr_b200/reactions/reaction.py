"""Minimal reaction descriptors.

The reference's ``reactions/reaction.py`` (957 lines: mass tables, Q-values, EXFOR strings) is
host bookkeeping outside the hot path.  The path only consumes
``reaction.target.{A, Z, m0}``, ``reaction.projectile.{A, Z, m0}``, ``reaction.process`` and
``reaction.kinematics(Elab)`` (xs/elastic.py:52-70, kduq.py:548-553), so these light classes keep
exactly that surface; rest masses are given explicitly (MeV/c^2) instead of looked up.  Any
object with the same attributes -- including the reference's own ``Reaction`` -- works as well.
"""
from __future__ import annotations

from dataclasses import dataclass

from ..constants import AMU, MASS_N, MASS_P
from ..utils.kinematics import semi_relativistic_kinematics


@dataclass(frozen=True)
class Particle:
    A: int
    Z: int
    m0: float


def _particle(p, mass):
    if isinstance(p, Particle):
        return p
    A, Z = p
    if mass is None:
        if (A, Z) == (1, 0):
            mass = MASS_N
        elif (A, Z) == (1, 1):
            mass = MASS_P
        else:
            # the reference looks the rest mass up in its mass table (reactions/reaction.py:167): same table, same default
            from ..data import rest_mass

            mass = rest_mass(A, Z)
            if mass is None:
                raise ValueError(f"no tabulated rest mass for (A, Z) = ({A}, {Z}): pass the mass explicitly (mass_target=...)")
    return Particle(int(A), int(Z), float(mass))


class Reaction:
    """Two-body reaction ``target(projectile, product)residual`` (reactions/reaction.py:343-623).  ``product`` and
    ``residual`` are only needed for exit-channel kinematics (quasi-elastic (p,n)); rest masses are passed
    explicitly, the Q value follows from them unless given."""

    def __init__(self, target, projectile, product=None, residual=None, process=None, mass_target=None,
                 mass_projectile=None, mass_product=None, mass_residual=None, Q=None, Ef=None):
        self.target = _particle(target, mass_target)
        self.projectile = _particle(projectile, mass_projectile)
        self.product = None if product is None else _particle(product, mass_product)
        self.residual = None if residual is None else _particle(residual, mass_residual)
        self.process = process
        if Q is None and self.product is not None and self.residual is not None:
            Q = self.target.m0 + self.projectile.m0 - self.product.m0 - self.residual.m0
        self.Q = Q
        self.Ef = Ef  # effective Fermi energy (reactions/reaction.py:528-538), used by the dispersive model only

    def kinematics(self, Elab):
        """reactions/reaction.py:540-556."""
        return semi_relativistic_kinematics(
            self.target.m0, self.projectile.m0, Elab, Zz=self.target.Z * self.projectile.Z
        )

    def kinematics_exit(self, entrance, residual_excitation_energy=0, product_excitation_energy=0):
        """Exit-channel kinematics from the entrance channel's and the excitation energies.
        reactions/reaction.py:580-623."""
        if self.Q is None:
            raise ValueError("kinematics_exit() requires a defined Q-value.")
        if self.residual is None or self.product is None:
            raise ValueError("kinematics_exit() requires both product and residual to be defined.")
        Ecm = entrance.Ecm + self.Q - residual_excitation_energy - product_excitation_energy
        Elab = (self.residual.m0 + self.product.m0) / self.residual.m0 * Ecm
        return semi_relativistic_kinematics(
            self.residual.m0 + residual_excitation_energy, self.product.m0, Elab, Zz=self.residual.Z * self.product.Z
        )


class ElasticReaction(Reaction):
    """reactions/reaction.py:715-738."""

    def __init__(self, target, projectile, mass_target=None, mass_projectile=None, Ef=None):
        super().__init__(target, projectile, process="el", mass_target=mass_target, mass_projectile=mass_projectile, Ef=Ef)
