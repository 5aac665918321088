"""Entrance-channel kinematics (scalar host math feeding the path).  utils/kinematics.py:16-121."""
from __future__ import annotations

from dataclasses import astuple, dataclass

import numpy as np

from ..constants import ALPHA, HBARC


@dataclass
class ChannelKinematics:
    """(Elab, Ecm, mu, k, eta) of one channel; iterable in field order (utils/kinematics.py:16-35)."""

    Elab: float
    Ecm: float
    mu: float
    k: float
    eta: float

    def __iter__(self):
        return iter(astuple(self))


def semi_relativistic_kinematics(mass_target, mass_projectile, Elab, Zz=0):
    """Ingemarsson (1974) approximation, utils/kinematics.py:38-73."""
    m_t, m_p = mass_target, mass_projectile
    Ecm = m_t / (m_t + m_p) * Elab
    Ep = Ecm + m_p
    k = m_t * np.sqrt(Elab * (Elab + 2 * m_p)) / np.sqrt((m_t + m_p) ** 2 + 2 * m_t * Elab) / HBARC
    mu = k**2 * Ep / (Ep**2 - m_p * m_p) * HBARC**2
    eta = (ALPHA * Zz * mu / HBARC) / k
    return ChannelKinematics(Elab, Ecm, mu, k, eta)


def classical_kinematics(mass_target, mass_projectile, Elab, Zz=0):
    """utils/kinematics.py:76-97."""
    mu = mass_target * mass_projectile / (mass_target + mass_projectile)
    Ecm = mass_target / (mass_target + mass_projectile) * Elab
    k = np.sqrt(2 * Ecm * mu) / HBARC
    eta = (ALPHA * Zz) * mu / (HBARC * k)
    return ChannelKinematics(Elab, Ecm, mu, k, eta)


def classical_kinematics_cm(mass_target, mass_projectile, Ecm, Zz=0):
    """utils/kinematics.py:100-121."""
    mu = mass_target * mass_projectile / (mass_target + mass_projectile)
    Elab = (mass_target + mass_projectile) / mass_target * Ecm
    k = np.sqrt(2 * Ecm * mu) / HBARC
    eta = (ALPHA * Zz) * mu / (HBARC * k)
    return ChannelKinematics(Elab, Ecm, mu, k, eta)
