"""Per-partial-wave channel metadata and Coulomb-Hankel asymptotics at the channel radius.

Same classes, fields and call signatures as reactions/system.py:18-227 of the reference; the
asymptotics come from ``free_solutions.coulomb_hankel_table`` (one table per (lmax, eta, a)
instead of eight mpmath calls per (l, channel)).
"""
from __future__ import annotations

import numpy as np

from ..free_solutions import coulomb_hankel_table


def scalar_couplings(l):
    """Default single-channel coupling matrix (reactions/system.py:18-20)."""
    return np.array([[1.0]])


def spin_half_orbit_coupling(l):
    """l.sigma expectation values for spin-1/2 on spin-0 (reactions/system.py:23-33)."""
    js = [l + 1.0 / 2] if l == 0 else [l + 1.0 / 2, l - 1.0 / 2]
    return np.diag([(j * (j + 1) - l * (l + 1) - 0.5 * (0.5 + 1)) for j in js])


class Asymptotics:
    """H+, H-, H+', H-' for the channels of one partial wave (reactions/system.py:36-64)."""

    def __init__(self, Hp, Hm, Hpp, Hmp):
        self.Hp, self.Hm, self.Hpp, self.Hmp = Hp, Hm, Hpp, Hmp
        self.size = Hp.shape[0]

    def decouple(self):
        return [
            Asymptotics(self.Hp[i : i + 1], self.Hm[i : i + 1], self.Hpp[i : i + 1], self.Hmp[i : i + 1])
            for i in range(self.size)
        ]


class Channels:
    """Channel metadata for one partial wave (reactions/system.py:67-113)."""

    def __init__(self, E, k, mu, eta, a, l, couplings):
        self.num_channels = E.shape[0]
        assert couplings.shape == (self.num_channels, self.num_channels)
        assert k.shape[0] == self.num_channels
        assert mu.shape[0] == self.num_channels
        assert eta.shape[0] == self.num_channels
        self.E, self.k, self.mu, self.eta, self.a, self.l = E, k, mu, eta, a, l
        self.size = couplings.shape[0]
        self.couplings = couplings

    def decouple(self):
        assert np.count_nonzero(self.couplings - np.diag(np.diagonal(self.couplings))) == 0
        d = np.diag(self.couplings)
        return [
            Channels(
                self.E[i : i + 1], self.k[i : i + 1], self.mu[i : i + 1], self.eta[i : i + 1],
                self.a, self.l[i : i + 1], np.array([[d[i]]]),
            )
            for i in range(self.size)
        ]


class ProjectileTargetSystem:
    """System-level information of a projectile-target partition (reactions/system.py:116-215)."""

    def __init__(
        self, channel_radius, lmax, mass_target=0.0, mass_projectile=0.0, Ztarget=0.0, Zproj=0.0,
        coupling=scalar_couplings, channel_levels=None,
    ):
        self.channel_radius = channel_radius
        self.lmax = lmax
        self.l = np.arange(0, lmax + 1, dtype=np.int64)
        self.couplings = [coupling(int(l)) for l in self.l]
        if channel_levels is None:
            channel_levels = np.zeros(self.couplings[0].shape[0], dtype=np.float64)
        self.channel_levels = channel_levels
        self.mass_target = mass_target
        self.mass_projectile = mass_projectile
        self.Ztarget = Ztarget
        self.Zproj = Zproj

    def get_partial_wave_channels(self, Elab, Ecm, mu, k, eta):
        """(list[Channels], list[Asymptotics]) for l = 0..lmax (reactions/system.py:157-215)."""
        channels, asymptotics = [], []
        tables = {}
        for l in range(0, self.lmax + 1):
            nch = self.couplings[l].shape[0]
            eta_array = uniform_array_from_scalar_or_array(eta, nch)
            channels.append(
                Channels(
                    uniform_array_from_scalar_or_array(Ecm, nch),
                    uniform_array_from_scalar_or_array(k, nch),
                    uniform_array_from_scalar_or_array(mu, nch),
                    eta_array,
                    self.channel_radius,
                    np.ones(nch) * l,
                    self.couplings[l],
                )
            )
            rows = []
            for ce in eta_array:
                ce = float(ce)
                if ce not in tables:
                    tables[ce] = coulomb_hankel_table(self.lmax, ce, float(self.channel_radius))
                rows.append(tables[ce][l])
            rows = np.array(rows, dtype=np.complex128)  # (nch, 4)
            asymptotics.append(Asymptotics(Hp=rows[:, 0].copy(), Hm=rows[:, 1].copy(), Hpp=rows[:, 2].copy(), Hmp=rows[:, 3].copy()))
        return channels, asymptotics

    def asymptotics_table(self, eta):
        """Packed (lmax+1, 4) complex128 table [H+, H-, H+', H-'] for a scalar eta (device upload)."""
        return coulomb_hankel_table(self.lmax, float(eta), float(self.channel_radius))


def uniform_array_from_scalar_or_array(scalar_or_array, size):
    """reactions/system.py:218-227."""
    if isinstance(scalar_or_array, (np.ndarray, list)):
        return np.asarray(scalar_or_array, dtype=np.float64)
    return np.full(size, scalar_or_array, dtype=np.float64)
