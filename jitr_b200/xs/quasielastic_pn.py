"""DWBA workspace for quasi-elastic (p,n) observables -- host mirror of the reference's
``xs/quasielastic_pn.py`` (System :18-80, Workspace :83-421) on the CUDA path.

The distorted waves of the entrance (proton) and exit (neutron) channel are solved with
``jitr_b200_solve_batched`` (wavefunction coefficients on), for all samples and both j of one l per launch;
``jitr_b200_qe_overlap`` forms T_lj, ``jitr_b200_qe_xs`` the angular sum.  ``tmatrix`` / ``xs`` keep the
reference's single-sample numpy signatures (B = 1 round trip); ``tmatrix_batch`` / ``xs_batch`` take CUDA tensors
with a leading sample axis.  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from .. import _cabi
from ..constants import HBARC
from ..reactions.system import ProjectileTargetSystem, spin_half_orbit_coupling
from ..rmatrix import Solver
from .elastic import check_angles


def _cg_half(l, ml, ms, j, mj):
    """<l ml; 1/2 ms | j mj> in closed form (Condon-Shortley), 0 when ml + ms != mj or |mj| > j."""
    if abs(ml + ms - mj) > 1e-12 or abs(mj) > j + 1e-12 or abs(ml) > l or j < 0:
        return 0.0
    if abs(j - (l + 0.5)) < 1e-12:
        return float(np.sqrt((l + 0.5 + (mj if ms > 0 else -mj)) / (2 * l + 1)))
    if abs(j - (l - 0.5)) < 1e-12:
        return float((-1.0 if ms > 0 else 1.0) * np.sqrt((l + 0.5 - (mj if ms > 0 else -mj)) / (2 * l + 1)))
    return 0.0


class System:
    """Entrance and exit partitions of a (p,n) reaction (quasielastic_pn.py:18-80)."""

    def __init__(self, channel_radius_fm, lmax, reaction, kinematics_entrance, kinematics_exit):
        self.channel_radius_fm = channel_radius_fm
        self.lmax = lmax
        self.l = np.arange(0, lmax + 1, dtype=np.int64)
        self.entrance = ProjectileTargetSystem(
            channel_radius=channel_radius_fm * kinematics_entrance.k, lmax=lmax, mass_target=reaction.target.m0,
            mass_projectile=reaction.projectile.m0, Ztarget=reaction.target.Z, Zproj=reaction.projectile.Z,
            coupling=spin_half_orbit_coupling,
        )
        if getattr(reaction, "residual", None) is None or getattr(reaction, "product", None) is None:
            raise ValueError("Reaction must define both residual and product for (p,n) scattering")
        self.exit = ProjectileTargetSystem(
            channel_radius=channel_radius_fm * kinematics_exit.k, lmax=lmax, mass_target=reaction.residual.m0,
            mass_projectile=reaction.product.m0, Ztarget=reaction.residual.Z, Zproj=reaction.product.Z,
            coupling=spin_half_orbit_coupling,
        )


class Workspace:
    """quasielastic_pn.py:83-421."""

    def __init__(self, reaction, kinematics_entrance, kinematics_exit, solver: Solver, angles, lmax, channel_radius_fm,
                 tmatrix_abs_tol=1e-6):
        from scipy.special import gamma, sph_harm_y

        self.lmax = lmax
        self.channel_radius_fm = channel_radius_fm
        self.tmatrix_abs_tol = tmatrix_abs_tol
        self.reaction = reaction
        self.sys = System(channel_radius_fm, lmax, reaction, kinematics_entrance, kinematics_exit)
        self.kinematics_entrance = kinematics_entrance
        self.kinematics_exit = kinematics_exit
        self.solver = solver
        check_angles(angles)
        self.angles = np.asarray(angles, dtype=np.float64)
        A, Z = reaction.target.A, reaction.target.Z
        Nn = A - Z
        self.isovector_factor = np.sqrt(np.fabs(Nn - Z)) / (Nn - Z - 1)
        # free matrices and asymptotics of both channels, all l (host, once)
        q = solver.kernel.quadrature
        eye = np.eye(q.nbasis)
        self.a_p, self.a_n = self.sys.entrance.channel_radius, self.sys.exit.channel_radius
        self.free_matrices_p = np.array([q.kinetic_matrix(self.a_p, int(l)) - eye for l in self.sys.l])
        self.free_matrices_n = np.array([q.kinetic_matrix(self.a_n, int(l)) - eye for l in self.sys.l])
        self.asym_p = self.sys.entrance.asymptotics_table(kinematics_entrance.eta)  # (lmax+1, 4)
        self.asym_n = self.sys.exit.asymptotics_table(kinematics_exit.eta)
        self.l_dot_s = np.array([np.diag(c) for c in self.sys.entrance.couplings[1:]])
        # purely geometric factors (:186-217)
        ke, kx = kinematics_entrance, kinematics_exit
        self.xs_factor = (kx.k / ke.k) * ke.mu * kx.mu / (4 * np.pi**2 * HBARC**4 * (2 * 1.0 / 2 + 1))
        self.sigma_c = np.angle(gamma(1 + self.sys.l + 1j * ke.eta))
        # G[m, m', l, j', theta] of :198-217.  The reference's Clebsch-Gordan product <l m-m' 1/2 m | j' m'><l 0 1/2 m | j' m>
        # vanishes unless m = m' (its first factor needs (m - m') + m = m'), and (-1)^(2j'+1) = +1 for half-integer j':
        # only the spin-diagonal entries are non-zero, each the square of one closed-form coefficient times Y_l0.
        G = np.zeros((2, 2, lmax + 1, 2, self.angles.shape[0]), dtype=np.complex128)
        pref = (4 * np.pi) ** 1.5 / (ke.k * kx.k)
        for l in range(lmax + 1):
            yl0 = sph_harm_y(l, 0, self.angles, 0)
            common = pref * np.exp(1j * self.sigma_c[l]) * np.sqrt(2 * l + 1) * yl0
            for im, m in enumerate((-0.5, 0.5)):
                G[im, im, l, 0] = common * _cg_half(l, 0, m, l + 0.5, m) ** 2
                if l > 0:
                    G[im, im, l, 1] = common * _cg_half(l, 0, m, l - 0.5, m) ** 2
        self.geometric_factor = G
        self._dev = {}

    # ------------------------------------------------------------------ host helpers
    def radial_grid(self):
        return self.solver.radial_grid(self.a_p, self.kinematics_entrance.k)

    def _local_potential(self, potential, name):
        v = np.asarray(potential, dtype=np.complex128)
        expected = (self.solver.kernel.quadrature.nbasis,)
        if v.shape != expected:
            raise ValueError(f"{name} must have shape {expected}")
        return v

    def _optional_local_potential(self, potential, name):
        if potential is None:
            return np.zeros(self.solver.kernel.quadrature.nbasis, dtype=np.complex128)
        return self._local_potential(potential, name)

    def _tables(self, dev):
        import torch

        key = str(dev)
        if key not in self._dev:
            self._dev[key] = dict(
                Fp=torch.tensor(self.free_matrices_p, dtype=torch.complex128, device=dev),
                Fn=torch.tensor(self.free_matrices_n, dtype=torch.complex128, device=dev),
                ap=torch.tensor(self.asym_p, dtype=torch.complex128, device=dev),
                an=torch.tensor(self.asym_n, dtype=torch.complex128, device=dev),
                G=torch.tensor(self.geometric_factor, dtype=torch.complex128, device=dev).contiguous(),
            )
        return self._dev[key]

    # ------------------------------------------------------------------ device path
    def tmatrix_batch(self, U_p_coulomb, U_p_central, U_p_spin_orbit=None, U_n_central=None, U_n_spin_orbit=None):
        """Batched ``tmatrix``: potentials are (B, N) complex128 CUDA tensors.  Returns CUDA tensors
        (Tpn, Sn, Sp) of shape (B, lmax + 1, 2) -- zero beyond each sample's early break -- and status (B,)."""
        import torch

        _cabi.require_device()
        if U_n_central is None:
            raise TypeError("U_n_central is required")
        N = self.solver.kernel.quadrature.nbasis
        dev = U_p_central.device
        t = self._tables(dev)

        def prep(v, name, optional=False):
            if v is None:
                if optional:
                    return None
                raise TypeError(f"{name} is required")
            v = v.to(device=dev, dtype=torch.complex128)
            if v.dim() != 2 or v.shape[1] != N:
                raise ValueError(f"{name} must have shape (B, {N})")
            return v.contiguous()

        pc, pcoul = prep(U_p_central, "U_p_central"), prep(U_p_coulomb, "U_p_coulomb")
        nc = prep(U_n_central, "U_n_central")
        pso, nso = prep(U_p_spin_orbit, "U_p_spin_orbit", True), prep(U_n_spin_orbit, "U_n_spin_orbit", True)
        B, L1 = pc.shape[0], self.lmax + 1
        zero = torch.zeros_like(pc)
        pso_z = zero if pso is None else pso
        nso_z = zero if nso is None else nso
        U1c = (-(nc - pc) * self.isovector_factor).contiguous()
        U1so = None if (pso is None and nso is None) else (-(nso_z - pso_z) * self.isovector_factor).contiguous()
        pcen = pc + pcoul
        ke, kx = self.kinematics_entrance, self.kinematics_exit
        scale = 1.0 / self.channel_radius_fm / ke.k / kx.k
        T = torch.zeros((B, L1, 2), dtype=torch.complex128, device=dev)
        Sn, Sp = torch.zeros_like(T), torch.zeros_like(T)
        status = torch.zeros((B,), dtype=torch.int32, device=dev)
        done = torch.zeros((B,), dtype=torch.bool, device=dev)
        L = _cabi.lib()
        for l in range(L1):
            lds = (0.0, 0.0) if l == 0 else tuple(float(v) for v in self.l_dot_s[l - 1])
            nj = 1 if l == 0 else 2
            if nj == 1:
                vn, vp = nc, pcen
            else:
                vn = torch.cat([nc + lds[0] * nso_z, nc + lds[1] * nso_z], dim=0)
                vp = torch.cat([pcen + lds[0] * pso_z, pcen + lds[1] * pso_z], dim=0)
            rn = self.solver.solve_batch(self.a_n, kx.Ecm, kx.k, 1, t["Fn"][l], t["an"][l].reshape(4, 1), local_potential=vn, wavefunction=True)
            rp = self.solver.solve_batch(self.a_p, ke.Ecm, ke.k, 1, t["Fp"][l], t["ap"][l].reshape(4, 1), local_potential=vp, wavefunction=True)
            Tl = torch.zeros((B, 2), dtype=torch.complex128, device=dev)
            rc = L.jitr_b200_qe_overlap(
                N, B, nj, _cabi.ptr(rp["coeffs"]), _cabi.ptr(rn["coeffs"]), _cabi.ptr(U1c), _cabi.ptr(U1so), lds[0], lds[1],
                scale, _cabi.ptr(Tl), 2, 1, _cabi.stream_ptr(),
            )
            _cabi.check(rc, "jitr_b200_qe_overlap")
            live = ~done
            snl = rn["S"][:, 0, 0].reshape(nj, B).transpose(0, 1)
            spl = rp["S"][:, 0, 0].reshape(nj, B).transpose(0, 1)
            st = torch.maximum(rn["status"].reshape(nj, B).max(dim=0).values, rp["status"].reshape(nj, B).max(dim=0).values)
            T[:, l, :] = torch.where(live[:, None], Tl, torch.zeros_like(Tl))
            Sn[:, l, :nj] = torch.where(live[:, None], snl, torch.zeros_like(snl))
            Sp[:, l, :nj] = torch.where(live[:, None], spl, torch.zeros_like(spl))
            status = torch.maximum(status, torch.where(live, st, torch.zeros_like(st)))
            if l >= 1:
                # the reference leaves its l loop when both |T_lj| fall below the tolerance (:336-340)
                done = done | (live & (Tl[:, 0].abs() < self.tmatrix_abs_tol) & (Tl[:, 1].abs() < self.tmatrix_abs_tol))
                if bool(done.all()):
                    break
        return T, Sn, Sp, status

    def xs_batch(self, U_p_coulomb, U_p_central, U_p_spin_orbit=None, U_n_central=None, U_n_spin_orbit=None, return_status=False):
        """Batched ``xs``: (B, n_angles) float64 CUDA tensor in mb/sr."""
        import torch

        T, _, _, status = self.tmatrix_batch(U_p_coulomb, U_p_central, U_p_spin_orbit, U_n_central, U_n_spin_orbit)
        B, nth = T.shape[0], self.angles.shape[0]
        out = torch.empty((B, nth), dtype=torch.float64, device=T.device)
        rc = _cabi.lib().jitr_b200_qe_xs(
            self.lmax, nth, B, _cabi.ptr(T), _cabi.ptr(self._tables(T.device)["G"]), float(self.xs_factor * 10), _cabi.ptr(out),
            _cabi.stream_ptr(),
        )
        _cabi.check(rc, "jitr_b200_qe_xs")
        return (out, status) if return_status else out

    # ------------------------------------------------------------------ the reference's single-sample API
    def _up(self, U_p_coulomb, U_p_central, U_p_spin_orbit, U_n_central, U_n_spin_orbit):
        import torch

        _cabi.require_device()
        if U_n_central is None:
            raise TypeError("U_n_central is required")
        dev = torch.device("cuda", torch.cuda.current_device())

        def up(v):
            return torch.tensor(v[None, :], dtype=torch.complex128, device=dev)

        return (
            up(self._local_potential(U_p_coulomb, "U_p_coulomb")), up(self._local_potential(U_p_central, "U_p_central")),
            None if U_p_spin_orbit is None else up(self._local_potential(U_p_spin_orbit, "U_p_spin_orbit")),
            up(self._local_potential(U_n_central, "U_n_central")),
            None if U_n_spin_orbit is None else up(self._local_potential(U_n_spin_orbit, "U_n_spin_orbit")),
        )

    def tmatrix(self, U_p_coulomb, U_p_central, U_p_spin_orbit=None, U_n_central=None, U_n_spin_orbit=None):
        """(Tpn, Sn, Sp), each (lmax + 1, 2) complex numpy arrays (:225-342)."""
        from .elastic import _raise_on_status

        T, Sn, Sp, status = self.tmatrix_batch(*self._up(U_p_coulomb, U_p_central, U_p_spin_orbit, U_n_central, U_n_spin_orbit))
        _raise_on_status(status[:1])
        return T[0].cpu().numpy(), Sn[0].cpu().numpy(), Sp[0].cpu().numpy()

    def xs(self, U_p_coulomb, U_p_central, U_p_spin_orbit=None, U_n_central=None, U_n_spin_orbit=None):
        """Differential (p,n) cross section in mb/sr at ``self.angles`` (:344-392)."""
        from .elastic import _raise_on_status

        out, status = self.xs_batch(*self._up(U_p_coulomb, U_p_central, U_p_spin_orbit, U_n_central, U_n_spin_orbit), return_status=True)
        _raise_on_status(status[:1])
        return out[0].cpu().numpy()
