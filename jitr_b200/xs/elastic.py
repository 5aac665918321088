"""Elastic-scattering observables built from the batched R-matrix solve.

Keeps the reference's workspace API (xs/elastic.py:20-389): ``ElasticXS``, ``IntegralWorkspace``,
``DifferentialWorkspace`` with the same constructor arguments, attributes and single-sample
methods (numpy in, numpy out) -- and adds the batched forms the UQ sweeps use:

* ``*_batch(central, spin_orbit, coulomb)``: user potentials as CUDA tensors ``(B, N)``;
* ``*_params(model, params)``: built-in potentials (KDUQ / CHUQ / LocalOpticalPotential / canonical
  shape numbers) evaluated on the mesh inside the fused kernel from ``(B, n_params)`` rows.

A workspace is one energy; ``ElasticSweep`` owns one workspace per energy and runs all
(sample, energy) pairs in one launch.  All numerics run on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.special import eval_legendre, gamma, lpmv

from .. import _cabi
from ..reactions import ProjectileTargetSystem, spin_half_orbit_coupling
from ..rmatrix import Solver


@dataclass
class ElasticXS:
    """dsdo [mb/sr], Ay, Q, total and reaction cross sections [mb] (xs/elastic.py:20-36).

    numpy arrays/scalars for the single-sample API, CUDA tensors with leading (B[, E]) axes for the
    batched API."""

    dsdo: object
    Ay: object
    Q: object
    t: object
    rxn: object


def check_angles(angles):
    """xs/elastic.py:384-389."""
    if angles.ndim != 1:
        raise ValueError("angles must be a 1D array")
    if angles[0] < 0 or angles[-1] > np.pi:
        raise ValueError("angles must a grid in radians on [0,pi)")


def energy_table_row(reaction, kinematics, a, radius_factor=None):
    """One row of the per-energy device table (layout: include/jitr_b200.h JITR_B200_ET_*)."""
    row = np.zeros(_cabi.ET_STRIDE, dtype=np.float64)
    A, Z = reaction.target.A, reaction.target.Z
    Ap, Zp = reaction.projectile.A, reaction.projectile.Z
    row[_cabi.ET_A] = a
    row[_cabi.ET_ECM] = kinematics.Ecm
    row[_cabi.ET_K] = kinematics.k
    row[_cabi.ET_ELAB] = kinematics.Elab
    row[_cabi.ET_RCH] = a / kinematics.k  # Solver.radial_grid: abscissa * (a / k0), rmatrix.py:26-28
    row[_cabi.ET_AT], row[_cabi.ET_ZT], row[_cabi.ET_ZP], row[_cabi.ET_AP] = A, Z, Zp, Ap
    row[_cabi.ET_RSCALE] = A ** (1 / 3) if radius_factor is None else radius_factor
    row[_cabi.ET_A13] = A ** (1.0 / 3.0)
    row[_cabi.ET_AM13] = A ** (-1.0 / 3.0)
    row[_cabi.ET_AM23] = A ** (-2.0 / 3.0)
    row[_cabi.ET_AM53] = A ** (-5.0 / 3.0)
    Ef = getattr(reaction, "Ef", None)
    row[_cabi.ET_EF] = 0.0 if Ef is None else float(Ef)  # effective Fermi energy (dispersive model)
    return row


class IntegralWorkspace:
    """Workspace for integral elastic observables with spin-orbit coupling (xs/elastic.py:39-211)."""

    def __init__(self, reaction, kinematics, channel_radius_fm, solver: Solver, lmax, smatrix_abs_tol=1e-6):
        if reaction.process is None or reaction.process.lower() != "el":
            raise ValueError("Reaction must be elastic!")
        self.smatrix_abs_tol = smatrix_abs_tol
        self.lmax = lmax
        self.channel_radius_fm = channel_radius_fm
        self.a = channel_radius_fm * kinematics.k
        self.kinematics = kinematics
        self.reaction = reaction
        self.sys = ProjectileTargetSystem(
            self.a, lmax, mass_target=reaction.target.m0, mass_projectile=reaction.projectile.m0,
            Ztarget=reaction.target.Z, Zproj=reaction.projectile.Z, coupling=spin_half_orbit_coupling,
        )
        self.solver = solver
        self.basis_boundary = solver.precompute_boundaries(self.a)
        channels, asymptotics = self.sys.get_partial_wave_channels(*self.kinematics)
        self.channels = [c.decouple() for c in channels]
        self.asymptotics = [a.decouple() for a in asymptotics]
        self.l_dot_s = np.array([np.diag(c) for c in self.sys.couplings[1:]])
        self.ls = self.sys.l[:, np.newaxis]
        self._free = None
        self._asym_table = self.sys.asymptotics_table(kinematics.eta)
        self._sweep = None

    @property
    def free_matrices(self):
        """Per-l free matrices T(l) - 1 (xs/elastic.py:73), built lazily: the device path never needs
        them (it assembles from That and the diagonal), they are kept for API parity."""
        if self._free is None:
            q = self.solver.kernel.quadrature
            eye = np.eye(q.nbasis)
            self._free = [q.kinetic_matrix(self.a, int(l)) - eye for l in self.sys.l]
        return self._free

    def radial_grid(self):
        return self.solver.radial_grid(self.a, self.kinematics.k)

    # -- validation (xs/elastic.py:84-102)
    def _local_potential(self, potential, name):
        arr = np.asarray(potential, dtype=np.complex128)
        expected = (self.solver.kernel.quadrature.nbasis,)
        if arr.shape != expected:
            raise ValueError(f"{name} must have shape {expected}")
        return arr

    def _optional_local_potential(self, potential, name):
        if potential is None:
            return None
        return self._local_potential(potential, name)

    def _as_sweep(self):
        if self._sweep is None:
            self._sweep = ElasticSweep([self])
        return self._sweep

    # -- single-sample API (numpy in, numpy out)
    def smatrix(self, central_potential, spin_orbit_potential=None, coulomb_potential=None):
        """(S+[:L], S-[:L]) for j = l +- 1/2 with the early break (xs/elastic.py:104-183)."""
        import torch

        c = self._local_potential(central_potential, "central_potential")
        so = self._optional_local_potential(spin_orbit_potential, "spin_orbit_potential")
        co = self._optional_local_potential(coulomb_potential, "coulomb_potential") if coulomb_potential is not None else None

        def up(v):
            return None if v is None else torch.from_numpy(np.ascontiguousarray(v)).cuda().reshape(1, 1, -1)

        S, nl, status = self._as_sweep().smatrix_tensors(up(c), up(so), up(co))
        _raise_on_status(status)
        L = int(nl.item())
        Sh = S[0, 0].cpu().numpy()
        return Sh[0, :L].copy(), Sh[1, :L].copy()

    def xs(self, central_potential, spin_orbit_potential=None, coulomb_potential=None):
        """(sigma_tot, sigma_rxn) in mb (xs/elastic.py:185-197, :310-329)."""
        sp, sm = self.smatrix(central_potential, spin_orbit_potential, coulomb_potential)
        return integral_elastic_xs(self.kinematics.k, sp, sm, self.ls)

    def transmission_coefficients(self, central_potential, spin_orbit_potential=None, coulomb_potential=None):
        """xs/elastic.py:199-211."""
        sp, sm = self.smatrix(central_potential, spin_orbit_potential, coulomb_potential)
        return 1.0 - np.absolute(sp) ** 2, 1.0 - np.absolute(sm) ** 2

    # -- batched API (CUDA tensors)
    def smatrix_batch(self, central, spin_orbit=None, coulomb=None):
        """central/spin_orbit/coulomb: (B, N) complex128 CUDA tensors -> S (B, 2, lmax+1), nl (B,), status (B,)."""
        def e1(v):
            return None if v is None else v.unsqueeze(1)

        S, nl, status = self._as_sweep().smatrix_tensors(e1(central), e1(spin_orbit), e1(coulomb))
        return S[:, 0], nl[:, 0], status[:, 0]

    def smatrix_params(self, model, params):
        """params: (B, n_params) rows of ``model`` -> S (B, 2, lmax+1), nl (B,), status (B,)."""
        S, nl, status = self._as_sweep().smatrix_params(model, params)
        return S[:, 0], nl[:, 0], status[:, 0]


def _raise_on_status(status):
    bad = int((status != _cabi.STATUS_OK).sum().item())
    if bad:
        raise np.linalg.LinAlgError(f"{bad} system(s) singular or non-finite")


def integral_elastic_xs(k, Splus, Sminus, ls=None):
    """(sigma_tot, sigma_rxn) from truncated S-matrices on the host (xs/elastic.py:310-329); tiny sums."""
    l = np.arange(Splus.shape[0])
    rxn = np.sum((l + 1) * (1 - np.absolute(Splus) ** 2) + l * (1 - np.absolute(Sminus) ** 2))
    tot = np.sum((l + 1) * (1 - np.real(Splus)) + l * (1 - np.real(Sminus)))
    return tot * 10 * 2 * np.pi / k**2, rxn * 10 * np.pi / k**2


class DifferentialWorkspace:
    """Workspace for angular elastic observables (xs/elastic.py:214-307)."""

    @classmethod
    def build_from_system(cls, reaction, kinematics, channel_radius_fm, solver, lmax, angles, smatrix_abs_tol=1e-6):
        return cls(IntegralWorkspace(reaction, kinematics, channel_radius_fm, solver, lmax, smatrix_abs_tol), angles)

    def __init__(self, integral_workspace, angles):
        self.integral_workspace = integral_workspace
        self.reaction = integral_workspace.reaction
        self.kinematics = integral_workspace.kinematics
        check_angles(angles)
        self.angles = angles
        self.ls = integral_workspace.ls
        self.P_l_costheta = eval_legendre(self.ls, np.cos(self.angles))
        self.P_1_l_costheta = lpmv(1, self.ls, np.cos(self.angles))
        self.Zz = self.reaction.projectile.Z * self.reaction.target.Z
        self.sigma_l = self.coulomb_phase_shift(self.ls.astype(np.float64))
        if self.Zz > 0:
            self.rutherford = self.rutherford_xs(self.angles)
            self.f_c = self.coulomb_amplitude(self.angles, self.sigma_l[0])
        else:
            self.f_c = np.zeros_like(angles)
            self.rutherford = None
        self._sweep = None

    def radial_grid(self):
        return self.integral_workspace.radial_grid()

    def rutherford_xs(self, angles):
        check_angles(angles)
        sin2 = np.sin(angles / 2.0) ** 2
        return 10 * self.kinematics.eta**2 / (4 * self.kinematics.k**2 * sin2**2)

    def coulomb_amplitude(self, angles, sigma_0):
        sin2 = np.sin(angles / 2.0)
        return np.asarray(
            -self.kinematics.eta / (2 * self.kinematics.k * sin2**2)
            * np.exp(2j * sigma_0 - 2j * self.kinematics.eta * np.log(sin2)),
            dtype=np.complex128,
        )

    def coulomb_phase_shift(self, ls):
        return np.angle(gamma(1 + ls + 1j * self.kinematics.eta))

    def _as_sweep(self):
        if self._sweep is None:
            self._sweep = ElasticSweep([self])
        return self._sweep

    def xs(self, central_potential, spin_orbit_potential=None, coulomb_potential=None):
        """Single sample: ElasticXS of numpy arrays (xs/elastic.py:283-307)."""
        import torch

        iw = self.integral_workspace
        c = iw._local_potential(central_potential, "central_potential")
        so = iw._optional_local_potential(spin_orbit_potential, "spin_orbit_potential")
        co = iw._local_potential(coulomb_potential, "coulomb_potential") if coulomb_potential is not None else None

        def up(v):
            return None if v is None else torch.from_numpy(np.ascontiguousarray(v)).cuda().reshape(1, 1, -1)

        x, status = self._as_sweep().xs_tensors(up(c), up(so), up(co), return_status=True)
        _raise_on_status(status)
        return ElasticXS(
            x.dsdo[0, 0].cpu().numpy(), x.Ay[0, 0].cpu().numpy(), x.Q[0, 0].cpu().numpy(),
            np.float64(x.t[0, 0].item()), np.float64(x.rxn[0, 0].item()),
        )

    def xs_batch(self, central, spin_orbit=None, coulomb=None):
        """(B, N) CUDA tensors -> ElasticXS of CUDA tensors with leading axis B."""
        def e1(v):
            return None if v is None else v.unsqueeze(1)

        x = self._as_sweep().xs_tensors(e1(central), e1(spin_orbit), e1(coulomb))
        return ElasticXS(x.dsdo[:, 0], x.Ay[:, 0], x.Q[:, 0], x.t[:, 0], x.rxn[:, 0])

    def xs_params(self, model, params):
        """(B, n_params) parameter rows of a built-in model -> ElasticXS of CUDA tensors (leading axis B)."""
        x = self._as_sweep().xs_params(model, params)
        return ElasticXS(x.dsdo[:, 0], x.Ay[:, 0], x.Q[:, 0], x.t[:, 0], x.rxn[:, 0])


def _on_device(fn):
    """Run a method of ElasticSweep with the sweep's device current: its tables live there, and the C ABI launches on the
    current device / torch's current stream of that device."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        import torch

        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)

    return wrapper


class ElasticSweep:
    """All (sample, energy) pairs of a UQ sweep in one fused launch.

    ``workspaces``: one ``DifferentialWorkspace`` (or ``IntegralWorkspace`` for S-matrix-only use) per
    energy, all sharing nbasis, lmax, tolerance and -- for observables -- the angle grid.  The
    per-energy tables (channel radius, kinematics, H+- for every l, Legendre tables, Coulomb
    amplitude) are packed once into small replicated device tensors.
    """

    def __init__(self, workspaces, device=None):
        import torch

        _cabi.require_device()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.diff = [w for w in workspaces if isinstance(w, DifferentialWorkspace)]
        self.integral = [w.integral_workspace if isinstance(w, DifferentialWorkspace) else w for w in workspaces]
        iw0 = self.integral[0]
        self.solver = iw0.solver
        self.N = iw0.solver.kernel.quadrature.nbasis
        self.lmax = iw0.lmax
        self.tol = iw0.smatrix_abs_tol
        self.nE = len(self.integral)
        for iw in self.integral:
            if iw.solver.kernel.quadrature.nbasis != self.N or iw.lmax != self.lmax or iw.smatrix_abs_tol != self.tol:
                raise ValueError("all workspaces of a sweep must share nbasis, lmax and smatrix_abs_tol")
        self.tabs = self.solver.device_tables(self.device)
        self._free_cache = {}
        etab = np.stack([energy_table_row(iw.reaction, iw.kinematics, iw.a) for iw in self.integral])
        asym = np.stack([iw._asym_table for iw in self.integral])  # (E, L1, 4)
        t = lambda v, dt: torch.tensor(np.ascontiguousarray(v), dtype=dt, device=self.device)
        self.etab = t(etab, torch.float64)
        self.asym = t(asym, torch.complex128)
        self.has_obs = len(self.diff) == self.nE
        if self.has_obs:
            ang0 = self.diff[0].angles
            for w in self.diff:
                if w.angles.shape != ang0.shape or not np.array_equal(w.angles, ang0):
                    raise ValueError("all workspaces of a sweep must share the angle grid")
            self.n_angles = ang0.shape[0]
            self.P = t(np.stack([w.P_l_costheta for w in self.diff]), torch.float64)
            self.P1 = t(np.stack([w.P_1_l_costheta for w in self.diff]), torch.float64)
            # exp(2 i sigma_l) / (2 i k)  (xs/elastic.py:355)
            self.phase = t(np.stack([np.exp(2j * w.sigma_l[:, 0]) / (2j * w.kinematics.k) for w in self.diff]), torch.complex128)
            self.f_c = t(np.stack([np.asarray(w.f_c, dtype=np.complex128) for w in self.diff]), torch.complex128)
            self.kvec = t(np.array([w.kinematics.k for w in self.diff]), torch.float64)

    def set_radius_factor(self, model):
        """LocalOpticalPotential radius factor depends on the model flag (omp.py:106-113)."""
        import torch

        rf = [model.radius_factor(iw.reaction) for iw in self.integral]
        self.etab[:, _cabi.ET_RSCALE] = torch.tensor(rf, dtype=torch.float64, device=self.device)

    # ------------------------------------------------------------------ nbasis > 64
    MAX_FUSED_NBASIS = 64

    @_on_device
    def _eval_potentials(self, model, params):
        """Built-in forms on the mesh of every energy: (B, E, N) central, spin_orbit (complex) and coulomb (real)
        CUDA tensors (jitr_b200_eval_potentials: the same device code the fused sweep inlines)."""
        import torch

        B = params.shape[0]
        c = torch.empty((B, self.nE, self.N), dtype=torch.complex128, device=self.device)
        so = torch.empty_like(c)
        co = torch.empty((B, self.nE, self.N), dtype=torch.float64, device=self.device)
        rc = _cabi.lib().jitr_b200_eval_potentials(
            self.N, model.model_id, B, self.nE, _cabi.ptr(self.tabs["x"]), _cabi.ptr(self.etab), _cabi.ptr(params),
            params.shape[-1], _cabi.ptr(c), _cabi.ptr(so), _cabi.ptr(co), _cabi.stream_ptr(),
        )
        _cabi.check(rc, "jitr_b200_eval_potentials")
        return c, so, co.to(torch.complex128)

    @_on_device
    def _smatrix_large(self, central, spin_orbit, coulomb, out=None):
        """nbasis beyond the fused kernel (N > 64, e.g. the reference's Frescox regression cases with N = 100): ONE batched
        solve per energy.  All samples, all partial waves and both j of an energy form one batch of
        jitr_b200_solve_batched: the free matrix of l = 0 is shared, the centrifugal term l(l+1)/(a x_n)^2 rides on the
        diagonal with the local potential,  A = F_0 + diag((V_c + V_C + (l.s) V_so)/E0 + l(l+1)/(a x_n)^2),  the asymptotics are
        per system.  The data-dependent break of xs/elastic.py:178-181 is applied afterwards on the device (first l >= 1
        with |1 - S+| < tol and |1 - S-| < tol is the last one kept): no host synchronisation anywhere."""
        import torch

        B, L1 = central.shape[0], self.lmax + 1
        S, nl, status = self._alloc_S(B) if out is None else out
        vc = central if coulomb is None else central + coulomb
        q = self.solver.kernel.quadrature
        dev = self.device
        # system s of a sample: s = 0 -> (l = 0, j = 1/2);  s = 2l - 1 -> (l, l + 1/2);  s = 2l -> (l, l - 1/2)
        ns = 2 * L1 - 1
        ls_of = np.zeros(ns, dtype=np.int64)
        lds_of = np.zeros(ns)
        for l in range(1, L1):
            ls_of[2 * l - 1], lds_of[2 * l - 1] = l, float(l)
            ls_of[2 * l], lds_of[2 * l] = l, -float(l + 1)
        l_idx = torch.tensor(ls_of, device=dev)
        lds = torch.tensor(lds_of, dtype=torch.float64, device=dev)
        x = torch.tensor(q.abscissa, dtype=torch.float64, device=dev)
        for e, iw in enumerate(self.integral):
            kin = iw.kinematics
            key = (e, 0)
            if key not in self._free_cache:
                self._free_cache[key] = torch.tensor(q.kinetic_matrix(iw.a, 0) - np.eye(self.N), dtype=torch.complex128, device=dev)
            cent = (l_idx * (l_idx + 1)).to(torch.float64)[:, None] / (iw.a * x[None, :]) ** 2  # (ns, N)
            loc = vc[:, e, None, :] + (kin.Ecm * cent)[None].to(torch.complex128)
            if spin_orbit is not None:
                loc = loc + lds[None, :, None] * spin_orbit[:, e, None, :]
            asym_sys = self.asym[e][l_idx].reshape(1, ns, 4, 1).expand(B, ns, 4, 1).reshape(B * ns, 4, 1).contiguous()
            r = self.solver.solve_batch(iw.a, kin.Ecm, kin.k, 1, self._free_cache[key], asym_sys, local_potential=loc.reshape(B * ns, self.N))
            Ss = r["S"].reshape(B, ns)
            st = r["status"].reshape(B, ns)
            zero = torch.zeros((B, 1), dtype=Ss.dtype, device=dev)
            sp = torch.cat([Ss[:, :1], Ss[:, 1::2]], dim=1)   # (B, L1)
            sm = torch.cat([zero, Ss[:, 2::2]], dim=1)
            stp = torch.cat([st[:, :1], st[:, 1::2]], dim=1)
            stm = torch.cat([torch.zeros((B, 1), dtype=st.dtype, device=dev), st[:, 2::2]], dim=1)
            if self.tol > 0 and L1 > 1:
                small = ((1 - sp[:, 1:]).abs() < self.tol) & ((1 - sm[:, 1:]).abs() < self.tol)
                none_before = small.to(torch.int32).cumsum(dim=1) == 0          # no hit up to and including l
                nle = torch.clamp(none_before.sum(dim=1) + 2, max=L1).to(torch.int32)  # leading misses + the hit itself + l = 0
            else:
                nle = torch.full((B,), L1, dtype=torch.int32, device=dev)
            keep = torch.arange(L1, device=dev)[None, :] < nle[:, None]
            S[:, e, 0] = torch.where(keep, sp, torch.zeros_like(sp))
            S[:, e, 1] = torch.where(keep, sm, torch.zeros_like(sm))
            status[:, e] = torch.where(keep, torch.maximum(stp, stm), torch.zeros_like(stp)).max(dim=1).values
            nl[:, e] = nle
        return S, nl, status

    # ------------------------------------------------------------------ S-matrix
    def _alloc_S(self, B):
        import torch

        L1 = self.lmax + 1
        S = torch.empty((B, self.nE, 2, L1), dtype=torch.complex128, device=self.device)
        nl = torch.empty((B, self.nE), dtype=torch.int32, device=self.device)
        status = torch.empty((B, self.nE), dtype=torch.int32, device=self.device)
        return S, nl, status

    @_on_device
    def smatrix_params(self, model, params, out=None):
        """Built-in potentials: params (B, n_params) CUDA float64 -> S (B, E, 2, lmax+1), nl (B, E), status (B, E)."""
        import torch

        if getattr(model, "pair_rows", None) is not None:
            # the model maps its parameters to shape numbers per (sample, energy) on the host (DOM): rows (B, E, 18)
            params = model.pair_rows([(iw.reaction, iw.kinematics) for iw in self.integral], params, self.device)
        else:
            params = model.pack(params, self.device)
        if hasattr(model, "radius_factor") and getattr(self, "_rf_model", None) is not model:
            self.set_radius_factor(model)
            self._rf_model = model
        B = params.shape[0]
        if self.N > self.MAX_FUSED_NBASIS:
            c, so, co = self._eval_potentials(model, params)
            return self._smatrix_large(c, so, co, out=out)
        S, nl, status = self._alloc_S(B) if out is None else out
        rc = _cabi.lib().jitr_b200_elastic_sweep_params(
            self.N, self.lmax, model.model_id, B, self.nE, _cabi.ptr(self.tabs["That"]), _cabi.ptr(self.tabs["x"]),
            _cabi.ptr(self.tabs["b"]), _cabi.ptr(self.etab), _cabi.ptr(self.asym), _cabi.ptr(params), params.shape[-1],
            float(self.tol), _cabi.ptr(S), _cabi.ptr(nl), _cabi.ptr(status), _cabi.stream_ptr(),
        )
        _cabi.check(rc, "jitr_b200_elastic_sweep_params")
        return S, nl, status

    @_on_device
    def smatrix_tensors(self, central, spin_orbit=None, coulomb=None, out=None):
        """User potentials: (B, E, N) complex128 CUDA tensors of mesh values (MeV)."""
        import torch

        def prep(v, name):
            if v is None:
                return None
            v = v.to(device=self.device, dtype=torch.complex128).contiguous()
            if v.dim() != 3 or v.shape[1] != self.nE or v.shape[2] != self.N:
                raise ValueError(f"{name} must have shape (B, {self.nE}, {self.N})")
            return v

        central = prep(central, "central_potential")
        spin_orbit = prep(spin_orbit, "spin_orbit_potential")
        coulomb = prep(coulomb, "coulomb_potential")
        B = central.shape[0]
        if self.N > self.MAX_FUSED_NBASIS:
            return self._smatrix_large(central, spin_orbit, coulomb, out=out)
        S, nl, status = self._alloc_S(B) if out is None else out
        rc = _cabi.lib().jitr_b200_elastic_sweep_tensors(
            self.N, self.lmax, B, self.nE, _cabi.ptr(self.tabs["That"]), _cabi.ptr(self.tabs["x"]), _cabi.ptr(self.tabs["b"]),
            _cabi.ptr(self.etab), _cabi.ptr(self.asym), _cabi.ptr(central), _cabi.ptr(spin_orbit), _cabi.ptr(coulomb),
            float(self.tol), _cabi.ptr(S), _cabi.ptr(nl), _cabi.ptr(status), _cabi.stream_ptr(),
        )
        _cabi.check(rc, "jitr_b200_elastic_sweep_tensors")
        return S, nl, status

    # ------------------------------------------------------------------ observables
    @_on_device
    def observables(self, S, nl, want_Ay=True, want_Q=True, out=None):
        """S (B, E, 2, lmax+1), nl (B, E) -> ElasticXS of CUDA tensors (B, E, n_angles) / (B, E)."""
        import torch

        if not self.has_obs:
            raise ValueError("observables need DifferentialWorkspace objects for every energy")
        B = S.shape[0]
        if out is None:
            dsdo = torch.empty((B, self.nE, self.n_angles), dtype=torch.float64, device=self.device)
            Ay = torch.empty_like(dsdo) if want_Ay else None
            Q = torch.empty_like(dsdo) if want_Q else None
            tr = torch.empty((B, self.nE, 2), dtype=torch.float64, device=self.device)
        else:
            dsdo, Ay, Q, tr = out
        rc = _cabi.lib().jitr_b200_elastic_observables(
            self.lmax, self.n_angles, B, self.nE, _cabi.ptr(S), _cabi.ptr(nl), _cabi.ptr(self.P), _cabi.ptr(self.P1),
            _cabi.ptr(self.phase), _cabi.ptr(self.f_c), _cabi.ptr(self.kvec), _cabi.ptr(dsdo), _cabi.ptr(Ay), _cabi.ptr(Q),
            _cabi.ptr(tr), _cabi.stream_ptr(),
        )
        _cabi.check(rc, "jitr_b200_elastic_observables")
        return ElasticXS(dsdo, Ay, Q, tr[..., 0], tr[..., 1])

    @_on_device
    def misfit(self, S, nl, y, weights, scale=None, want_dsdo=False):
        """Weighted misfit against data, reduced on the device: sum_theta w (scale * dsdo - y)^2 per (sample, energy).

        ``y``, ``weights``, ``scale``: (E, n_angles) arrays (``scale`` e.g. 1 / Rutherford for ratio data, None = 1).
        Returns ``(misfit (B, E), sigma_tot (B, E), sigma_rxn (B, E)[, dsdo (B, E, n_angles)])`` CUDA tensors; without
        ``want_dsdo`` the angular distributions never leave the kernel."""
        import torch

        if not self.has_obs:
            raise ValueError("observables need DifferentialWorkspace objects for every energy")
        B = S.shape[0]

        def tab(v, name):
            if v is None:
                return None
            t = torch.as_tensor(np.asarray(v.cpu() if isinstance(v, torch.Tensor) else v, dtype=np.float64), device=self.device)
            if tuple(t.shape) != (self.nE, self.n_angles):
                raise ValueError(f"{name} must have shape ({self.nE}, {self.n_angles})")
            return t.contiguous()

        yt, wt, st = tab(y, "y"), tab(weights, "weights"), tab(scale, "scale")
        mis = torch.empty((B, self.nE), dtype=torch.float64, device=self.device)
        tr = torch.empty((B, self.nE, 2), dtype=torch.float64, device=self.device)
        dsdo = torch.empty((B, self.nE, self.n_angles), dtype=torch.float64, device=self.device) if want_dsdo else None
        rc = _cabi.lib().jitr_b200_elastic_misfit(
            self.lmax, self.n_angles, B, self.nE, _cabi.ptr(S), _cabi.ptr(nl), _cabi.ptr(self.P), _cabi.ptr(self.P1),
            _cabi.ptr(self.phase), _cabi.ptr(self.f_c), _cabi.ptr(self.kvec), _cabi.ptr(yt), _cabi.ptr(wt), _cabi.ptr(st),
            _cabi.ptr(mis), _cabi.ptr(tr), _cabi.ptr(dsdo), _cabi.stream_ptr(),
        )
        _cabi.check(rc, "jitr_b200_elastic_misfit")
        out = (mis, tr[..., 0], tr[..., 1])
        return out + (dsdo,) if want_dsdo else out

    def misfit_params(self, model, params, y, weights, scale=None):
        """Built-in potentials -> misfit (B, E): the batched likelihood kernel of a calibration loop."""
        S, nl, status = self.smatrix_params(model, params)
        return self.misfit(S, nl, y, weights, scale) + (status,)

    @_on_device
    def xs_params_to_host(self, model, params_host, dsdo_host, tr_host, chunk=8192, status_host=None):
        """End-to-end sweep with HOST buffers: ``params_host`` (B, n_params) pinned float64 ->
        ``dsdo_host`` (B, E, n_angles) and ``tr_host`` (B, E, 2) pinned float64, and -- when given --
        ``status_host`` (B, E) pinned int32 (JITR_B200_STATUS_* per (sample, energy): a non-finite or singular
        system is reported, not silent).  The sample axis is processed in chunks on two device buffer sets so that
        the device->host copy of chunk i overlaps the solve of chunk i+1 (copy stream + events).  Returns the
        number of systems solved."""
        import torch

        B = params_host.shape[0]
        cur = torch.cuda.current_stream(self.device)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.device)
            self._bufs = {}
        key = (chunk, params_host.shape[1])
        if key not in self._bufs:
            mk = lambda *sh, dt=torch.float64: torch.empty(sh, dtype=dt, device=self.device)
            self._bufs[key] = [
                dict(p=mk(chunk, params_host.shape[1]), S=self._alloc_S(chunk),
                     obs=(mk(chunk, self.nE, self.n_angles), None, None, mk(chunk, self.nE, 2)), done=None)
                for _ in range(2)
            ]
        bufs = self._bufs[key]
        nsolve = torch.zeros((), dtype=torch.int64, device=self.device)
        for i, s0 in enumerate(range(0, B, chunk)):
            n = min(chunk, B - s0)
            b = bufs[i % 2]
            if b["done"] is not None:
                cur.wait_event(b["done"])  # previous D2H out of this buffer set has finished
            S, nl, status = (t[:n] for t in b["S"])
            if getattr(model, "pair_rows", None) is not None:
                # parameter map on the host (one row per sample and energy): hand it the host rows
                self.smatrix_params(model, params_host[s0 : s0 + n], out=(S, nl, status))
            else:
                b["p"][:n].copy_(params_host[s0 : s0 + n], non_blocking=True)
                self.smatrix_params(model, b["p"][:n], out=(S, nl, status))
            dsdo, _, _, tr = b["obs"]
            self.observables(S, nl, want_Ay=False, want_Q=False, out=(dsdo[:n], None, None, tr[:n]))
            nsolve += (2 * nl.to(torch.int64) - 1).sum()
            ev = torch.cuda.Event()
            ev.record(cur)
            self._copy_stream.wait_event(ev)
            with torch.cuda.stream(self._copy_stream):
                dsdo_host[s0 : s0 + n].copy_(dsdo[:n], non_blocking=True)
                tr_host[s0 : s0 + n].copy_(tr[:n], non_blocking=True)
                if status_host is not None:
                    status_host[s0 : s0 + n].copy_(status, non_blocking=True)
                b["done"] = torch.cuda.Event()
                b["done"].record(self._copy_stream)
        cur.wait_stream(self._copy_stream)
        return nsolve

    def xs_params(self, model, params, return_status=False):
        S, nl, status = self.smatrix_params(model, params)
        x = self.observables(S, nl)
        return (x, status) if return_status else x

    def xs_tensors(self, central, spin_orbit=None, coulomb=None, return_status=False):
        S, nl, status = self.smatrix_tensors(central, spin_orbit, coulomb)
        x = self.observables(S, nl)
        return (x, status) if return_status else x
