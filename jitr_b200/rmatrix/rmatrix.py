"""``Solver``: the reference's R-matrix solver API (rmatrix/rmatrix.py:19-246) on the B200 path.

Setup helpers (``radial_grid``, ``free_matrix``, ``interaction_matrix`` ...) are float64 host code
returning numpy arrays with the reference's values.  ``solve`` keeps the single-system signature
and return convention but runs on the GPU through the C ABI (``jitr_b200_solve_batched``);
``solve_batch`` is the same operator with a leading batch axis on the potentials (torch CUDA
tensors, complex128, C-contiguous) -- the form the UQ sweeps use.  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from .. import _cabi
from ..quadrature import Kernel
from ..utils.misc import block


class Solver:
    """Solve coupled-channel Schroedinger equations with the calculable R-matrix method."""

    def __init__(self, nbasis, basis="Legendre", **args):
        self.kernel = Kernel(nbasis, basis)
        self._dev = {}

    # ------------------------------------------------------------------ host setup (reference values)
    def radial_grid(self, a, k0):
        """rmatrix.py:26-28."""
        return self.kernel.radial_grid(a / k0)

    def nonlocal_radial_grids(self, a, k0):
        """rmatrix.py:30-34."""
        return self.kernel.nonlocal_radial_grids(a / k0)

    def precompute_boundaries(self, a=None):
        """Basis functions at the channel radius, b_n = f_n(a) (rmatrix.py:36-42); a-independent."""
        q = self.kernel.quadrature
        N = q.nbasis
        n = np.arange(1, N + 1)
        x = q.abscissa
        # legendre(n, a, s=a): x = 1, P_N(1) = 1
        return ((-1.0) ** (N - n) * np.sqrt((1 - x) / x) / (1.0 - x)).astype(np.complex128)

    def get_channel_block(self, matrix, i, j=None):
        """rmatrix.py:44-54."""
        nb = self.kernel.quadrature.nbasis
        if j is None:
            j = i
        return np.asarray(block(np.asarray(matrix), (i, j), (nb, nb)))

    def kinetic_matrix(self, a, l, mu=None):
        """rmatrix.py:56-81."""
        l_array = np.atleast_1d(np.asarray(l))
        mu_array = np.ones(l_array.shape, dtype=np.float64) if mu is None else np.asarray(mu, dtype=np.float64)
        nb = self.kernel.quadrature.nbasis
        nch = int(np.size(l_array))
        K = np.zeros((nb * nch, nb * nch), dtype=np.complex128)
        for i in range(nch):
            sl = slice(i * nb, (i + 1) * nb)
            K[sl, sl] += self.kernel.quadrature.kinetic_matrix(a, int(l_array[i])) * mu_array[0] / mu_array[i]
        return K

    def energy_matrix(self, a, l, E=None):
        """rmatrix.py:83-104."""
        l_array = np.atleast_1d(np.asarray(l))
        scale = np.ones(l_array.shape, dtype=np.float64) if E is None else np.asarray(E, dtype=np.float64)
        nb = self.kernel.quadrature.nbasis
        nch = int(np.size(l_array))
        Em = np.zeros((nb * nch, nb * nch), dtype=np.complex128)
        for i in range(nch):
            sl = slice(i * nb, (i + 1) * nb)
            Em[sl, sl] += self.kernel.overlap * scale[i]
        return Em / scale[0]

    def free_matrix(self, a, l, E=None, mu=None, coupled=True):
        """rmatrix.py:106-124."""
        l_array = np.atleast_1d(np.asarray(l))
        F = self.kinetic_matrix(a, l_array, mu) - self.energy_matrix(a, l_array, E)
        if coupled:
            return F
        return [self.get_channel_block(F, i) for i in range(int(np.size(l_array)))]

    def interaction_matrix(self, k0, E0, a, nch, local_potential=None, nonlocal_potential=None):
        """rmatrix.py:126-178 (host numpy; the device path assembles this inside the kernel)."""
        nb = self.kernel.quadrature.nbasis
        size = nb * nch
        V = np.zeros((size, size), dtype=np.complex128)
        radius = a / k0
        if local_potential is not None:
            loc = self.kernel.matrix_local(local_potential)
            if loc.ndim == 1:
                if nch != 1:
                    raise ValueError("1D local potentials are only valid for one channel")
                loc = loc.reshape(1, 1, nb)
            elif loc.shape[:2] != (nch, nch):
                raise ValueError(
                    f"local_potential leading dimensions {loc.shape[:2]} do not match nch={nch}; "
                    f"expected shape ({nch}, {nch}, {nb})"
                )
            for i in range(nch):
                for j in range(nch):
                    V[i * nb : (i + 1) * nb, j * nb : (j + 1) * nb] = np.diag(loc[i, j, ...])
        if nonlocal_potential is not None:
            nl = self.kernel.matrix_nonlocal(nonlocal_potential, radius)
            if nl.ndim == 2:
                if nch != 1:
                    raise ValueError("2D nonlocal potentials are only valid for one channel")
                nl = nl.reshape(1, 1, nb, nb)
            elif nl.shape[:2] != (nch, nch):
                raise ValueError(
                    f"nonlocal_potential leading dimensions {nl.shape[:2]} do not match nch={nch}; "
                    f"expected shape ({nch}, {nch}, {nb}, {nb})"
                )
            V += nl.swapaxes(1, 2).reshape(size, size, order="C") / k0
        return V / E0

    # ------------------------------------------------------------------ device tables
    def device_tables(self, device=None):
        """Replicated KB-sized tables on the device: That (N,N) f64, mesh x (N,), b (N,), sqrt(w) (N,)."""
        import torch

        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        key = str(device)
        if key not in self._dev:
            q = self.kernel.quadrature
            self._dev[key] = dict(
                That=torch.tensor(np.ascontiguousarray(q.kinetic_hat()), dtype=torch.float64, device=device),
                x=torch.tensor(q.abscissa, dtype=torch.float64, device=device),
                b=torch.tensor(np.real(self.precompute_boundaries()), dtype=torch.float64, device=device),
                wsqrt=torch.tensor(np.sqrt(q.weights), dtype=torch.float64, device=device),
            )
        return self._dev[key]

    # ------------------------------------------------------------------ the solve
    def solve(
        self, channels, asymptotics, local_potential=None, nonlocal_potential=None, interaction_matrix=None,
        free_matrix=None, basis_boundary=None, weights=None, wavefunction=False,
    ):
        """One coupled set of channels; same contract as rmatrix.py:180-246.

        Returns ``(R, S, uext_prime_boundary)`` or ``(R, S, coeffs, uext_prime_boundary)`` as numpy arrays."""
        import torch

        nb = self.kernel.quadrature.nbasis
        nch = channels.size
        size = nch * nb
        if free_matrix is None:
            free_matrix = self.free_matrix(channels.a, channels.l, channels.E, channels.mu, coupled=True)
        free_matrix = np.asarray(free_matrix, dtype=np.complex128)
        if basis_boundary is None:
            basis_boundary = self.precompute_boundaries(channels.a)
        basis_boundary = np.asarray(basis_boundary)
        assert free_matrix.shape == (size, size)
        assert basis_boundary.shape == (nb,)
        if interaction_matrix is not None:
            interaction_matrix = np.asarray(interaction_matrix, dtype=np.complex128)
            assert interaction_matrix.shape == (size, size)
            local_t = nonlocal_t = None
            inter_t = torch.from_numpy(np.ascontiguousarray(interaction_matrix)).cuda().unsqueeze(0)
        else:
            inter_t = None
            local_t, nonlocal_t = self._validated_potentials(nch, local_potential, nonlocal_potential, batch=False)
        asym = np.stack([asymptotics.Hp, asymptotics.Hm, asymptotics.Hpp, asymptotics.Hmp]).astype(np.complex128)
        out = self.solve_batch(
            a=float(channels.a), E0=float(channels.E[0]), k0=float(channels.k[0]), nch=nch,
            free_matrix=torch.from_numpy(np.ascontiguousarray(free_matrix)).cuda(),
            asym=torch.from_numpy(np.ascontiguousarray(asym)).cuda(),
            local_potential=local_t, nonlocal_potential=nonlocal_t, interaction_matrix=inter_t,
            basis_boundary=torch.from_numpy(np.ascontiguousarray(np.real(basis_boundary).astype(np.float64))).cuda(),
            weights=weights, wavefunction=wavefunction, batch=1,
        )
        R, S, uext = (out[k][0].cpu().numpy() for k in ("R", "S", "uext"))
        st = int(out["status"][0].item())
        if st != _cabi.STATUS_OK:
            raise np.linalg.LinAlgError("Matrix is singular or non-finite (status %d)" % st)
        if not wavefunction:
            return R, S, uext
        return R, S, out["coeffs"][0].cpu().numpy(), uext

    def _validated_potentials(self, nch, local_potential, nonlocal_potential, batch):
        """Shape checks of rmatrix.py:141-176 / kernel.py:180-214; returns CUDA complex128 tensors with a batch axis."""
        import torch

        nb = self.kernel.quadrature.nbasis

        def as_tensor(v):
            if isinstance(v, torch.Tensor):
                return v.to(device="cuda", dtype=torch.complex128)
            return torch.from_numpy(np.ascontiguousarray(np.asarray(v, dtype=np.complex128))).cuda()

        loc = nl = None
        if local_potential is not None:
            loc = as_tensor(local_potential)
            core = loc.shape[1:] if batch else loc.shape
            if len(core) == 1 and core == (nb,):
                if nch != 1:
                    raise ValueError("1D local potentials are only valid for one channel")
            elif len(core) == 3 and core[0] == core[1] and core[2] == nb:
                if tuple(core[:2]) != (nch, nch):
                    raise ValueError(
                        f"local_potential leading dimensions {tuple(core[:2])} do not match nch={nch}; "
                        f"expected shape ({nch}, {nch}, {nb})"
                    )
            else:
                raise ValueError(f"local potential values must have shape ({nb},) or (nchannels, nchannels, {nb})")
            loc = loc.reshape(-1, nch, nch, nb).contiguous()
        if nonlocal_potential is not None:
            nl = as_tensor(nonlocal_potential)
            core = nl.shape[1:] if batch else nl.shape
            if len(core) == 2 and tuple(core) == (nb, nb):
                if nch != 1:
                    raise ValueError("2D nonlocal potentials are only valid for one channel")
            elif len(core) == 4 and core[0] == core[1] and tuple(core[2:]) == (nb, nb):
                if tuple(core[:2]) != (nch, nch):
                    raise ValueError(
                        f"nonlocal_potential leading dimensions {tuple(core[:2])} do not match nch={nch}; "
                        f"expected shape ({nch}, {nch}, {nb}, {nb})"
                    )
            else:
                raise ValueError(
                    f"nonlocal potential values must have shape ({nb}, {nb}) or (nchannels, nchannels, {nb}, {nb})"
                )
            nl = nl.reshape(-1, nch, nch, nb, nb).contiguous()
        return loc, nl

    def solve_batch(
        self, a, E0, k0, nch, free_matrix, asym, local_potential=None, nonlocal_potential=None,
        interaction_matrix=None, basis_boundary=None, weights=None, wavefunction=False, batch=None, free_index=None,
    ):
        """Batched ``solve`` on the device.

        ``free_index`` (B,) int32 CUDA tensor: the batch mixes partial waves -- ``free_matrix`` is then (n_free, n, n) and system b
        uses ``free_matrix[free_index[b]]`` (one call for all (sample, l) systems of a sweep; give ``asym`` per system).

        ``free_matrix``: (n, n) shared or (B, n, n) complex128 CUDA tensor; ``asym``: (4, nch) shared or
        (B, 4, nch) rows H+, H-, H+', H-'; potentials with a leading batch axis:
        local (B, nch, nch, N) [or (B, N) for nch = 1], nonlocal (B, nch, nch, N, N) [or (B, N, N)],
        or a precomputed ``interaction_matrix`` (B, n, n).  ``a``, ``E0``, ``k0``: scalars shared by the
        batch, or (B,) tensors/arrays.  Returns a dict of CUDA tensors R, S (B, nch, nch), uext (B, nch),
        status (B,) int32 and, if ``wavefunction``, coeffs (B, nch, N).
        """
        import torch

        _cabi.require_device()
        L = _cabi.lib()
        nb = self.kernel.quadrature.nbasis
        n = nb * nch
        # the batch lives where its inputs live: the free matrix decides the device (the current device for host inputs)
        dev = free_matrix.device if isinstance(free_matrix, torch.Tensor) and free_matrix.is_cuda else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            return self._solve_batch_on(dev, a, E0, k0, nch, free_matrix, asym, local_potential, nonlocal_potential, interaction_matrix,
                                        basis_boundary, weights, wavefunction, batch, free_index)

    def _solve_batch_on(self, dev, a, E0, k0, nch, free_matrix, asym, local_potential, nonlocal_potential, interaction_matrix,
                        basis_boundary, weights, wavefunction, batch, free_index=None):
        import torch

        L = _cabi.lib()
        nb = self.kernel.quadrature.nbasis
        n = nb * nch
        tabs = self.device_tables(dev)
        loc = nl = None
        if interaction_matrix is None:
            loc, nl = self._validated_potentials(nch, local_potential, nonlocal_potential, batch=True) if batch is None else (
                local_potential, nonlocal_potential)
            if batch is None:
                batch = (loc if loc is not None else nl).shape[0] if (loc is not None or nl is not None) else 1
            else:
                loc = None if loc is None else loc.reshape(-1, nch, nch, nb).contiguous()
                nl = None if nl is None else nl.reshape(-1, nch, nch, nb, nb).contiguous()
        else:
            interaction_matrix = interaction_matrix.to(device=dev, dtype=torch.complex128).reshape(-1, n, n).contiguous()
            if batch is None:
                batch = interaction_matrix.shape[0]
        B = int(batch)
        free_matrix = free_matrix.to(device=dev, dtype=torch.complex128).contiguous()
        free_stride = 0 if (free_matrix.dim() == 2 or free_index is not None) else n * n
        assert free_matrix.shape[-2:] == (n, n)
        if free_index is not None:
            free_index = free_index.to(device=dev, dtype=torch.int32).contiguous()
            assert free_matrix.dim() == 3 and free_index.shape == (B,)
        asym = asym.to(device=dev, dtype=torch.complex128).contiguous()
        asym_stride = 0 if asym.dim() == 2 else 4 * nch
        assert asym.shape[-2:] == (4, nch)

        def per_sys(v):
            t = torch.as_tensor(v, dtype=torch.float64, device=dev)
            return t.expand(B) if t.dim() == 0 else t.reshape(B)

        scalar_sys = all(np.ndim(v) == 0 and not isinstance(v, torch.Tensor) for v in (a, E0, k0))
        if scalar_sys:
            sys_t = torch.tensor([[float(a), float(E0), float(k0), 0.0]], dtype=torch.float64, device=dev)
            sys_stride = 0
        else:
            sys_t = torch.stack([per_sys(a), per_sys(E0), per_sys(k0), torch.zeros(B, dtype=torch.float64, device=dev)], dim=1).contiguous()
            sys_stride = 4
        bvec = tabs["b"] if basis_boundary is None else basis_boundary.to(device=dev, dtype=torch.float64).contiguous()
        if weights is None:
            w = np.zeros(nch, dtype=np.float64)
            w[0] = 1
        else:
            w = np.asarray(weights, dtype=np.float64)
        w_t = torch.tensor(w, dtype=torch.float64, device=dev)
        R = torch.empty((B, nch, nch), dtype=torch.complex128, device=dev)
        S = torch.empty((B, nch, nch), dtype=torch.complex128, device=dev)
        uext = torch.empty((B, nch), dtype=torch.complex128, device=dev)
        coeffs = torch.empty((B, nch, nb), dtype=torch.complex128, device=dev) if wavefunction else None
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        # device scratch of the CTA-per-system kernels (one slab per resident CTA; none for the warp kernel)
        ws_bytes = L.jitr_b200_solve_workspace_bytes(nb, nch, B, int(bool(wavefunction)))
        if ws_bytes < 0:
            raise ValueError("jitr_b200_solve_workspace_bytes: bad arguments")
        ws = torch.empty((ws_bytes // 16,), dtype=torch.complex128, device=dev) if ws_bytes else None
        if free_index is not None:
            rc = L.jitr_b200_solve_batched_indexed(
                nb, nch, B, _cabi.ptr(free_matrix), int(free_matrix.shape[0]), _cabi.ptr(free_index), _cabi.ptr(loc), _cabi.ptr(nl),
                _cabi.ptr(interaction_matrix), _cabi.ptr(tabs["wsqrt"]), _cabi.ptr(sys_t), sys_stride, _cabi.ptr(bvec), _cabi.ptr(asym),
                asym_stride, _cabi.ptr(w_t), _cabi.ptr(R), _cabi.ptr(S), _cabi.ptr(uext), _cabi.ptr(coeffs), _cabi.ptr(ws),
                _cabi.ptr(status), _cabi.stream_ptr(),
            )
        else:
            rc = L.jitr_b200_solve_batched(
                nb, nch, B, _cabi.ptr(free_matrix), free_stride, _cabi.ptr(loc), _cabi.ptr(nl), _cabi.ptr(interaction_matrix),
                _cabi.ptr(tabs["wsqrt"]), _cabi.ptr(sys_t), sys_stride, _cabi.ptr(bvec), _cabi.ptr(asym), asym_stride,
                _cabi.ptr(w_t), _cabi.ptr(R), _cabi.ptr(S), _cabi.ptr(uext), _cabi.ptr(coeffs), _cabi.ptr(ws),
                _cabi.ptr(status), _cabi.stream_ptr(),
            )
        _cabi.check(rc, "jitr_b200_solve_batched")
        out = dict(R=R, S=S, uext=uext, status=status)
        if wavefunction:
            out["coeffs"] = coeffs
        return out
