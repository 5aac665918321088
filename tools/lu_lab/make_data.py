#!/usr/bin/env python
"""Write lab_data_N<N>.bin for lu_lab: realistic bordered systems of the elastic sweep.

File layout (little endian): int32 N, int32 K; That[N*N] f64; b[N] f64; diag[K][N] c128;
corner_ref[K] c128  (corner = -b^T (That + diag)^-1 b, numpy LAPACK solve).
The K systems are the (l, j) partial waves of n+208Pb at 30 MeV with the KD03 potential.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from jitr_b200.quadrature import Kernel  # noqa: E402
from oracle import jitr_oracle as orc  # noqa: E402


def main(N, out):
    ker = Kernel(N)
    q = ker.quadrature
    That = q.kinetic_hat()
    x = q.abscissa
    b = (-1.0) ** (N - np.arange(1, N + 1)) / np.sqrt(x * (1 - x))
    g = np.load(os.path.join(ROOT, "tests", "golden", "config1_n208Pb_30MeV.npz"))
    kin = g["kin"]  # (Elab, Ecm, mu, k, eta)
    Ecm, k = float(kin[1]), float(kin[3])
    Rch = 12.0
    a = k * Rch
    r = x * Rch
    p = np.load(os.path.join(ROOT, "tests", "golden", "kduq_params.npz"))["kd03_n"]
    c, so, co = orc.kduq_potentials(r, (1, 0), (208, 82), 30.0, p)
    diags = []
    for l in range(0, 31):
        for ls in ([l] if l == 0 else [l, -(l + 1)]):
            d = np.diag(That) + l * (l + 1) / x**2 + a * a * ((c + co) / Ecm - 1.0) + ls * a * a * so / Ecm
            diags.append(d.astype(np.complex128))
    diags = np.array(diags)
    ref = []
    for d in diags:
        A = That.astype(np.complex128).copy()
        A[np.diag_indices(N)] = d
        ref.append(-b @ np.linalg.solve(A, b.astype(np.complex128)))
    ref = np.array(ref)
    with open(out, "wb") as f:
        f.write(np.array([N, len(diags)], dtype=np.int32).tobytes())
        f.write(That.astype(np.float64).tobytes())
        f.write(b.astype(np.float64).tobytes())
        f.write(diags.tobytes())
        f.write(ref.tobytes())
    print(f"wrote {out}: N={N}, K={len(diags)}")


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    main(N, os.path.join(os.path.dirname(os.path.abspath(__file__)), f"lab_data_N{N}.bin"))
