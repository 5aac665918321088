"""Numpy model of the warp-level blocked bordered LU used by the CUDA solve kernel.

Not product code: a design aid that mirrors, step by step, what csrc/lu_warp.cuh does
(panel-in-registers with implicit pivot search + LAPACK-style row positions, U-phase,
T-phase) so that the algebra (bordered matrix M = [[A, b],[b^T, 0]], corner = -b^T A^-1 b)
can be checked on the CPU before the CUDA translation is run on a GPU.
"""
import numpy as np


def bordered_lu_R(A, b, NB=8):
    N = A.shape[0]
    M = np.zeros((N + 1, N + 1), dtype=np.complex128)
    M[:N, :N] = A
    M[:N, N] = b
    M[N, :N] = b
    nswaps = 0
    for j0 in range(0, N, NB):
        nb = min(NB, N - j0)
        rows = list(range(j0, N + 1))  # physical rows held by lane-slots (row N = z row)
        # ---- phase P: panel in "registers": pan[r] = M[r, j0:j0+nb] + y
        pan = {r: np.concatenate([M[r, j0:j0 + nb], [M[r, N]]]) for r in rows}
        pos = {r: r for r in rows}          # current logical position of the row held by each lane-slot
        alive = {r: True for r in rows}
        for kk in range(nb):
            cand = [(abs(pan[r][kk]) ** 2, r) for r in rows if alive[r] and r != N]
            _, P = max(cand, key=lambda t: (t[0], -t[1]))
            pp = pos[P]
            # LAPACK-style position swap: whoever sits at j0+kk goes to pp
            Q = [r for r in rows if pos[r] == j0 + kk][0]
            pos[Q] = pp
            pos[P] = j0 + kk
            if pp != j0 + kk:
                nswaps += 1
                # physical swap of the trailing columns (j0+nb .. N-1) in "smem"
                c = slice(j0 + nb, N)
                t = M[j0 + kk, c].copy()
                M[j0 + kk, c] = M[pp, c]
                M[pp, c] = t
            u = pan[P].copy()
            rinv = 1.0 / u[kk]
            alive[P] = False
            for r in rows:
                if alive[r]:
                    l = pan[r][kk] * rinv
                    pan[r][kk] = l
                    pan[r][kk + 1:] -= l * u[kk + 1:]
        # write back panel columns + y at final positions
        for r in rows:
            M[pos[r], j0:j0 + nb] = pan[r][:nb]
            M[pos[r], N] = pan[r][nb]
        mt = N - j0 - nb
        if mt == 0:
            break
        # ---- phase U: lane c owns column j0+nb+c
        L11 = M[j0:j0 + nb, j0:j0 + nb]
        lz = M[N, j0:j0 + nb]
        for c in range(mt):
            col = j0 + nb + c
            a = M[j0:j0 + nb, col].copy()
            for kk in range(nb):
                for k2 in range(kk):
                    a[kk] -= L11[kk, k2] * a[k2]
            M[j0:j0 + nb, col] = a
            M[N, col] -= np.dot(lz, a)
        # ---- phase T
        r0 = j0 + nb
        M[r0:N, r0:N] -= M[r0:N, j0:j0 + nb] @ M[j0:j0 + nb, r0:N]
    return -M[N, N], nswaps


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N in (40, 60, 17):
        worst = 0
        for t in range(20):
            A = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
            A = A + A.T
            b = rng.standard_normal(N)
            ref = b @ np.linalg.solve(A, b)
            got, ns = bordered_lu_R(A, b)
            worst = max(worst, abs(got - ref) / abs(ref))
        print(N, "worst rel err", worst, "swaps", ns)
