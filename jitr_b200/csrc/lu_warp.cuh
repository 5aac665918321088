// Warp-per-system blocked LU (complex128, partial pivoting) of the BORDERED matrix
//
//        M = [ A   b ]      (N+1) x (N+1), stored row-major in shared memory with LD = N+1
//            [ b^T 0 ]
//
// eliminated on its first N columns with the pivot search restricted to the first N rows; the
// corner that is left is the Schur complement  -b^T A^-1 b , i.e. minus the (scaled) R-matrix
// (rmatrix/core.py:7-22 computes the same number through the explicit inverse).  Nothing of L or
// U is needed afterwards, so no triangular back-substitution is run.
//
// One warp owns one system.  Right-looking, panel width NB = 8, three phases per panel:
//   P  panel factorisation with the panel held in REGISTERS, one matrix row per lane (two row
//      slots while more than 32 rows are alive): pivot search = REDUX.MAX on the high word of
//      |a|^2 + ballot, pivot-row broadcast = shuffles, rank-1 updates = DFMA on registers.
//      Column N (the right-hand side y) rides along as a ninth panel column, row N (z = U^-T b)
//      as one more row that is never a pivot candidate.  Row interchanges of the trailing columns
//      are done in shared memory as soon as the pivot is known (lane c always owns column c, so no
//      barrier is needed and the LDS/STS latency hides under the register updates).
//   U  U12 = L11^-1 A12: lane c owns trailing column c and runs the 8-step forward substitution
//      in registers (L11 read as shared-memory broadcasts); the z row is updated here too.
//   T  A22 -= L21 U12 with 4x4 complex register tiles per lane (8 row groups x 4 column groups,
//      cyclic, conflict-free for the odd leading dimension LD = N+1).
#pragma once
#include <cuda_runtime.h>

namespace jitr {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
// acc -= a * b
__device__ __forceinline__ void cfnma(double2& acc, double2 a, double2 b) {
    acc.x = fma(-a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ double2 crecip(double2 a) {
    double d = 1.0 / fma(a.x, a.x, a.y * a.y);
    return make_double2(a.x * d, -a.y * d);
}
__device__ __forceinline__ double2 cdiv(double2 n, double2 d) { return cmul(n, crecip(d)); }
__device__ __forceinline__ double2 shfl2(double2 v, int src) {
    return make_double2(__shfl_sync(kFull, v.x, src), __shfl_sync(kFull, v.y, src));
}

// ---- phase T ------------------------------------------------------------------------------------
template <int N, int J0, int NBc, int RB, int CB>
__device__ __forceinline__ void trailing_update(double2* __restrict__ M, int lane) {
    constexpr int LD = N + 1;
    constexpr int MT = N - J0 - NBc;
    constexpr int R0 = J0 + NBc;
    constexpr int TI = (MT + 7) / 8;
    constexpr int TJ = (MT + 3) / 4;
    const int rg = lane & 7, cg = lane >> 3;
#pragma unroll
    for (int tj0 = 0; tj0 < TJ; tj0 += CB) {
#pragma unroll
        for (int ti0 = 0; ti0 < TI; ti0 += RB) {
            int rowoff[RB], coloff[CB];
            bool rok[RB], cok[CB];
#pragma unroll
            for (int ti = 0; ti < RB; ++ti) {
                const int r = rg + 8 * (ti0 + ti);
                rok[ti] = (8 * (ti0 + ti) + 7 < MT) || (r < MT);
                rowoff[ti] = (R0 + (rok[ti] ? r : 0)) * LD;
            }
#pragma unroll
            for (int tj = 0; tj < CB; ++tj) {
                const int c = cg + 4 * (tj0 + tj);
                cok[tj] = (4 * (tj0 + tj) + 3 < MT) || (c < MT);
                coloff[tj] = R0 + (cok[tj] ? c : 0);
            }
            double2 acc[RB][CB];
#pragma unroll
            for (int ti = 0; ti < RB; ++ti)
#pragma unroll
                for (int tj = 0; tj < CB; ++tj)
                    if (ti0 + ti < TI && tj0 + tj < TJ) acc[ti][tj] = M[rowoff[ti] + coloff[tj]];
#pragma unroll
            for (int k = 0; k < NBc; ++k) {
                double2 l[RB], u[CB];
#pragma unroll
                for (int ti = 0; ti < RB; ++ti)
                    if (ti0 + ti < TI) l[ti] = M[rowoff[ti] + J0 + k];
#pragma unroll
                for (int tj = 0; tj < CB; ++tj)
                    if (tj0 + tj < TJ) u[tj] = M[(J0 + k) * LD + coloff[tj]];
#pragma unroll
                for (int ti = 0; ti < RB; ++ti)
#pragma unroll
                    for (int tj = 0; tj < CB; ++tj)
                        if (ti0 + ti < TI && tj0 + tj < TJ) cfnma(acc[ti][tj], l[ti], u[tj]);
            }
#pragma unroll
            for (int ti = 0; ti < RB; ++ti)
#pragma unroll
                for (int tj = 0; tj < CB; ++tj)
                    if (ti0 + ti < TI && tj0 + tj < TJ && rok[ti] && cok[tj]) M[rowoff[ti] + coloff[tj]] = acc[ti][tj];
        }
    }
}

// ---- one panel stage (phases P, U, T), recursing on J0 -------------------------------------------
template <int N, int J0>
struct Stage {
    static constexpr int LD = N + 1;
    static constexpr int NBc = (N - J0 < 8) ? (N - J0) : 8;
    static constexpr int ROWS = N - J0 + 1;  // rows J0..N (row N is the z row)
    static constexpr int SLOTS = (ROWS + 31) / 32;
    static constexpr int MT = N - J0 - NBc;
    static constexpr int R0 = J0 + NBc;
    static_assert(SLOTS <= 2, "warp LU supports N <= 63");

    __device__ __forceinline__ static void run(double2* __restrict__ M, int lane, int& singular, double2& corner) {
        // ------------------------------------------------------------------ phase P
        double2 p[SLOTS][NBc + 1];
        int pos[SLOTS];
        bool alive[SLOTS], valid[SLOTS], cand[SLOTS];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int row = J0 + 32 * s + lane;
            valid[s] = (32 * s + 31 < ROWS) || (row <= N);
            alive[s] = valid[s];
            cand[s] = valid[s] && (row != N);
            pos[s] = row;
            const int base = (valid[s] ? row : J0) * LD;
#pragma unroll
            for (int j = 0; j < NBc; ++j) p[s][j] = M[base + J0 + j];
            p[s][NBc] = M[base + N];
        }
#pragma unroll
        for (int kk = 0; kk < NBc; ++kk) {
            double best = -1.0;
            int bs = 0;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const double m = fma(p[s][kk].x, p[s][kk].x, p[s][kk].y * p[s][kk].y);
                if (cand[s] && alive[s] && m > best) {
                    best = m;
                    bs = s;
                }
            }
            const int key = __double2hiint(best);
            const int kmax = __reduce_max_sync(kFull, key);
            const unsigned ball = __ballot_sync(kFull, key == kmax);
            const int pl = __ffs(ball) - 1;
            if (kmax <= 0) singular = 1;  // |pivot|^2 is zero/denormal (or no candidate)
            const int ps = (SLOTS > 1) ? __shfl_sync(kFull, bs, pl) : 0;
            double2 u[NBc + 1];
#pragma unroll
            for (int j = kk; j <= NBc; ++j) {
                double2 src = p[0][j];
                if (SLOTS > 1 && ps) src = p[SLOTS - 1][j];
                u[j] = shfl2(src, pl);
            }
            int mypos = pos[0];
            if (SLOTS > 1 && ps) mypos = pos[SLOTS - 1];
            const int pp = __shfl_sync(kFull, mypos, pl);
            const int target = J0 + kk;
            if (MT > 0 && pp != target) {  // warp-uniform: interchange rows `target` and `pp` in the trailing columns
#pragma unroll
                for (int c0 = 0; c0 < MT; c0 += 32) {
                    const int c = c0 + lane;
                    if ((c0 + 31 < MT) || (c < MT)) {
                        const double2 t1 = M[target * LD + R0 + c];
                        const double2 t2 = M[pp * LD + R0 + c];
                        M[target * LD + R0 + c] = t2;
                        M[pp * LD + R0 + c] = t1;
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
                if (pos[s] == target) pos[s] = pp;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
                if (lane == pl && s == ps) {
                    pos[s] = target;
                    alive[s] = false;
                }
            const double2 rinv = crecip(u[kk]);
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                double2 l = make_double2(0.0, 0.0);
                if (alive[s]) {
                    l = cmul(p[s][kk], rinv);
                    p[s][kk] = l;
                }
#pragma unroll
                for (int j = kk + 1; j <= NBc; ++j) cfnma(p[s][j], l, u[j]);
            }
        }
        if constexpr (MT == 0) {
            // last panel: the z row (row N) now holds the corner in its y slot
            constexpr int zl = (N - J0) % 32, zs = (N - J0) / 32;
            corner = shfl2(p[zs][NBc], zl);
        } else {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            if (valid[s]) {
#pragma unroll
                for (int j = 0; j < NBc; ++j) M[pos[s] * LD + J0 + j] = p[s][j];
                M[pos[s] * LD + N] = p[s][NBc];
            }
        }
        __syncwarp();
        // ------------------------------------------------------------------ phase U
#pragma unroll
        for (int c0 = 0; c0 < MT; c0 += 32) {
            const int c = c0 + lane;
            const bool act = (c0 + 31 < MT) || (c < MT);
            const int col = R0 + (act ? c : 0);
            double2 a[NBc];
#pragma unroll
            for (int kk = 0; kk < NBc; ++kk) a[kk] = M[(J0 + kk) * LD + col];
#pragma unroll
            for (int kk = 1; kk < NBc; ++kk)
#pragma unroll
                for (int k2 = 0; k2 < kk; ++k2) cfnma(a[kk], M[(J0 + kk) * LD + J0 + k2], a[k2]);
            double2 z = M[N * LD + col];
#pragma unroll
            for (int kk = 0; kk < NBc; ++kk) cfnma(z, M[N * LD + J0 + kk], a[kk]);
            if (act) {
#pragma unroll
                for (int kk = 0; kk < NBc; ++kk) M[(J0 + kk) * LD + col] = a[kk];
                M[N * LD + col] = z;
            }
        }
        __syncwarp();
        // ------------------------------------------------------------------ phase T
        trailing_update<N, J0, NBc, 4, 4>(M, lane);
        __syncwarp();
        Stage<N, R0>::run(M, lane, singular, corner);
        }
    }
};

// -b^T A^-1 b for the bordered matrix held in M (destroyed).  All lanes get the result.
template <int N>
__device__ __forceinline__ double2 bordered_schur(double2* __restrict__ M, int lane, int& singular) {
    double2 corner = make_double2(0.0, 0.0);
    Stage<N, 0>::run(M, lane, singular, corner);
    return corner;
}

}  // namespace jitr
