// Warp-per-system blocked LU (complex128, partial pivoting) of the BORDERED matrix
//
//        M = [ A   0  b ]     rows 0..N-1   : the system (N = nbasis)
//            [ 0   0  0 ]     rows N..P-1   : zero padding up to P = 8*ceil(N/8)
//            [ b^T 0  0 ]     row  P        : the "z" row
//
// stored row-major in shared memory with leading dimension LD = P + 1 (odd: a column walk by the 32
// lanes is bank-conflict free), column P holding the right-hand side.  The first N columns are
// eliminated with the pivot search restricted to the first N rows; the corner M[P][P] that is left
// is the Schur complement  -b^T A^-1 b , i.e. minus the (scaled) R-matrix that rmatrix/core.py:7-22
// obtains through the explicit inverse.  Neither L nor U is needed afterwards: no back-substitution.
//
// One warp owns one system; right-looking, panel width 8, three phases per panel.  The code is a
// RUNTIME loop over panels with nbasis a runtime value: the first version of this file unrolled every
// panel at compile time and ran instruction-fetch bound (400 KB of SASS, I-cache hit rate 38 %, ncu
// profiles/r01_v1_*); this one is ~1.5k instructions and serves every nbasis <= 60.
//   P  panel factorisation with the panel in REGISTERS, one matrix row per lane (two row slots while
//      more than 32 rows are alive): pivot search = REDUX.MAX on the high word of |a|^2 + ballot,
//      pivot-row broadcast = shuffles, rank-1 updates = DFMA on registers.  Column P (the RHS y) rides
//      along as a ninth panel column, row P (z = U^-T b) as one more row that is never a pivot
//      candidate.  Row interchanges of the trailing columns are done in shared memory as soon as the
//      pivot is known (lane c always owns column c, so no barrier is needed).
//   U  U12 = L11^-1 A12: lane c owns trailing column c, 8-step forward substitution in registers
//      (L11 read as shared-memory broadcasts); the z row is updated here too.
//   T  A22 -= L21 U12 on the FP64 tensor-core path: 8x8 complex tiles, K = 8 = two m8n8k4 DMMA steps,
//      four real DMMAs per complex step; a warp register-blocks 1x2 tiles.
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
__device__ __forceinline__ double dneg(double x) {
    return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));
}
// D(8x8) += A(8x4, row) * B(4x8, col): a = A[lane/4][lane%4], b = B[lane%4][lane/4], c{0,1} = C[lane/4][2(lane%4)+{0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct LuGeom {
    int N;    // nbasis (real rows/columns)
    int P;    // N rounded up to a multiple of 8; z row = row P, rhs column = column P
    int LD;   // P + 1
};

// ---- phase T: A22 -= L21 * U12, 8x8 complex tiles on DMMA --------------------------------------------
__device__ __forceinline__ void trailing_update_dmma(double2* __restrict__ M, const LuGeom g, int J0, int lane) {
    const int R0 = J0 + 8;
    const int nt = (g.P - R0) >> 3;  // tiles per dimension
    const int LD = g.LD;
    const int fr = lane >> 2, fk = lane & 3;
    for (int tr = 0; tr < nt; ++tr) {
        // A fragments of this tile row (both k-steps): -L and the sign-flipped parts, kept in registers
        const int arow = (R0 + 8 * tr + fr) * LD + J0 + fk;
        const double2 l0 = M[arow], l1 = M[arow + 4];
        const double nl0x = dneg(l0.x), nl0y = dneg(l0.y), nl1x = dneg(l1.x), nl1y = dneg(l1.y);
        for (int tc = 0; tc < nt; tc += 2) {
            const bool two = tc + 1 < nt;  // warp-uniform
            const int bcol = R0 + 8 * tc + fr;
            const int brow = (J0 + fk) * LD;
            const double2 u0 = M[brow + bcol], u1 = M[brow + 4 * LD + bcol];
            const int crow = (R0 + 8 * tr + fr) * LD + R0 + 8 * tc + 2 * fk;
            double2 ca = M[crow], cb = M[crow + 1];
            // C_re += (-Lre) Ure + (Lim) Uim ;  C_im += (-Lre) Uim + (-Lim) Ure
            dmma884(ca.x, cb.x, nl0x, u0.x);
            dmma884(ca.y, cb.y, nl0x, u0.y);
            dmma884(ca.x, cb.x, l0.y, u0.y);
            dmma884(ca.y, cb.y, nl0y, u0.x);
            dmma884(ca.x, cb.x, nl1x, u1.x);
            dmma884(ca.y, cb.y, nl1x, u1.y);
            dmma884(ca.x, cb.x, l1.y, u1.y);
            dmma884(ca.y, cb.y, nl1y, u1.x);
            M[crow] = ca;
            M[crow + 1] = cb;
            if (two) {
                const double2 v0 = M[brow + bcol + 8], v1 = M[brow + 4 * LD + bcol + 8];
                double2 cc = M[crow + 8], cd = M[crow + 9];
                dmma884(cc.x, cd.x, nl0x, v0.x);
                dmma884(cc.y, cd.y, nl0x, v0.y);
                dmma884(cc.x, cd.x, l0.y, v0.y);
                dmma884(cc.y, cd.y, nl0y, v0.x);
                dmma884(cc.x, cd.x, nl1x, v1.x);
                dmma884(cc.y, cd.y, nl1x, v1.y);
                dmma884(cc.x, cd.x, l1.y, v1.y);
                dmma884(cc.y, cd.y, nl1y, v1.x);
                M[crow + 8] = cc;
                M[crow + 9] = cd;
            }
        }
    }
}

// ---- the whole factorisation; returns the corner -b^T A^-1 b to every lane ---------------------------
__device__ __forceinline__ double2 bordered_schur(double2* __restrict__ M, const LuGeom g, int lane, int& singular) {
    const int N = g.N, P = g.P, LD = g.LD;
    double2 corner = make_double2(0.0, 0.0);
    for (int J0 = 0; J0 < N; J0 += 8) {
        const int nreal = N - J0;              // real rows still alive (before this panel)
        const int nb = nreal < 8 ? nreal : 8;  // panel width
        const int rows = nreal + 1;            // + z row
        const bool two = rows > 32;            // warp-uniform: second row slot in use
        const int MTr = N - J0 - 8;            // real trailing columns (<= 0 on the last panel)
        const int R0 = J0 + 8;
        // ------------------------------------------------------------------ phase P
        double2 p[2][9];
        int pos[2];
        bool alive[2], cand[2], valid[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int i = 32 * s + lane;
            valid[s] = i < rows;
            cand[s] = i < nreal;
            alive[s] = valid[s];
            pos[s] = cand[s] ? J0 + i : P;  // the z row sits at row P
            if (s == 0 || two) {
                const int base = (valid[s] ? pos[s] : J0) * LD;
#pragma unroll
                for (int j = 0; j < 8; ++j) p[s][j] = M[base + J0 + j];
                p[s][8] = M[base + P];
            } else {
#pragma unroll
                for (int j = 0; j < 9; ++j) p[s][j] = make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            if (kk < nb) {  // warp-uniform
                double best = -1.0;
                int bs = 0;
                {
                    const double m = fma(p[0][kk].x, p[0][kk].x, p[0][kk].y * p[0][kk].y);
                    if (cand[0] && alive[0] && m > best) best = m;
                }
                if (two) {
                    const double m = fma(p[1][kk].x, p[1][kk].x, p[1][kk].y * p[1][kk].y);
                    if (cand[1] && alive[1] && m > best) {
                        best = m;
                        bs = 1;
                    }
                }
                const int key = __double2hiint(best);
                const int kmax = __reduce_max_sync(kFull, key);
                const unsigned ball = __ballot_sync(kFull, key == kmax);
                const int pl = __ffs(ball) - 1;
                if (kmax <= 0) singular = 1;  // |pivot|^2 is zero/denormal (or no candidate)
                const int ps = two ? __shfl_sync(kFull, bs, pl) : 0;
                double2 u[9];
#pragma unroll
                for (int j = kk; j < 9; ++j) {
                    double2 src = p[0][j];
                    if (ps) src = p[1][j];
                    u[j] = shfl2(src, pl);
                }
                const int pp = __shfl_sync(kFull, ps ? pos[1] : pos[0], pl);
                const int target = J0 + kk;
                if (MTr > 0 && pp != target) {  // interchange rows `target` and `pp` in the trailing columns
                    for (int c = lane; c < MTr; c += 32) {
                        const double2 t1 = M[target * LD + R0 + c];
                        const double2 t2 = M[pp * LD + R0 + c];
                        M[target * LD + R0 + c] = t2;
                        M[pp * LD + R0 + c] = t1;
                    }
                }
#pragma unroll
                for (int s = 0; s < 2; ++s)
                    if (pos[s] == target) pos[s] = pp;
                if (lane == pl) {
                    if (ps) { pos[1] = target; alive[1] = false; }
                    else    { pos[0] = target; alive[0] = false; }
                }
                const double2 rinv = crecip(u[kk]);
                {
                    double2 l = make_double2(0.0, 0.0);
                    if (alive[0]) {
                        l = cmul(p[0][kk], rinv);
                        p[0][kk] = l;
                    }
#pragma unroll
                    for (int j = kk + 1; j < 9; ++j) cfnma(p[0][j], l, u[j]);
                }
                if (two) {
                    double2 l = make_double2(0.0, 0.0);
                    if (alive[1]) {
                        l = cmul(p[1][kk], rinv);
                        p[1][kk] = l;
                    }
#pragma unroll
                    for (int j = kk + 1; j < 9; ++j) cfnma(p[1][j], l, u[j]);
                }
            }
        }
        if (MTr <= 0) {
            // last panel: the z row now holds the corner in its y slot
            const int zl = nreal & 31, zs = nreal >> 5;
            corner = shfl2(zs ? p[1][8] : p[0][8], zl);
            break;
        }
        if (valid[0]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) M[pos[0] * LD + J0 + j] = p[0][j];
            M[pos[0] * LD + P] = p[0][8];
        }
        if (two && valid[1]) {
#pragma unroll
            for (int j = 0; j < 8; ++j) M[pos[1] * LD + J0 + j] = p[1][j];
            M[pos[1] * LD + P] = p[1][8];
        }
        __syncwarp();
        // ------------------------------------------------------------------ phase U
        for (int c0 = 0; c0 < MTr; c0 += 32) {
            const int c = c0 + lane;
            const bool act = c < MTr;
            const int col = R0 + (act ? c : 0);
            double2 a[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) a[kk] = M[(J0 + kk) * LD + col];
#pragma unroll
            for (int kk = 1; kk < 8; ++kk)
#pragma unroll
                for (int k2 = 0; k2 < kk; ++k2) cfnma(a[kk], M[(J0 + kk) * LD + J0 + k2], a[k2]);
            double2 z = M[P * LD + col];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) cfnma(z, M[P * LD + J0 + kk], a[kk]);
            if (act) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) M[(J0 + kk) * LD + col] = a[kk];
                M[P * LD + col] = z;
            }
        }
        __syncwarp();
        // ------------------------------------------------------------------ phase T
        trailing_update_dmma(M, g, J0, lane);
        __syncwarp();
    }
    return corner;
}

}  // namespace jitr
