// Warp-per-system blocked LU (complex128, partial pivoting) of the BORDERED matrix
//
//        M = [ A   0  b ]     rows 0..N-1   : the system (N = nbasis)
//            [ 0   I  0 ]     rows N..P-1   : identity padding up to P = 8*ceil(N/8)
//            [ b^T 0  0 ]     row  P        : the "z" row
//
// stored row-major in shared memory with leading dimension LD = P + 1 (odd: a column walk by the 32
// lanes is bank-conflict free), column P holding the right-hand side.  The first N columns are
// eliminated with the pivot search restricted to the first P rows (the identity padding makes every
// panel a full 8 columns and changes nothing else); the corner M[P][P] that is left
// is the Schur complement  -b^T A^-1 b , i.e. minus the (scaled) R-matrix that rmatrix/core.py:7-22
// obtains through the explicit inverse.  Neither L nor U is needed afterwards: no back-substitution.
//
// One warp owns one system; right-looking, panel width 8, three phases per panel, nbasis a runtime
// value (P + 1 <= 96 rows: up to three row slots per lane).
//   P  panel factorisation with the panel in REGISTERS, one matrix row per lane (a second row slot
//      while more than 32 rows are alive).  Pivot search: |re|+|im| (LAPACK izamax), REDUX.MAX on the
//      high word + ballot; pivot-row broadcast by SHFL (measured on B200: a shared-memory store costs
//      4 SM-cycles per 128-bit instruction whatever the number of active lanes, a SHFL 1).  Column P
//      (the RHS y) rides along as a ninth panel column, row P (z = U^-T b) as one more row that is
//      never a pivot candidate.  Row interchanges of the trailing columns are done in shared memory
//      as soon as the pivot is known (lane c always owns column c, so no barrier is needed).
//   U  U12 = L11^-1 A12: lane c owns trailing column c, 8-step forward substitution in registers
//      (L11 read as shared-memory broadcasts); the z row is updated here too.
//   T  A22 -= L21 U12 on the FP64 tensor-core path (DMMA m8n8k4) in the REAL representation of the
//      complex product: an 8 x 4 complex half-tile of C is an 8 x 8 real tile whose accumulator
//      fragment (c0, c1) is exactly one interleaved complex number, so C moves between shared memory
//      and the accumulators with one 128-bit access and no register shuffling; the 8 complex = 16
//      real contraction indices take four DMMAs.
#pragma once
#include <cuda_runtime.h>

#ifndef JITR_PH
#define JITR_PH_INIT
#define JITR_PH(k)
#endif

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
// 1/x to ~1 ulp without the slow-path code of the IEEE division: MUFU.RCP64H seed (2^-23) + two
// Newton steps.  x must be finite, normal and non-zero (callers flag zero pivots separately).
__device__ __forceinline__ double rcp_newton(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
// the same with ONE cubic (Halley) step, r (1 + e + e^2), e = 1 - x r: three dependent FMAs instead of four (2^-69 after the
// 2^-23 seed) -- for the pivot chain of the packed path, where the reciprocal is the longest link
__device__ __forceinline__ double rcp_halley(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double2 crecip_fast(double2 a) {
    const double d = rcp_newton(fma(a.x, a.x, a.y * a.y));
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

// ---- phase P: factorise eight panel columns (+ the RHS column) held in registers -------------------
// p[s][0..7] panel columns and p[s][8] the RHS of the row in slot s of this lane; pos[s] the row's
// current position (LAPACK-style interchanges are applied to positions); cand[s] whether the row is
// still a pivot candidate.  On return p[s][k] (k < 8) holds the row's multipliers, p[s][8] its updated
// RHS; a row that became pivot kk has pos = `first` + kk and keeps its multipliers p[0..kk-1].
// swap (warp-uniform): also interchange the trailing columns (R0.. R0+MTr-1) of the two rows in shared memory.
template <int NS>
__device__ __forceinline__ void panel_eliminate(double2 (&p)[NS][9], int (&pos)[NS], bool (&cand)[NS],
                                                double2* __restrict__ M, const int LD, const int first, const int R0,
                                                const int MTr, const bool swap, const int lane, int& singular) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        // pivot search over the candidate rows: max |re| + |im|
        // (a NaN entry counts as the largest: the column is then eliminated like any other and the system is flagged
        // non-finite at the end; a lane without candidate rows never wins the search)
        double best = -1.0;
        int bs = 0;
        bool any = false;
        double2 cbest = p[0][kk];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const double m = fabs(p[s][kk].x) + fabs(p[s][kk].y);
            if (cand[s] && !(m <= best)) {
                best = m;
                bs = s;
                cbest = p[s][kk];
                any = true;
            }
        }
        // speculative reciprocal of this lane's own candidate: runs beside the search instead of
        // after the broadcast (it is the longest dependent chain of the step)
        const double2 myrinv = crecip_fast(cbest);
        const int key = any ? (__double2hiint(best) & 0x7fffffff) : -1;  // the GPU's canonical FP64 NaN is negative
        const int kmax = __reduce_max_sync(kFull, key);
        const int pl = __ffs(__ballot_sync(kFull, key == kmax)) - 1;
        if (kmax <= 0) singular = 1;  // pivot is zero/denormal (or there is no finite candidate)
        // (slot, position) of the pivot row, one shuffle
        int mypk = pos[0];
#pragma unroll
        for (int s = 1; s < NS; ++s)
            if (bs == s) mypk = pos[s] | (s << 8);
        const int pk = __shfl_sync(kFull, mypk, pl);
        const int ps = pk >> 8;  // warp-uniform
        const int pp = pk & 0xff;
        const int target = first + kk;
        // broadcast 1/pivot and the pivot row (columns kk+1.. and the RHS); the row is final from now on
        const double2 rinv = shfl2(myrinv, pl);
        double2 u[9];
#pragma unroll
        for (int j = kk + 1; j < 9; ++j) {
            double2 src = p[0][j];
#pragma unroll
            for (int s = 1; s < NS; ++s)
                if (ps == s) src = p[s][j];
            u[j] = shfl2(src, pl);
        }
        if (lane == pl) {  // the pivot row leaves the candidate set and is parked at `target`
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (ps == s) {
                    cand[s] = false;
                    pos[s] = -1;
                }
        }
        if (swap && pp != target) {  // interchange rows `target` and `pp` in the trailing columns
            for (int c = lane; c < MTr; c += 32) {
                const double2 t1 = M[target * LD + R0 + c];
                const double2 t2 = M[pp * LD + R0 + c];
                M[target * LD + R0 + c] = t2;
                M[pp * LD + R0 + c] = t1;
            }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (pos[s] == target) pos[s] = pp;  // whoever lived at `target` moves to the pivot's old home
        if (lane == pl) {
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (ps == s) pos[s] = target;
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            // A row that already is a pivot keeps its multipliers p[0..own kk-1] (all that phase U
            // needs from it); what it computes from here on is never used.
            const double2 l = cmul(p[s][kk], rinv);
            p[s][kk] = l;
#pragma unroll
            for (int j = kk + 1; j < 9; ++j) cfnma(p[s][j], l, u[j]);
        }
    }
}

// Panel J0 of the matrix in shared memory.  NS = number of row slots per lane (rows alive, z row
// included, <= 32 NS).  Returns true when this was the last panel; `corner` then holds -b^T A^-1 b.
template <int NS>
__device__ __forceinline__ void panel_load(double2 (&p)[NS][9], int (&pos)[NS], bool (&cand)[NS], bool (&wb)[NS],
                                           const double2* __restrict__ M, const LuGeom g, const int J0, const int lane) {
    const int P = g.P, LD = g.LD;
    const int nreal = P - J0;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int i = 32 * s + lane;
        cand[s] = i < nreal;
        wb[s] = i <= nreal;              // row exists: real rows + the z row (i == nreal)
        pos[s] = cand[s] ? J0 + i : P;   // the z row sits at row P
        const int base = (wb[s] ? pos[s] : J0) * LD;
#pragma unroll
        for (int j = 0; j < 8; ++j) p[s][j] = M[base + J0 + j];
        p[s][8] = M[base + P];
    }
}
template <int NS>
__device__ __forceinline__ void panel_store(const double2 (&p)[NS][9], const int (&pos)[NS], const bool (&wb)[NS],
                                            double2* __restrict__ M, const LuGeom g, const int J0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (wb[s]) {
            double2* dst = M + pos[s] * g.LD;
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[J0 + j] = p[s][j];
            dst[g.P] = p[s][8];
        }
    }
}
template <int NS>
__device__ __forceinline__ bool panel_phase(double2* __restrict__ M, const LuGeom g, const int J0, const int lane,
                                            int& singular, double2& corner) {
    const int nreal = g.P - J0;  // candidate rows still alive (before this panel), a multiple of 8
    const int MTr = nreal - 8;   // trailing columns (0 on the last panel)
    double2 p[NS][9];
    int pos[NS];
    bool cand[NS], wb[NS];
    panel_load<NS>(p, pos, cand, wb, M, g, J0, lane);
    panel_eliminate<NS>(p, pos, cand, M, g.LD, J0, J0 + 8, MTr, MTr > 0, lane, singular);
    if (MTr <= 0) {
        // last panel: the z row (slot 0 of lane 8) now holds the corner in its y slot
        corner = shfl2(p[0][8], nreal & 31);
        return true;
    }
    panel_store<NS>(p, pos, wb, M, g, J0);
    return false;
}

// ---- phase U: U12 = L11^-1 A12 and the z row --------------------------------------------------------
__device__ __forceinline__ void upper_phase(double2* __restrict__ M, const LuGeom g, const int J0, const int lane) {
    const int P = g.P, LD = g.LD;
    const int MTr = P - J0 - 8;
    const int R0 = J0 + 8;
#pragma unroll 1
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
            for (int kk = 1; kk < 8; ++kk) M[(J0 + kk) * LD + col] = a[kk];
            M[P * LD + col] = z;
        }
    }
}

// ---- phase T: A22 -= L21 * U12 on DMMA, real representation of the complex product -------------------
// Fragment <-> matrix mapping (lane = 4 fr + fk):
//   C  8 rows x 4 complex columns: row rho(fr) = (fr >> 1) + 4 (fr & 1), complex column fk; the accumulator
//      pair (c0, c1) is (re, im).  rho puts the two rows of a quarter-warp 4 rows apart, which makes the
//      128-bit accesses conflict free with the odd leading dimension.
//   contraction index of step s (0..3): complex k = s + 4 (fk >> 1), part = fk & 1 (0 re, 1 im)
//   A  -L[row rho(fr)][k].part
//   B  column n = fr -> complex column fr >> 1, pn = fr & 1:  component (part ^ pn) of U[k][col],
//      negated when part = 1 and pn = 0      ( [re im] += [Lre Lim] [[Ure Uim], [-Uim Ure]] )
struct TFrag {
    int a_off, b_off, bflip, c_off;
};
// NQ tile columns starting at tile column tc0, all nt tile rows
template <int NQ>
__device__ __forceinline__ void trailing_block(double2* __restrict__ M, const TFrag f, const int LD, const int J0,
                                               const int nt, const int tc0) {
    const int R0 = J0 + 8;
    const double* Md = reinterpret_cast<const double*>(M);
    // B fragments of the NQ tile columns (2 half-tiles x 4 steps each), kept in registers over the tile rows
    double b[NQ][2][4];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double v = Md[f.b_off + 2 * (s * LD + R0 + 8 * (tc0 + q) + 4 * h)];
                b[q][h][s] = __hiloint2double(__double2hiint(v) ^ f.bflip, __double2loint(v));
            }
#pragma unroll 1
    for (int tr = 0; tr < nt; ++tr) {
        const int rbase = (R0 + 8 * tr) * LD;
        double a[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) a[s] = dneg(Md[f.a_off + 2 * (rbase + s)]);
        double2* Crow = M + rbase + f.c_off + 8 * tc0;
        double2 c[NQ][2];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            c[q][0] = Crow[8 * q];
            c[q][1] = Crow[8 * q + 4];
        }
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                dmma884(c[q][0].x, c[q][0].y, a[s], b[q][0][s]);
                dmma884(c[q][1].x, c[q][1].y, a[s], b[q][1][s]);
            }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            Crow[8 * q] = c[q][0];
            Crow[8 * q + 4] = c[q][1];
        }
    }
}

__device__ __forceinline__ void trailing_update_dmma(double2* __restrict__ M, const LuGeom g, int J0, int lane) {
    const int R0 = J0 + 8;
    const int nt = (g.P - R0) >> 3;  // tiles per dimension
    const int LD = g.LD;
    const int fr = lane >> 2, fk = lane & 3;
    const int rho = (fr >> 1) + 4 * (fr & 1);
    const int part = fk & 1, pn = fr & 1;
    TFrag f;
    // per-lane element offsets (in doubles) of the A and B operands of step s = 0 (then + 2 s / + 2 s LD)
    f.a_off = 2 * (rho * LD + J0 + 4 * (fk >> 1)) + part;
    f.b_off = 2 * ((J0 + 4 * (fk >> 1)) * LD + (fr >> 1)) + (part ^ pn);
    f.bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
    f.c_off = rho * LD + R0 + fk;
    int tc0 = 0;
#pragma unroll 1
    for (; tc0 + 4 <= nt; tc0 += 4) trailing_block<4>(M, f, LD, J0, nt, tc0);
    const int rem = nt - tc0;  // warp-uniform
    if (rem == 3) trailing_block<3>(M, f, LD, J0, nt, tc0);
    else if (rem == 2) trailing_block<2>(M, f, LD, J0, nt, tc0);
    else if (rem == 1) trailing_block<1>(M, f, LD, J0, nt, tc0);
}

// ---- the whole factorisation; returns the corner -b^T A^-1 b to every lane ---------------------------
// MAXNS: largest number of row slots the instantiation supports (P + 1 <= 32 MAXNS).
template <int MAXNS>
__device__ __forceinline__ double2 bordered_schur(double2* __restrict__ M, const LuGeom g, int lane, int& singular,
                                                  const int J0start = 0) {
    double2 corner = make_double2(0.0, 0.0);
    JITR_PH_INIT
#pragma unroll 1
    for (int J0 = J0start; J0 < g.P; J0 += 8) {
        const int rows = g.P - J0 + 1;
        bool last;
        if (MAXNS >= 3 && rows > 64) last = panel_phase<(MAXNS >= 3 ? 3 : 1)>(M, g, J0, lane, singular, corner);
        else if (MAXNS >= 2 && rows > 32) last = panel_phase<(MAXNS >= 2 ? 2 : 1)>(M, g, J0, lane, singular, corner);
        else last = panel_phase<1>(M, g, J0, lane, singular, corner);
        JITR_PH(0)
        if (last) break;
        __syncwarp();
        upper_phase(M, g, J0, lane);
        __syncwarp();
        JITR_PH(1)
        trailing_update_dmma(M, g, J0, lane);
        __syncwarp();
        JITR_PH(2)
    }
    return corner;
}

}  // namespace jitr
