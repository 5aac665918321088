// Symmetric fast path: threshold-pivoted elimination that keeps the complex symmetry of the system.
//
// Every system of the elastic sweep is complex SYMMETRIC,  a^2 A = That + diag(d),  and so is the bordered
// matrix [[A, b], [b^T, 0]].  As long as the diagonal entry passes the threshold-pivoting test
//        |a_kk| >= tau * max_{i >= k} |a_ik|          (tau = 2^-JITR_SYM_TAU_LOG2, |.| = max(|re|, |im|))
// it is taken as the pivot (classical threshold partial pivoting: every multiplier is bounded by 1/tau),
// no rows move, and the elimination is A = L D L^T:
//   * the pivot row is the transpose of the pivot column (no row search result to fetch, no interchanges),
//   * U12 = D L21^T comes for free (the unscaled column entries), so there is no triangular solve (phase U),
//   * only the lower triangle of the trailing matrix is updated (20 instead of 30 DMMA tiles at N = 40),
//   * the border needs no z row: the corner is  -sum_k y_k^2 / d_k.
// The first time a diagonal entry FAILS the test the factorisation stops and reports it; the caller then
// solves that system with the general partial-pivoting path (lu_fast.cuh).  On the KDUQ sweep the
// kinetic diagonal dominates: 0.013 % of the systems fall back at tau = 1/64 (0.2 % at 1/16), element
// growth of the unpivoted elimination is 1.0 and the corner agrees with LAPACK to <= 2e-11 either way
// (profiles/r01_pivoting_stats.md).
//
// Layout and phases follow lu_fast.cuh / lu_warp.cuh: first panel straight from That (the system never
// exists as a whole), Schur complement W of order n = PT - 8 with its RHS column in shared memory,
// panels of 8 columns in registers (one row per lane), trailing update on DMMA in the real representation.
#pragma once
#include "lu_warp.cuh"

#ifndef JITR_SYM_TAU_LOG2
#define JITR_SYM_TAU_LOG2 6
#endif

namespace jitr {

template <int PT>
struct SymGeom {
    static constexpr int n = PT - 8;       // order of the Schur complement kept in W
    static constexpr int LDW = n + 1;      // odd leading dimension; column n = RHS (no z row)
    static constexpr int NT = n / 8;
    static constexpr int NS0 = PT > 32 ? 2 : 1;
    static constexpr int WSZ = n * LDW;
    // scratch inside W while the first panel is in flight:
    //   WB  the unscaled columns of ALL rows 0..PT-1 (rows < 8: the pivot block), at the start of W
    //   LB  the multipliers L21 of rows 8..PT-1, at the END of W: tile row tr of the Schur complement,
    //       stored to W rows 8 tr .. 8 tr + 7, then only overwrites LB rows of tile rows <= tr, so the
    //       A fragments can be fetched one tile row at a time instead of being held in registers
    static constexpr int LB_LD = 9;
    static constexpr int WB_OFF = 0;
    static constexpr int LB_MIN = n * (n - 8);  // = WSZ - n * LB_LD
    static constexpr int LB_OFF = LB_MIN > PT * LB_LD ? LB_MIN : PT * LB_LD;
    static constexpr int WALLOC = LB_OFF + n * LB_LD;  // >= WSZ
    static_assert(n >= 8 && n <= 32, "symmetric fast path covers 16 <= PT <= 40");
};

// One panel of eight columns, rows i = 32 s + lane (relative to the panel start), nrows of them alive.
// p[s][0..7] panel columns, p[s][8] the RHS.  The unscaled column entries (= the pivot-row block by
// symmetry) are handed to `store_w(s, kk, value)` as they become final -- for every row below the
// diagonal -- and the pivot row is read back with `load_u(kk, j)` (entry of row j > kk, a shared-memory
// broadcast: 2 SM-cycles against the 4 of a 128-bit shuffle; the store is issued anyway).  On return
// p[s][k] are the multipliers and p[s][8] the updated RHS.  cacc accumulates the corner; fail is set when
// a diagonal entry fails the threshold test (or is zero / not finite).
// BC: 1/d and y of the pivot row travel through shared memory as well -- store_w(-1 / -2, kk, v) by lane kk, load_u(kk, -1 / -2)
// by all -- instead of two 128-bit shuffles: 8 LSU wavefronts -> 4 per column; the LSU pipe is the second-busiest unit of the
// N = 40 sweep (69 %), measured +2.1 % (profiles/r02_sweep_variants.md).
template <int NS, bool BC = false, class StoreW, class LoadU>
__device__ __forceinline__ void sym_panel_eliminate(double2 (&p)[NS][9], const int nrows, const int lane,
                                                    double2& cacc, int& fail, StoreW store_w, LoadU load_u) {
#ifdef JITR_SYM_FASTCHAIN
    // Variant of the pivot chain (lu_packed.cuh): no column-maximum REDUX -- the threshold test is applied to the multipliers
    // themselves, |Re l|, |Im l| <= 2^TAU for the rows below the pivot -- and  l = (w conj(d)) / |d|^2  so that the complex
    // product runs beside the reciprocal.
    double2 ysave = make_double2(0.0, 0.0), rsave = make_double2(0.0, 0.0);
    int myfail = 0;
    constexpr int kLim = 0x3ff00000 + (JITR_SYM_TAU_LOG2 << 20);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const double2 dmine = p[0][kk];
        const double mys = rcp_halley(fma(dmine.x, dmine.x, dmine.y * dmine.y));
        if (lane == kk) {
            const int m = max(__double2hiint(dmine.x) & 0x7fffffff, __double2hiint(dmine.y) & 0x7fffffff);
            if (m < 0x00100000 || m >= 0x7ff00000) myfail = 1;
            ysave = p[0][8];
            rsave = make_double2(dmine.x * mys, -dmine.y * mys);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) store_w(s, kk, p[s][kk]);
        const double2 piv = shfl2(dmine, kk);
        double2 u[9];
        u[8] = shfl2(p[0][8], kk);
        __syncwarp();
#pragma unroll
        for (int j = kk + 1; j < 8; ++j) u[j] = load_u(kk, j);
        const double sc = __shfl_sync(kFull, mys, kk);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int i = 32 * s + lane;
            const double2 w = p[s][kk];
            const double tx = fma(w.x, piv.x, w.y * piv.y), ty = fma(w.y, piv.x, -(w.x * piv.y));
            const double2 l = make_double2(tx * sc, ty * sc);
            if (i > kk && i < nrows && max(__double2hiint(l.x) & 0x7fffffff, __double2hiint(l.y) & 0x7fffffff) > kLim) myfail = 1;
            p[s][kk] = l;
#pragma unroll
            for (int j = kk + 1; j < 9; ++j) cfnma(p[s][j], l, u[j]);
        }
    }
#else
    // lane k (k < 8) keeps y_k and 1/d_k of "its" pivot for the corner term, formed once per panel
    double2 ysave = make_double2(0.0, 0.0), rsave = make_double2(0.0, 0.0);
    int myfail = 0;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        // column maximum over the rows at and below the diagonal, on the high words of max(|re|, |im|)
        // (integer pipe; within a factor 2 of |re| + |im|, which is all a threshold test needs)
        int key = -1, key0 = 0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int i = 32 * s + lane;
            const int m = max(__double2hiint(p[s][kk].x) & 0x7fffffff, __double2hiint(p[s][kk].y) & 0x7fffffff);
            if (s == 0) key0 = m;
            if (i >= kk && i < nrows) key = max(key, m);
        }
        const double2 myrinv = crecip_fast(p[0][kk]);  // speculative: only lane kk's is used
        const int kmax = __reduce_max_sync(kFull, key);
        // threshold test by the diagonal's own lane: |d| * 2^TAU >= colmax; zero, denormal, Inf or NaN fail
        if (lane == kk) {
            if (key0 < 0x00100000 || key0 + (JITR_SYM_TAU_LOG2 << 20) < kmax || kmax >= 0x7ff00000) myfail = 1;
            ysave = p[0][8];
            rsave = myrinv;
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) store_w(s, kk, p[s][kk]);
        double2 rinv;
        double2 u[9];
        if constexpr (BC) {
            if (lane == kk) {
                store_w(-1, kk, myrinv);
                store_w(-2, kk, p[0][8]);
            }
            __syncwarp();
            rinv = load_u(kk, -1);
            u[8] = load_u(kk, -2);
        } else {
            rinv = shfl2(myrinv, kk);
            u[8] = shfl2(p[0][8], kk);
            __syncwarp();
        }
        // pivot row = transposed pivot column: the entries the rows kk+1..7 have just stored
#pragma unroll
        for (int j = kk + 1; j < 8; ++j) u[j] = load_u(kk, j);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const double2 l = cmul(p[s][kk], rinv);
            p[s][kk] = l;
#pragma unroll
            for (int j = kk + 1; j < 9; ++j) cfnma(p[s][j], l, u[j]);
        }
    }
#endif
    // corner -= sum_k y_k^2 / d_k : one term per lane k < 8 (zero on the others), kept PER LANE across the panels; the owner of
    // cacc sums the eight lanes once, after the last panel (corner_total) -- 16 shuffles per solve instead of per panel
    const double2 term = cmul(cmul(ysave, rsave), ysave);
    cacc.x -= term.x;
    cacc.y -= term.y;
    if (__any_sync(kFull, myfail)) fail = 1;
}

// Sum of the per-lane corner terms of the panels (lanes 0..7), on every lane.
__device__ __forceinline__ double2 corner_total(double2 t) {
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        t.x += __shfl_xor_sync(kFull, t.x, o);
        t.y += __shfl_xor_sync(kFull, t.y, o);
    }
    return shfl2(t, 0);  // lanes >= 8 hold zeros: take the sum from the first group
}

// lower-triangle trailing update of the Schur complement in W after panel J0 (nt <= 3 tile rows)
__device__ __forceinline__ void sym_trailing_update(double2* __restrict__ M, const int LD, const int J0, const int nt,
                                                    const int lane) {
    const int R0 = J0 + 8;
    const int fr = lane >> 2, fk = lane & 3;
    const int rho = (fr >> 1) + 4 * (fr & 1);
    const int part = fk & 1, pn = fr & 1;
    const double* Md = reinterpret_cast<const double*>(M);
    const int a_off = 2 * (rho * LD + J0 + 4 * (fk >> 1)) + part;
    const int b_off = 2 * ((J0 + 4 * (fk >> 1)) * LD + (fr >> 1)) + (part ^ pn);
    const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
    const int c_off = rho * LD + R0 + fk;
    double b[3][2][4];
#pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < nt) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const double v = Md[b_off + 2 * (s * LD + R0 + 8 * q + 4 * h)];
                    b[q][h][s] = __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v));
                }
        }
#pragma unroll 1
    for (int tr = 0; tr < nt; ++tr) {
        const int rbase = (R0 + 8 * tr) * LD;
        double a[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) a[s] = dneg(Md[a_off + 2 * (rbase + s)]);
        double2* Crow = M + rbase + c_off;
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (q <= tr) {  // lower triangle (diagonal tiles in full)
                double2 c0 = Crow[8 * q], c1 = Crow[8 * q + 4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    dmma884(c0.x, c0.y, a[s], b[q][0][s]);
                    dmma884(c1.x, c1.y, a[s], b[q][1][s]);
                }
                Crow[8 * q] = c0;
                Crow[8 * q + 4] = c1;
            }
    }
}

// lower-triangle trailing update after panel J0 of an n x n matrix of any order (tile columns in blocks of 3)
__device__ __forceinline__ void sym_trailing_update_blocks(double2* __restrict__ M, const int LD, const int J0,
                                                           const int nt, const int lane) {
    const int R0 = J0 + 8;
    const int fr = lane >> 2, fk = lane & 3;
    const int rho = (fr >> 1) + 4 * (fr & 1);
    const int part = fk & 1, pn = fr & 1;
    const double* Md = reinterpret_cast<const double*>(M);
    const int a_off = 2 * (rho * LD + J0 + 4 * (fk >> 1)) + part;
    const int b_off = 2 * ((J0 + 4 * (fk >> 1)) * LD + (fr >> 1)) + (part ^ pn);
    const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
    const int c_off = rho * LD + R0 + fk;
#pragma unroll 1
    for (int q0 = 0; q0 < nt; q0 += 3) {
        double b[3][2][4];
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (q0 + q < nt) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const double v = Md[b_off + 2 * (s * LD + R0 + 8 * (q0 + q) + 4 * h)];
                        b[q][h][s] = __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v));
                    }
            }
#pragma unroll 1
        for (int tr = q0; tr < nt; ++tr) {
            const int rbase = (R0 + 8 * tr) * LD;
            double a[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) a[s] = dneg(Md[a_off + 2 * (rbase + s)]);
            double2* Crow = M + rbase + c_off + 8 * q0;
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (q0 + q <= tr) {  // lower triangle (diagonal tiles in full)
                    double2 c0 = Crow[8 * q], c1 = Crow[8 * q + 4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        dmma884(c0.x, c0.y, a[s], b[q][0][s]);
                        dmma884(c1.x, c1.y, a[s], b[q][1][s]);
                    }
                    Crow[8 * q] = c0;
                    Crow[8 * q + 4] = c1;
                }
        }
    }
}

// One panel (columns J0..J0+7) of the threshold-pivoted symmetric elimination of a matrix ASSEMBLED in shared
// memory: rows/columns 0..P-1, RHS in column P, leading dimension LD (the layout of lu_warp.cuh without its z
// row).  NS row slots per lane.  Returns true after the last panel.
template <int NS>
__device__ __forceinline__ bool sym_panel_inplace(double2* __restrict__ M, const int P, const int LD, const int J0,
                                                  const int lane, double2& cacc, int& fail) {
    const int nrows = P - J0;
    double2 p[NS][9];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int i = 32 * s + lane;
        const int row = J0 + (i < nrows ? i : 0);
#pragma unroll
        for (int j = 0; j < 8; ++j) p[s][j] = M[row * LD + J0 + j];
        p[s][8] = M[row * LD + P];
    }
    sym_panel_eliminate<NS>(
        p, nrows, lane, cacc, fail,
        [&](int s, int kk, double2 v) {
            const int i = 32 * s + lane;  // the unscaled column entry of row J0 + i is U[kk][i]: its (upper) position
            if (i > kk && i < nrows) M[(J0 + kk) * LD + J0 + i] = v;
        },
        [&](int kk, int j) { return M[(J0 + kk) * LD + J0 + j]; });
    if (nrows <= 8) return true;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int i = 32 * s + lane;
        if (i >= 8 && i < nrows) {
            double2* dst = M + (J0 + i) * LD;
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[J0 + j] = p[s][j];
            dst[P] = p[s][8];
        }
    }
    return false;
}

// Threshold-pivoted symmetric elimination of an assembled bordered system (generic solver modes, P <= 64).
// Returns -b^T A^-1 b; fail != 0: a diagonal entry failed the test (the caller re-assembles and uses
// partial pivoting).
__device__ __forceinline__ double2 sym_schur_inplace(double2* __restrict__ M, const int P, const int LD, const int lane,
                                                     int& fail) {
    double2 cacc = make_double2(0.0, 0.0);
    JITR_PH_INIT
#pragma unroll 1
    for (int J0 = 0; J0 < P; J0 += 8) {
        const bool last = (P - J0 > 32) ? sym_panel_inplace<2>(M, P, LD, J0, lane, cacc, fail)
                                        : sym_panel_inplace<1>(M, P, LD, J0, lane, cacc, fail);
        JITR_PH(0)
        if (last) break;
        __syncwarp();
        sym_trailing_update_blocks(M, LD, J0, (P - J0 - 8) >> 3, lane);
        __syncwarp();
        JITR_PH(2)
    }
    return corner_total(cacc);
}

// Returns -b^T (That + diag d)^-1 b on every lane; fail != 0: a pivot failed the threshold test and the
// returned value is meaningless.  Arguments as fast_schur (lu_fast.cuh); W needs SymGeom<PT>::WALLOC elements.
template <int PT>
__device__ __forceinline__ double2 sym_schur(double2* __restrict__ W, double2* __restrict__ dvec,
                                             const double* __restrict__ Tp, const double* __restrict__ bsh,
                                             const double2 (&d)[SymGeom<PT>::NS0], const int lane, int& fail) {
    using G = SymGeom<PT>;
    constexpr int NS = G::NS0, n = G::n, LDW = G::LDW, NT = G::NT;
    JITR_PH_INIT
    double2 cacc = make_double2(0.0, 0.0);
    // ------------------------------------------------------------------ first panel, from That
    {
        double2 p[NS][9];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int i = 32 * s + lane;
            const bool ex = i < PT;
            const int ic = ex ? i : 0;
            if (ex) dvec[i] = d[s];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const double t = Tp[j * PT + ic];  // That is symmetric: conflict-free read across the lanes
                p[s][j] = (i == j) ? d[s] : make_double2(t, 0.0);
            }
            p[s][8] = make_double2(bsh[ic], 0.0);
        }
        sym_panel_eliminate<NS, true>(
            p, PT, lane, cacc, fail,
            [&](int s, int kk, double2 v) {
                if (s < 0) {  // lane kk: 1/d in the diagonal slot, y in the pad slot of WB row kk (both otherwise unused)
                    W[G::WB_OFF + kk * G::LB_LD + (s == -1 ? kk : 8)] = v;
                    return;
                }
                const int i = 32 * s + lane;
                if (i > kk && i < PT) W[G::WB_OFF + i * G::LB_LD + kk] = v;
            },
            [&](int kk, int j) { return j < 0 ? W[G::WB_OFF + kk * G::LB_LD + (j == -1 ? kk : 8)] : W[G::WB_OFF + j * G::LB_LD + kk]; });
        double2 ykeep[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int i = 32 * s + lane;
            ykeep[s] = p[s][8];
            if (i >= 8 && i < PT) {
                double2* dst = W + G::LB_OFF + (i - 8) * G::LB_LD;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = p[s][j];
            }
        }
        __syncwarp();
        JITR_PH(0)
        // ---- trailing update: W[r][c] = That[8+r][8+c] + d - sum_k L[r][k] w[c][k], lower tiles only
        const int fr = lane >> 2, fk = lane & 3;
        const int rho = (fr >> 1) + 4 * (fr & 1);
        const int part = fk & 1, pn = fr & 1;
        const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
        const double* Wd = reinterpret_cast<const double*>(W);
        double b[NT][2][4];
#pragma unroll
        for (int q = 0; q < NT; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const double v =
                        Wd[2 * (G::WB_OFF + (8 + 8 * q + 4 * h + (fr >> 1)) * G::LB_LD + s + 4 * (fk >> 1)) + (part ^ pn)];
                    b[q][h][s] = __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v));
                }
        __syncwarp();  // the B fragments are in registers: WB (the start of W) may be overwritten
#pragma unroll
        for (int tr = 0; tr < NT; ++tr) {
            double a[4];
#pragma unroll
            for (int s = 0; s < 4; ++s)
                a[s] = dneg(Wd[2 * (G::LB_OFF + (8 * tr + rho) * G::LB_LD + s + 4 * (fk >> 1)) + part]);
            __syncwarp();  // (the last tile row overwrites its own LB rows)
            const int row = 8 + 8 * tr + rho;
            double2 c[NT][2];
#pragma unroll
            for (int q = 0; q <= tr; ++q)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = 8 + 8 * q + 4 * h + fk;
                    c[q][h] = (row == col) ? dvec[col] : make_double2(Tp[col * PT + row], 0.0);
                }
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int q = 0; q <= tr; ++q) {
                    dmma884(c[q][0].x, c[q][0].y, a[s], b[q][0][s]);
                    dmma884(c[q][1].x, c[q][1].y, a[s], b[q][1][s]);
                }
            double2* Crow = W + (8 * tr + rho) * LDW + fk;
#pragma unroll
            for (int q = 0; q <= tr; ++q) {
                Crow[8 * q] = c[q][0];
                Crow[8 * q + 4] = c[q][1];
            }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int i = 32 * s + lane;
            if (i >= 8 && i < PT) W[(i - 8) * LDW + n] = ykeep[s];
        }
        __syncwarp();
        JITR_PH(2)
    }
    // ------------------------------------------------------------------ panels of the Schur complement in W
#pragma unroll 1
    for (int J0 = 0; J0 < n; J0 += 8) {
        const int nrows = n - J0;  // <= 32: one row per lane
        double2 p[1][9];
        {
            const int row = J0 + (lane < nrows ? lane : 0);
#pragma unroll
            for (int j = 0; j < 8; ++j) p[0][j] = W[row * LDW + J0 + j];
            p[0][8] = W[row * LDW + n];
        }
        sym_panel_eliminate<1, true>(
            p, nrows, lane, cacc, fail,
            [&](int s, int kk, double2 v) {
                if (s < 0) {  // lane kk: 1/d on the diagonal, y in the RHS column of the pivot row (neither is read again)
                    W[(J0 + kk) * LDW + (s == -1 ? J0 + kk : n)] = v;
                    return;
                }
                if (lane > kk && lane < nrows) W[(J0 + kk) * LDW + J0 + lane] = v;
            },
            [&](int kk, int j) { return j < 0 ? W[(J0 + kk) * LDW + (j == -1 ? J0 + kk : n)] : W[(J0 + kk) * LDW + J0 + j]; });
        if (nrows <= 8) {
            JITR_PH(0)
            break;
        }
        if (lane >= 8 && lane < nrows) {
            double2* dst = W + (J0 + lane) * LDW;
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[J0 + j] = p[0][j];
            dst[n] = p[0][8];
        }
        __syncwarp();
        JITR_PH(0)
        sym_trailing_update(W, LDW, J0, (nrows - 8) >> 3, lane);
        __syncwarp();
        JITR_PH(2)
    }
    return corner_total(cacc);
}

}  // namespace jitr
