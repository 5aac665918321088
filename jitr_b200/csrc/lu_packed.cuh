// Packed symmetric path: complex-symmetric LDL^T of a system whose LOWER TRIANGLE ALONE lives in shared memory.
//
// The fast paths of lu_sym.cuh keep a full n x n Schur complement per system (the unscaled pivot rows sit in the upper
// triangle), which is what limits the fused sweep to 3 resident systems per SM at N = 60 (BASELINE config 3) and is why
// the nonlocal solve (config 4) went through an L2 slab.  Here a system is stored by 8-row blocks, block row tr holding
// the columns 0 .. 8 tr + 7 of its rows and nothing right of the diagonal tile:
//
//      row r = 8 tr + i  ->  elements  off(tr) + i * ld(tr) + c,   c = 0 .. 8 tr + 7,   ld(tr) = 8 tr + 9  (odd),
//      off(tr) = 32 tr^2 + 40 tr,      the pad element  c = 8 tr + 8  of row r holds the right-hand side y_r.
//
// P = 64: 2 368 elements = 37.9 KB per system instead of 67.6 KB -- six systems per SM.  Every row stride is odd, so the
// one-row-per-lane column walks of the panel phase and the (rho-permuted) 128-bit accumulator accesses of the trailing
// update stay bank-conflict free exactly as in lu_sym.cuh.
//
// Elimination (threshold pivoting on the diagonal, as lu_sym.cuh):
//   * panel of 8 columns in registers, one row per lane (two slots while more than 32 rows are alive).  The UNSCALED
//     column entries w_ik are written back in place; the pivot row is the transposed pivot column and is read back from
//     the rows of the diagonal block as shared-memory broadcasts.  Nothing is ever stored above the diagonal.
//   * the threshold test is applied to the multipliers themselves: |Re l|, |Im l| <= 2^JITR_SYM_TAU_LOG2 for every row
//     below the pivot (equivalent to |a_kk| >= max_i |a_ik| / 64, without the REDUX of the column maximum), checked on
//     the high words (integer pipe), accumulated per lane and voted on ONCE per system.
//   * trailing update of the lower tiles on DMMA m8n8k4 in the real representation of the complex product (fragment
//     mapping of lu_warp.cuh).  A = -L is formed on the fly from the stored w and the panel's eight reciprocals
//     (one DMUL + one DFMA per fragment element), B = W^T comes straight from the stored rows: no second panel buffer.
//   * the border: corner = -sum_k y_k^2 / d_k.
// A system that fails the test is reported to the caller (the fused sweep defers its (sample, energy) pair, the batched
// solve its system, to the partial-pivoting kernels).
#pragma once
#include "lu_sym.cuh"

#ifndef JITR_PACKED_TR_UNROLL
#define JITR_PACKED_TR_UNROLL 1
#endif
#define JITR_PRAGMA_STR(x) _Pragma(#x)
#define JITR_PRAGMA_UNROLL(n) JITR_PRAGMA_STR(unroll n)
#define JITR_PACKED_TR_PRAGMA JITR_PRAGMA_UNROLL(JITR_PACKED_TR_UNROLL)

namespace jitr {

template <int P>
struct PackedGeom {
    static_assert(P % 8 == 0 && P >= 8 && P <= 64, "packed path: 8 <= P <= 64, two row slots per lane");
    static constexpr int NB = P / 8;
    __host__ __device__ static constexpr int ld(int tr) { return 8 * tr + 9; }
    __host__ __device__ static constexpr int off(int tr) { return 32 * tr * tr + 40 * tr; }
    __host__ __device__ static constexpr int row(int r) { return off(r >> 3) + (r & 7) * ld(r >> 3); }
    __host__ __device__ static constexpr int rhs(int r) { return row(r) + 8 * (r >> 3) + 8; }  // pad element = y_r
    static constexpr int SIZE = off(NB);       // complex elements of the packed matrix (incl. the y slots)
    static constexpr int WARP_ELEMS = SIZE;    // (the reciprocal of pivot k takes the place of the diagonal entry k once it is used)
};

// high words of |re|, |im| beyond 2^TAU (or NaN / Inf): the multiplier fails the threshold test
__device__ __forceinline__ bool mult_too_large(const double2 l) {
    constexpr int kLim = 0x3ff00000 + (JITR_SYM_TAU_LOG2 << 20);  // high word of 2^TAU
    const int hx = __double2hiint(l.x) & 0x7fffffff, hy = __double2hiint(l.y) & 0x7fffffff;
    // |v| <= 2^TAU  <=>  hi < kLim, or hi == kLim with a zero mantissa; the cheap form hi >= kLim + 1 lets (2^TAU, 2^TAU(1+2^-20)) pass
    return max(hx, hy) > kLim;
}

// One panel (columns 8 kb .. 8 kb + 7), rows i = 32 s + lane relative to the panel start, nrows alive.
template <int P, int NS>
__device__ __forceinline__ void packed_panel(double2* __restrict__ M, const int kb, const int lane, double2& cacc, int& bad) {
    using G = PackedGeom<P>;
    const int J0 = 8 * kb, nrows = P - J0;
    const int ldk = G::ld(kb);
    const int dbase = G::off(kb) + J0;  // entry (J0 + j, J0 + k) of the diagonal block: dbase + j * ldk + k
    int rb[NS], ry[NS];
    double2 p[NS][9];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int i = 32 * s + lane;
        const int ic = i < nrows ? i : 0;
        const int tr = kb + (ic >> 3);
        rb[s] = G::off(tr) + (ic & 7) * G::ld(tr) + J0;
        ry[s] = rb[s] - J0 + 8 * tr + 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) p[s][j] = M[rb[s] + j];
        p[s][8] = M[ry[s]];
    }
    double2 ysave = make_double2(0.0, 0.0), rsave = make_double2(0.0, 0.0);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        // The pivot chain is the critical path of a warp-per-system elimination (64 columns in sequence, every link a
        // dependent FP64 instruction of ~23 cycles), so the multiplier is formed as  l = (w conj(d)) / |d|^2 : the product
        // w conj(d) needs only the pivot itself (broadcast at once) and runs beside the reciprocal of |d|^2, which every lane
        // computes speculatively on its own entry (only lane kk's is used) -- two links shorter than  l = w (1/d).
        const double2 dmine = p[0][kk];
        const double mys = rcp_halley(fma(dmine.x, dmine.x, dmine.y * dmine.y));
        if (lane == kk) {
            const int m = max(__double2hiint(dmine.x) & 0x7fffffff, __double2hiint(dmine.y) & 0x7fffffff);
            if (m < 0x00100000 || m >= 0x7ff00000) bad = 1;  // zero, denormal, Inf or NaN pivot
            ysave = p[0][8];
            rsave = make_double2(dmine.x * mys, -dmine.y * mys);
        }
        // the unscaled column entries, in place (rows below the pivot): W of the trailing update, and the pivot row
#pragma unroll
        // the pivot and the pivot row's y go through shared memory too (their own positions: the diagonal entry is replaced by
        // 1/d after the panel, the right-hand side slot gets the same value again): 3 LSU wavefronts instead of 8 for two shuffles
        for (int s = 0; s < NS; ++s) {
            const int i = 32 * s + lane;
            if (i >= kk + (s ? 1 : 0) && i < nrows) M[rb[s] + kk] = p[s][kk];
        }
        if (lane == kk) M[ry[0]] = p[0][8];
        __syncwarp();
        const double2 piv = M[dbase + kk * ldk + kk];
        double2 u[9];
        u[8] = M[dbase - J0 + kk * ldk + 8 * kb + 8];
#pragma unroll
        for (int j = kk + 1; j < 8; ++j) u[j] = M[dbase + j * ldk + kk];  // pivot row = transposed pivot column
        const double sc = __shfl_sync(kFull, mys, kk);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int i = 32 * s + lane;
            const double2 w = p[s][kk];
            const double tx = fma(w.x, piv.x, w.y * piv.y), ty = fma(w.y, piv.x, -(w.x * piv.y));  // w conj(d)
            const double2 l = make_double2(tx * sc, ty * sc);
            if (i > kk && i < nrows && mult_too_large(l)) bad = 1;
#pragma unroll
            for (int j = kk + 1; j < 9; ++j) cfnma(p[s][j], l, u[j]);
        }
    }
    // corner -= sum_k y_k^2 / d_k : one term per lane k < 8, per lane across the panels (packed_ldlt sums the lanes once)
    const double2 term = cmul(cmul(ysave, rsave), ysave);
    cacc.x -= term.x;
    cacc.y -= term.y;
    // 1/d_k replaces the diagonal entry (trailing update, back substitution); every row keeps its final right-hand side y
    if (lane < 8) {
        M[rb[0] + lane] = rsave;
        M[ry[0]] = ysave;
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int i = 32 * s + lane;
        if (i >= 8 && i < nrows) M[ry[s]] = p[s][8];
    }
}

// lower tiles (tr, q), kb < q <= tr < NB, -= L[tr] W[q]^T  after panel kb; tile columns in blocks of three
template <int P>
__device__ __forceinline__ void packed_trailing(double2* __restrict__ M, const int kb, const int lane) {
    using G = PackedGeom<P>;
    constexpr int NB = G::NB;
    const int J0 = 8 * kb;
    const int fr = lane >> 2, fk = lane & 3;
    const int rho = (fr >> 1) + 4 * (fr & 1);
    const int part = fk & 1, pn = fr & 1;
    const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
    const int kcol = J0 + 4 * (fk >> 1);
    const double* Md = reinterpret_cast<const double*>(M);
    // a[s] = -(component `part` of w * rinv) = -(w.x c1 + w.y c2)
    double c1[4], c2[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int jd = 4 * (fk >> 1) + s;
        const double2 r = M[G::off(kb) + jd * G::ld(kb) + J0 + jd];  // 1/d of column J0 + jd
        c1[s] = part ? r.y : r.x;
        c2[s] = part ? r.x : -r.y;
    }
#pragma unroll 1
    for (int q0 = kb + 1; q0 < NB; q0 += 3) {
        double b[3][2][4];
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (q0 + q < NB) {
                const int qb = G::off(q0 + q) + (fr >> 1) * G::ld(q0 + q) + kcol;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const double v = Md[2 * (qb + 4 * h * G::ld(q0 + q) + s) + (part ^ pn)];
                        b[q][h][s] = __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v));
                    }
            }
        JITR_PACKED_TR_PRAGMA
        for (int tr = q0; tr < NB; ++tr) {
            const int rbase = G::off(tr) + rho * G::ld(tr);
            double a[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double2 w = M[rbase + kcol + s];
                a[s] = fma(-w.x, c1[s], -(w.y * c2[s]));
            }
            double2* Crow = M + rbase + 8 * q0 + fk;
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (q0 + q <= tr) {  // lower triangle (diagonal tiles in full)
                    double2 c0 = Crow[8 * q], c1v = Crow[8 * q + 4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        dmma884(c0.x, c0.y, a[s], b[q][0][s]);
                        dmma884(c1v.x, c1v.y, a[s], b[q][1][s]);
                    }
                    Crow[8 * q] = c0;
                    Crow[8 * q + 4] = c1v;
                }
        }
    }
}

// In-place elimination of a packed system (lower triangle + y slots).  Returns -y^T D^-1 y = -b^T A^-1 b on every lane;
// bad != 0 (warp-uniform on return): a pivot failed the threshold test, the value is meaningless.
template <int P>
__device__ __forceinline__ double2 packed_ldlt(double2* __restrict__ M, const int lane, int& bad,
                                               const double2 cacc0 = make_double2(0.0, 0.0)) {
    // cacc0: PER-LANE corner terms of panels eliminated before the call (packed_first_panel); summed with this function's own
    using G = PackedGeom<P>;
    double2 cacc = cacc0;
    int mybad = 0;
    JITR_PH_INIT
#pragma unroll 1
    for (int kb = 0; kb < G::NB; ++kb) {
        if (P - 8 * kb > 32) packed_panel<P, 2>(M, kb, lane, cacc, mybad);
        else packed_panel<P, 1>(M, kb, lane, cacc, mybad);
        JITR_PH(0)
        if (kb == G::NB - 1) break;
        __syncwarp();
        packed_trailing<P>(M, kb, lane);
        __syncwarp();
        JITR_PH(2)
    }
    bad = __any_sync(kFull, mybad) ? 1 : 0;
    return corner_total(cacc);
}

// First panel of  That + diag(d)  straight from That (the fused sweep: the system never exists as a whole, lu_sym.cuh), for the
// packed layout: the eight pivot columns are eliminated in registers from That's columns (global memory, L1/L2 resident), and only
// the Schur complement of order n = PT - 8 is written -- into a PackedGeom<PT - 8> region (P = 64: 29.6 KB instead of 37.9 KB:
// SEVEN systems per SM, and no assembly pass).  Scratch inside the region while the panel is in flight:
//   WB  the unscaled columns of all PT rows (pivot-row source and, rows 8.., the B operands), at the START of the region: all B
//       fragments are taken into registers before the first block row is written over it;
//   LB  the multipliers of rows 8..PT-1, at the END: block rows are written in ascending order and reach LB only at the last
//       two, whose own A fragments are loaded before they are stored.
// cacc: the panel's PER-LANE corner terms are accumulated into it (hand it to packed_ldlt as cacc0); `fail` as sym_panel_eliminate.  d[s]: diagonal entries of rows lane + 32 s
// (the caller passes 1 for the identity padding), bsh: boundary vector b (shared memory, zero padded), N: real rows.
template <int PT>
__device__ __forceinline__ void packed_first_panel(double2* __restrict__ W, const double* __restrict__ That, const double* __restrict__ bsh,
                                                   const double2 (&d)[2], const int N, const int lane, double2& cacc, int& fail) {
    using G = PackedGeom<PT - 8>;
    constexpr int n = PT - 8, NT = n / 8, LB_LD = 9;
    constexpr int WB_OFF = 0, LB_OFF = G::SIZE - n * LB_LD;
    static_assert(PT > 32 && PT <= 64, "two row slots");
    static_assert(LB_OFF >= PT * LB_LD, "scratch of the first panel must fit the region");
    double2 p[2][9];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int i = 32 * s + lane;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double t = (i < N && j < N && i != j) ? __ldg(That + j * N + i) : 0.0;  // That is symmetric: coalesced across the lanes
            p[s][j] = (i == j) ? d[s] : make_double2(t, 0.0);
        }
        p[s][8] = make_double2(i < PT ? bsh[i] : 0.0, 0.0);
    }
    sym_panel_eliminate<2, true>(
        p, PT, lane, cacc, fail,
        [&](int s, int kk, double2 v) {
            if (s < 0) {
                W[WB_OFF + kk * LB_LD + (s == -1 ? kk : 8)] = v;
                return;
            }
            const int i = 32 * s + lane;
            if (i > kk && i < PT) W[WB_OFF + i * LB_LD + kk] = v;
        },
        [&](int kk, int j) { return j < 0 ? W[WB_OFF + kk * LB_LD + (j == -1 ? kk : 8)] : W[WB_OFF + j * LB_LD + kk]; });
    double2 ykeep[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int i = 32 * s + lane;
        ykeep[s] = p[s][8];
        if (i >= 8 && i < PT) {
            double2* dst = W + LB_OFF + (i - 8) * LB_LD;
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = p[s][j];
        }
    }
    __syncwarp();
    // ---- Schur complement  W[r][c] = That[8+r][8+c] (+ d on the diagonal) - sum_k L[r][k] w[c][k], lower tiles only
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
                const double v = Wd[2 * (WB_OFF + (8 + 8 * q + 4 * h + (fr >> 1)) * LB_LD + s + 4 * (fk >> 1)) + (part ^ pn)];
                b[q][h][s] = __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v));
            }
    __syncwarp();  // the B fragments are in registers: WB may be overwritten
#pragma unroll
    for (int tr = 0; tr < NT; ++tr) {
        const int row = 8 + 8 * tr + rho;  // row of the full system
        // the entries of That (and the diagonal) this lane's accumulators start from: issued first, they arrive during the A loads
        double2 c[NT][2];
        const double2 dv = shfl2(row >= 32 ? d[1] : d[0], row & 31);  // d of this lane's accumulator row (used where row == col)
#pragma unroll
        for (int q = 0; q <= tr; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = 8 + 8 * q + 4 * h + fk;
                const double t = (row != col && row < N && col < N) ? __ldg(That + col * N + row) : 0.0;
                c[q][h] = (row == col) ? dv : make_double2(t, 0.0);
            }
        double a[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) a[s] = dneg(Wd[2 * (LB_OFF + (8 * tr + rho) * LB_LD + s + 4 * (fk >> 1)) + part]);
        __syncwarp();  // (the last block rows overwrite LB rows, their own included)
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int q = 0; q <= tr; ++q) {
                dmma884(c[q][0].x, c[q][0].y, a[s], b[q][0][s]);
                dmma884(c[q][1].x, c[q][1].y, a[s], b[q][1][s]);
            }
        double2* Crow = W + G::off(tr) + rho * G::ld(tr) + fk;
#pragma unroll
        for (int q = 0; q <= tr; ++q) {
            Crow[8 * q] = c[q][0];
            Crow[8 * q + 4] = c[q][1];
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int i = 32 * s + lane;
        if (i >= 8 && i < PT) W[G::rhs(i - 8)] = ykeep[s];
    }
    __syncwarp();
}

// z = A^-1 b from the factors packed_ldlt leaves behind (W below the diagonal, 1/d on it, y = L^-1 b in the y slots):
// z_i = (y_i - sum_{j > i} w_ji z_j) / d_i, column oriented -- row j of the packed storage holds w_ji for all i < j, lane i
// (+32) accumulates its own entry, z_j is broadcast as it becomes final.  z[h] = entry lane + 32 h.  (core.py:63-82)
template <int P>
__device__ __forceinline__ void packed_back_substitute(const double2* __restrict__ M, const int lane, double2 (&z)[2]) {
    using G = PackedGeom<P>;
    double2 t[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        t[h] = i < P ? M[G::rhs(i)] : make_double2(0.0, 0.0);
        z[h] = make_double2(0.0, 0.0);
    }
#pragma unroll 1
    for (int j = P - 1; j >= 0; --j) {
        const int rowj = G::row(j);
        const double2 tj = shfl2(j >= 32 ? t[1] : t[0], j & 31);
        const double2 zj = cmul(tj, M[rowj + j]);
        if (lane == (j & 31)) {
            if (j >= 32) z[1] = zj;
            else z[0] = zj;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;
            if (i < j) cfnma(t[h], M[rowj + i], zj);
        }
    }
}

}  // namespace jitr
