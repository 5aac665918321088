// Fast path of the bordered LU for nbasis <= 40: the first panel is factorised straight from the
// shared kinetic matrix, so a system never exists in shared memory as a whole.
//
// Every system of a sweep is  a^2 A = That + diag(d)  with the SAME real matrix That (kept once per CTA
// in shared memory, identity-padded to PT = 8 ceil(N/8)) and a per-system complex diagonal d.  The
// first eight columns are therefore read from That into the panel registers, factorised there, and
// only the Schur complement of that panel -- an (PT-8) x (PT-8) complex matrix plus the border -- is
// ever written to the warp's shared-memory region W:
//     storage per system  (PT-7)^2 x 16 B   (17.4 KB for PT = 40, against 26.9 KB for the whole matrix)
// which lifts the number of resident systems (= warps) per SM from 8 to 12 and removes the assembly pass.
// The multipliers L21 and the row block U12 of the first panel pass through W as scratch on their way
// into DMMA fragments held in registers; then the trailing tiles  That[rows, cols] + d - L21 U12  are
// formed in the accumulators and stored to W, and the generic panels of lu_warp.cuh take over on W.
#pragma once
#include "lu_warp.cuh"

namespace jitr {

template <int PT>
struct FastGeom {
    static constexpr int n = PT - 8;        // order of the Schur complement kept in W
    static constexpr int LDW = n + 1;       // odd leading dimension; column n = RHS, row n = z row
    static constexpr int NT = n / 8;        // tiles per dimension of the first trailing update
    static constexpr int NS0 = (PT + 1 > 32) ? 2 : 1;
    static constexpr int WSZ = LDW * LDW;   // complex elements of W
    // scratch inside W while the first panel is in flight
    static constexpr int LB_LD = 9;                  // L21 rows (position - 8), 8 multipliers each, odd stride
    static constexpr int LB_OFF = 0;                 // (n + 1) rows: the last one is the z row
    static constexpr int L11_OFF = LB_OFF + (n + 1) * LB_LD;  // 8 x 8 multipliers of the pivot rows
    static constexpr int UB_LD = n + 1;              // U12 rows k = 0..7, n columns, odd stride
    static constexpr int UB_OFF = L11_OFF + 64;
    static constexpr int SCRATCH = UB_OFF + 8 * UB_LD;
    static constexpr int WALLOC = WSZ > SCRATCH ? WSZ : SCRATCH;  // complex elements reserved for W
    static_assert(n >= 8 && n <= 32, "fast path covers 16 <= PT <= 40");
    // per-warp bytes: W, the diagonal d (PT complex), the original row index of every position
    static constexpr int ORIG_BYTES = ((PT + 1 + 15) / 16) * 16;
    static constexpr int WARP_BYTES = WALLOC * 16 + PT * 16 + ORIG_BYTES;
};

// Tp   [PT][PT] real, shared by the CTA: That padded with the identity
// bsh  [PT] real: boundary vector, 0 in the padding
// d    this lane's diagonal entries d[s] of rows 32 s + lane ((1, 0) in the padding)
// W, dvec, origv: this warp's shared-memory regions (FastGeom)
// Returns -b^T (That + diag d)^-1 b on every lane.
template <int PT>
__device__ __forceinline__ double2 fast_schur(double2* __restrict__ W, double2* __restrict__ dvec,
                                              unsigned char* __restrict__ origv, const double* __restrict__ Tp,
                                              const double* __restrict__ bsh, const double2 (&d)[FastGeom<PT>::NS0],
                                              const int lane, int& singular) {
    using G = FastGeom<PT>;
    constexpr int NS = G::NS0, n = G::n, LDW = G::LDW, NT = G::NT;
    JITR_PH_INIT
    LuGeom g;
    g.N = n;
    g.P = n;
    g.LD = LDW;
    // panels that need NS row slots go through ONE instance of the unrolled elimination loop: the first
    // panel (from That) and, when the Schur complement still has more than 32 rows, its first panel
    constexpr int NA = 1 + ((NS == 2 && n + 1 > 32) ? 1 : 0);
    double2 p[NS][9];
    int pos[NS];
    bool cand[NS], exists[NS];
    double2 zc = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int it = 0; it < NA; ++it) {
    if (it == 0) {
        // -------------------------------------------------------------- first panel, phase P: load
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int i = 32 * s + lane;
            cand[s] = i < PT;
            exists[s] = i <= PT;  // the z row is row PT
            pos[s] = cand[s] ? i : PT;
            const int ic = cand[s] ? i : 0;
            if (cand[s]) dvec[i] = d[s];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // That is symmetric: row j, column i is a conflict-free read across the lanes
                const double t = cand[s] ? Tp[j * PT + ic] : bsh[j];
                p[s][j] = (i == j) ? d[s] : make_double2(t, 0.0);
            }
            p[s][8] = make_double2(cand[s] ? bsh[ic] : 0.0, 0.0);
        }
    } else {
        panel_load<NS>(p, pos, cand, exists, W, g, 0, lane);
    }
    panel_eliminate<NS>(p, pos, cand, W, LDW, 0, 8, n - 8, it > 0, lane, singular);
    if (it > 0) {
        panel_store<NS>(p, pos, exists, W, g, 0);
        JITR_PH(0)
        __syncwarp();
        upper_phase(W, g, 0, lane);
        __syncwarp();
        JITR_PH(1)
        trailing_update_dmma(W, g, 0, lane);
        __syncwarp();
        JITR_PH(2)
        continue;
    }
    // multipliers -> scratch, original row of every position
    double2 ykeep[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        ykeep[s] = p[s][8];
        if (exists[s]) {
            origv[pos[s]] = (unsigned char)(32 * s + lane);
            double2* dst = pos[s] >= 8 ? W + G::LB_OFF + (pos[s] - 8) * G::LB_LD : W + G::L11_OFF + pos[s] * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = p[s][j];
        }
    }
    __syncwarp();
    JITR_PH(0)
    // ------------------------------------------------------------------ first panel, phase U
    // lane c owns trailing column 8 + c: rows = the eight pivot rows of That (+ d on the diagonal)
    {
        const int c = lane < n ? lane : 0;
        const int col = 8 + c;
        double2 a[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int o = origv[kk];  // warp-uniform
            a[kk] = (o == col) ? dvec[col] : make_double2(Tp[o * PT + col], 0.0);
        }
#pragma unroll
        for (int kk = 1; kk < 8; ++kk)
#pragma unroll
            for (int k2 = 0; k2 < kk; ++k2) cfnma(a[kk], W[G::L11_OFF + kk * 8 + k2], a[k2]);
        zc = make_double2(bsh[col], 0.0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) cfnma(zc, W[G::LB_OFF + n * G::LB_LD + kk], a[kk]);
        if (lane < n) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) W[G::UB_OFF + kk * G::UB_LD + c] = a[kk];
        }
    }
    __syncwarp();
    JITR_PH(1)
    // ------------------------------------------------------------------ first panel, phase T
    // (fragment <-> matrix mapping: see trailing_update_dmma in lu_warp.cuh)
    {
        const int fr = lane >> 2, fk = lane & 3;
        const int rho = (fr >> 1) + 4 * (fr & 1);
        const int part = fk & 1, pn = fr & 1;
        const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
        const double* Wd = reinterpret_cast<const double*>(W);
        double b[NT][2][4], a[NT][4];
        int orow[NT];
#pragma unroll
        for (int q = 0; q < NT; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const double v =
                        Wd[2 * (G::UB_OFF + (s + 4 * (fk >> 1)) * G::UB_LD + 8 * q + 4 * h + (fr >> 1)) + (part ^ pn)];
                    b[q][h][s] = __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v));
                }
#pragma unroll
        for (int tr = 0; tr < NT; ++tr) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
                a[tr][s] = dneg(Wd[2 * (G::LB_OFF + (8 * tr + rho) * G::LB_LD + s + 4 * (fk >> 1)) + part]);
            orow[tr] = origv[8 + 8 * tr + rho];
        }
        __syncwarp();  // every fragment is in registers: the scratch may be overwritten
#pragma unroll
        for (int tr = 0; tr < NT; ++tr) {
            double2 c[NT][2];
#pragma unroll
            for (int q = 0; q < NT; ++q)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = 8 + 8 * q + 4 * h + fk;
                    c[q][h] = (orow[tr] == col) ? dvec[col] : make_double2(Tp[col * PT + orow[tr]], 0.0);
                }
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int q = 0; q < NT; ++q) {
                    dmma884(c[q][0].x, c[q][0].y, a[tr][s], b[q][0][s]);
                    dmma884(c[q][1].x, c[q][1].y, a[tr][s], b[q][1][s]);
                }
            double2* Crow = W + (8 * tr + rho) * LDW + fk;
#pragma unroll
            for (int q = 0; q < NT; ++q) {
                Crow[8 * q] = c[q][0];
                Crow[8 * q + 4] = c[q][1];
            }
        }
        // border: RHS column (rows that were not pivots; the z row gives the corner) and the z row
#pragma unroll
        for (int s = 0; s < NS; ++s)
            if (exists[s] && pos[s] >= 8) W[(pos[s] - 8) * LDW + n] = ykeep[s];
        if (lane < n) W[n * LDW + lane] = zc;
    }
    __syncwarp();
    JITR_PH(2)
    }  // it
    // ------------------------------------------------------------------ remaining panels on W
    return bordered_schur<1>(W, g, lane, singular, 8 * (NA - 1));
}

}  // namespace jitr
