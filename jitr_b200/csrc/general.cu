// General batched R-matrix solve: any number of channels, local and/or nonlocal interaction,
// optional wavefunction coefficients.  Replaces Solver.solve (rmatrix/rmatrix.py:180-246) +
// rmatrix/core.py:7-82 for one system per call in the reference.  Three kernels, chosen by size:
//   warp_solve_kernel     single channel, no coefficients, nbasis <= 48: one warp per system (lu_warp.cuh)
//   block_solve_kernel    up to 256 bordered rows (config 4: N = 60 nonlocal; config 5: 3 x 50 coupled;
//                         elastic workspaces with nbasis > 64): one CTA per system, blocked DMMA LU on an
//                         L2-resident slab
//   general_solve_kernel  anything larger: unblocked elimination, one CTA per system
// The fused warp-per-system kernel in elastic.cu is the fast path of the single-channel elastic sweep.
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"
#include "jitr_b200.h"
#include "general_decl.cuh"
#include "lu_warp.cuh"

namespace jitr {

constexpr int kGenThreads = 256;

// R (in small[0 .. nch^2)) -> Z+-, S = Z+^-1 Z-, u'_ext; writes R, S, uext, status of system sysi.  Called by the
// whole CTA after a barrier that made R visible; ends with a barrier (ue = small + 3 nch^2 is then readable).
__device__ __forceinline__ void finish_system(const GenArgs& A, const long long sysi, const double a, const int bad_flags,
                                              double2* small, const int tid, const int nt) {
    const int nch = A.nch;
    double2* Rm = small;                          // [nch][nch]
    double2* Zp = small + nch * nch;              // [nch][nch]
    double2* Zm = small + 2 * nch * nch;          // [nch][nch] -> S
    double2* ue = small + 3 * nch * nch;          // [nch]
    const double2* as = A.asym + (size_t)sysi * A.asym_stride;  // [4][nch]: Hp, Hm, Hpp, Hmp
    for (int ij = tid; ij < nch * nch; ij += nt) {
        const int i = ij / nch, j = ij - i * nch;
        // Z+- = diag(H+-) - a R diag(H+-')   (core.py:51-53)
        const double2 r = Rm[ij];
        const double2 tp = cmul(r, as[2 * nch + j]), tm = cmul(r, as[3 * nch + j]);
        double2 zp = make_double2(-a * tp.x, -a * tp.y), zm = make_double2(-a * tm.x, -a * tm.y);
        if (i == j) {
            zp.x += as[j].x; zp.y += as[j].y;
            zm.x += as[nch + j].x; zm.y += as[nch + j].y;
        }
        Zp[ij] = zp;
        Zm[ij] = zm;
    }
    __syncthreads();
    if (tid == 0) {
        // S = Z+^-1 Z-  by Gaussian elimination with partial pivoting (nch <= kMaxCh)
        for (int k = 0; k < nch; ++k) {
            int p = k;
            double best = -1.0;
            for (int i = k; i < nch; ++i) {
                const double2 v = Zp[i * nch + k];
                const double m = v.x * v.x + v.y * v.y;
                if (m > best) { best = m; p = i; }
            }
            if (p != k)
                for (int c = 0; c < nch; ++c) {
                    double2 t = Zp[k * nch + c]; Zp[k * nch + c] = Zp[p * nch + c]; Zp[p * nch + c] = t;
                    t = Zm[k * nch + c]; Zm[k * nch + c] = Zm[p * nch + c]; Zm[p * nch + c] = t;
                }
            const double2 rinv = crecip(Zp[k * nch + k]);
            for (int i = k + 1; i < nch; ++i) {
                const double2 l = cmul(Zp[i * nch + k], rinv);
                for (int c = k + 1; c < nch; ++c) cfnma(Zp[i * nch + c], l, Zp[k * nch + c]);
                for (int c = 0; c < nch; ++c) cfnma(Zm[i * nch + c], l, Zm[k * nch + c]);
            }
        }
        for (int c = 0; c < nch; ++c)
            for (int i = nch - 1; i >= 0; --i) {
                double2 v = Zm[i * nch + c];
                for (int m = i + 1; m < nch; ++m) cfnma(v, Zp[i * nch + m], Zm[m * nch + c]);
                Zm[i * nch + c] = cmul(v, crecip(Zp[i * nch + i]));
            }
        // u'_ext = (i/2)(H-' w - S H+')   (core.py:58)
        int bad = bad_flags;
        for (int i = 0; i < nch; ++i) {
            double2 v = make_double2(as[3 * nch + i].x * A.weights[i], as[3 * nch + i].y * A.weights[i]);
            for (int j = 0; j < nch; ++j) cfnma(v, Zm[i * nch + j], as[2 * nch + j]);
            ue[i] = make_double2(-0.5 * v.y, 0.5 * v.x);
            if (!(isfinite(v.x) && isfinite(v.y))) bad |= 1;
        }
        A.status[sysi] = (bad & 1) ? JITR_B200_STATUS_NONFINITE : ((bad & 2) ? JITR_B200_STATUS_SINGULAR : JITR_B200_STATUS_OK);
    }
    __syncthreads();
    for (int ij = tid; ij < nch * nch; ij += nt) {
        A.R[(size_t)sysi * nch * nch + ij] = Rm[ij];
        A.S[(size_t)sysi * nch * nch + ij] = Zm[ij];
    }
    for (int i = tid; i < nch; i += nt) A.uext[(size_t)sysi * nch + i] = ue[i];
}

template <bool IN_SMEM>
__global__ void __launch_bounds__(kGenThreads) general_solve_kernel(const GenArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = A.n, N = A.N, nch = A.nch, ld = n + nch;
    const int tid = threadIdx.x, nt = blockDim.x;
    // shared scratch: pivot column multipliers, pivot row, reduction slots, small matrices
    double2* lcol = reinterpret_cast<double2*>(smem_raw);  // [n]
    double2* urow = lcol + n;                               // [ld]
    double* redv = reinterpret_cast<double*>(urow + ld);    // [32]
    int* redi = reinterpret_cast<int*>(redv + 32);          // [36]: 32 slots + pivot row + flags
    double2* small = reinterpret_cast<double2*>(redi + 36); // [4 * kMaxCh * kMaxCh]
    double2* Msm = small + 4 * kMaxCh * kMaxCh;

    for (long long sysi = blockIdx.x; sysi < A.batch; sysi += gridDim.x) {
        double2* M = IN_SMEM ? Msm : A.workspace + (size_t)sysi * n * ld;
        const double* sy = A.sys + (size_t)sysi * A.sys_stride;
        const double a = sy[0], E0 = sy[1], k0 = sy[2];
        const double2* F = A.free_matrix + (A.free_index ? (size_t)A.free_index[sysi] * A.n * A.n : (size_t)sysi * A.free_stride);
        // ---------------- assemble [A | B]
        const double nl_scale = (a / k0) / k0 / E0;  // kernel.py:202 (x radius), rmatrix.py:175-178 (/k0, /E0)
        for (int idx = tid; idx < n * ld; idx += nt) {
            const int r = idx / ld, c = idx - r * ld;
            double2 v;
            if (c < n) {
                v = F[(size_t)r * n + c];
                const int ci = r / N, nn = r - ci * N, cj = c / N, mm = c - cj * N;
                if (A.interaction) {
                    const double2 w = A.interaction[(size_t)sysi * n * n + (size_t)r * n + c];
                    v.x += w.x; v.y += w.y;
                } else {
                    if (A.local && nn == mm) {
                        const double2 w = A.local[(((size_t)sysi * nch + ci) * nch + cj) * N + nn];
                        v.x += w.x / E0; v.y += w.y / E0;
                    }
                    if (A.nonlocal_) {
                        const double2 w = A.nonlocal_[((((size_t)sysi * nch + ci) * nch + cj) * N + nn) * N + mm];
                        const double s = A.wsqrt[nn] * A.wsqrt[mm] * nl_scale;
                        v.x += w.x * s; v.y += w.y * s;
                    }
                }
            } else {
                const int j = c - n, ci = r / N;
                v = make_double2(ci == j ? A.bvec[r - ci * N] : 0.0, 0.0);
            }
            M[idx] = v;
        }
        if (tid == 0) redi[33] = 0;
        __syncthreads();
        // ---------------- Gaussian elimination with partial pivoting on the augmented matrix
        for (int k = 0; k < n; ++k) {
            // pivot search over rows k..n-1 of column k
            double best = -1.0;
            int bi = k;
            for (int i = k + tid; i < n; i += nt) {
                const double2 v = M[(size_t)i * ld + k];
                const double m = v.x * v.x + v.y * v.y;
                if (m > best) { best = m; bi = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_down_sync(kFull, best, o);
                const int oi = __shfl_down_sync(kFull, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if ((tid & 31) == 0) { redv[tid >> 5] = best; redi[tid >> 5] = bi; }
            __syncthreads();
            if (tid < 32) {
                best = (tid < (nt >> 5)) ? redv[tid] : -1.0;
                bi = (tid < (nt >> 5)) ? redi[tid] : k;
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_down_sync(kFull, best, o);
                    const int oi = __shfl_down_sync(kFull, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (tid == 0) {
                    redi[32] = bi;
                    if (!(best > 0.0)) redi[33] |= (best == 0.0) ? 2 : 1;  // zero pivot / NaN column
                }
            }
            __syncthreads();
            const int p = redi[32];
            // swap rows k and p (columns k..ld-1) while staging the pivot row
            for (int c = k + tid; c < ld; c += nt) {
                const double2 vp = M[(size_t)p * ld + c];
                if (p != k) {
                    M[(size_t)p * ld + c] = M[(size_t)k * ld + c];
                    M[(size_t)k * ld + c] = vp;
                }
                urow[c] = vp;
            }
            __syncthreads();
            const double2 rinv = crecip(urow[k]);
            for (int i = k + 1 + tid; i < n; i += nt) lcol[i] = cmul(M[(size_t)i * ld + k], rinv);
            __syncthreads();
            // rank-1 update of the trailing block (RHS columns included)
            const int w = ld - (k + 1), h = n - (k + 1);
            for (int idx = tid; idx < w * h; idx += nt) {
                const int i = k + 1 + idx / w, c = k + 1 + idx % w;
                double2 v = M[(size_t)i * ld + c];
                cfnma(v, lcol[i], urow[c]);
                M[(size_t)i * ld + c] = v;
            }
            __syncthreads();
        }
        // ---------------- back substitution for the nch right-hand sides (one warp per RHS column)
        {
            const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
            for (int j = warp; j < nch; j += nw) {
                for (int i = n - 1; i >= 0; --i) {
                    double2 acc = make_double2(0.0, 0.0);
                    for (int c = i + 1 + lane; c < n; c += 32) {
                        const double2 u = M[(size_t)i * ld + c], x = M[(size_t)c * ld + n + j];
                        acc.x += u.x * x.x - u.y * x.y;
                        acc.y += u.x * x.y + u.y * x.x;
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        acc.x += __shfl_down_sync(kFull, acc.x, o);
                        acc.y += __shfl_down_sync(kFull, acc.y, o);
                    }
                    if (lane == 0) {
                        double2 y = M[(size_t)i * ld + n + j];
                        y.x -= acc.x; y.y -= acc.y;
                        M[(size_t)i * ld + n + j] = cmul(y, crecip(M[(size_t)i * ld + i]));
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        // ---------------- R_ij = b^T X_(i-block, j) / a^2     (core.py:15-22)
        double2* Rm = small;                          // [nch][nch]
        for (int ij = tid; ij < nch * nch; ij += nt) {
            const int i = ij / nch, j = ij - i * nch;
            double2 acc = make_double2(0.0, 0.0);
            for (int m = 0; m < N; ++m) {
                const double2 x = M[(size_t)(i * N + m) * ld + n + j];
                const double b = A.bvec[m];
                acc.x += b * x.x; acc.y += b * x.y;
            }
            Rm[ij] = make_double2(acc.x / (a * a), acc.y / (a * a));
        }
        __syncthreads();
        finish_system(A, sysi, a, redi[33], small, tid, nt);
        const double2* ue = small + 3 * nch * nch;
        if (A.coeffs) {
            // x = A^-1 vec_i(b u'_ext,i) = sum_i u'_ext,i X[:, i]    (core.py:81-82)
            for (int r = tid; r < n; r += nt) {
                double2 acc = make_double2(0.0, 0.0);
                for (int i = 0; i < nch; ++i) {
                    const double2 t = cmul(ue[i], M[(size_t)r * ld + n + i]);
                    acc.x += t.x; acc.y += t.y;
                }
                A.coeffs[(size_t)sysi * n + r] = acc;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Single-channel systems (nch = 1, no coefficients) up to P = 8 ceil(N/8) <= JITR_WARP_SOLVE_MAXP: one WARP per system, the
// bordered matrix [[F + V, b], [b^T, 0]] assembled in the warp's shared-memory region straight from the
// global arrays (the nonlocal kernel is the HBM stream of this path: N^2 x 16 B per system) and eliminated
// by the blocked partial-pivoting LU of lu_warp.cuh (register panel + DMMA trailing update).
// ---------------------------------------------------------------------------------------------
template <int MAXNS>
__global__ void __launch_bounds__(256, 1) warp_solve_kernel(const GenArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = A.N;
    LuGeom g;
    g.N = N;
    g.P = (N + 7) & ~7;
    g.LD = g.P + 1;
    const int P = g.P, LD = g.LD, msz = LD * LD;
    const int nwarps = blockDim.x >> 5;
    const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    double2* M = reinterpret_cast<double2*>(smem_raw) + (size_t)warp * msz;
    const double2 zero2 = make_double2(0.0, 0.0);
    for (long long sysi = (long long)blockIdx.x * nwarps + warp; sysi < A.batch; sysi += (long long)gridDim.x * nwarps) {
        const double* sy = A.sys + (size_t)sysi * A.sys_stride;
        const double a = sy[0], E0 = sy[1], k0 = sy[2];
        const double2* F = A.free_matrix + (A.free_index ? (size_t)A.free_index[sysi] * A.n * A.n : (size_t)sysi * A.free_stride);
        const double nl_scale = (a / k0) / k0 / E0;  // kernel.py:202 (x radius), rmatrix.py:175-178 (/k0, /E0)
        const double2* Vint = A.interaction ? A.interaction + (size_t)sysi * N * N : nullptr;
        const double2* Vnl = (!A.interaction && A.nonlocal_) ? A.nonlocal_ + (size_t)sysi * N * N : nullptr;
        const double2* Vloc = (!A.interaction && A.local) ? A.local + (size_t)sysi * N : nullptr;
        // ---- assemble rows 0..N-1 (coalesced 16-byte loads, lane = column), identity padding, border
        for (int r = 0; r < N; ++r) {
            const double wr = Vnl ? A.wsqrt[r] * nl_scale : 0.0;
            for (int c = lane; c < P; c += 32) {
                double2 v = zero2;
                if (c < N) {
                    v = F[(size_t)r * N + c];
                    if (Vint) {
                        const double2 w = Vint[(size_t)r * N + c];
                        v.x += w.x; v.y += w.y;
                    }
                    if (Vnl) {
                        const double2 w = Vnl[(size_t)r * N + c];
                        const double sc = wr * A.wsqrt[c];
                        v.x = fma(w.x, sc, v.x); v.y = fma(w.y, sc, v.y);
                    }
                    if (Vloc && c == r) {
                        const double2 w = Vloc[r];
                        v.x += w.x / E0; v.y += w.y / E0;
                    }
                }
                M[r * LD + c] = v;
            }
        }
        for (int r = N; r < P; ++r)
            for (int c = lane; c < LD; c += 32) M[r * LD + c] = make_double2(r == c ? 1.0 : 0.0, 0.0);
        __syncwarp();
        for (int r = lane; r < N; r += 32) {
            const double2 bb = make_double2(A.bvec[r], 0.0);
            M[r * LD + P] = bb;
            M[P * LD + r] = bb;
        }
        for (int c = N + lane; c <= P; c += 32) M[P * LD + c] = zero2;
        __syncwarp();
        int singular = 0;
        const double2 corner = bordered_schur<MAXNS>(M, g, lane, singular);  // -b^T A^-1 b
        __syncwarp();
        if (lane == 0) {
            const double2* as = A.asym + (size_t)sysi * A.asym_stride;  // [4][1]: Hp, Hm, Hpp, Hmp
            const double2 Hp = as[0], Hm = as[1], Hpp = as[2], Hmp = as[3];
            const double2 R = make_double2(-corner.x / (a * a), -corner.y / (a * a));  // core.py:15-22
            const double2 aR = make_double2(a * R.x, a * R.y);
            const double2 tp = cmul(aR, Hpp), tm = cmul(aR, Hmp);
            // S = Z+^-1 Z-, Z+- = H+- - a R H+-'   (core.py:51-56)
            const double2 S = cdiv(make_double2(Hm.x - tm.x, Hm.y - tm.y), make_double2(Hp.x - tp.x, Hp.y - tp.y));
            // u'_ext = (i/2)(H-' w - S H+')   (core.py:58)
            const double w0 = A.weights[0];
            const double2 sh = cmul(S, Hpp);
            const double2 v = make_double2(Hmp.x * w0 - sh.x, Hmp.y * w0 - sh.y);
            A.R[sysi] = R;
            A.S[sysi] = S;
            A.uext[sysi] = make_double2(-0.5 * v.y, 0.5 * v.x);
            const bool bad = !(isfinite(S.x) && isfinite(S.y));
            A.status[sysi] = bad ? JITR_B200_STATUS_NONFINITE : (singular ? JITR_B200_STATUS_SINGULAR : JITR_B200_STATUS_OK);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Coupled-channel / large systems (BASELINE config 5: 3 channels x N = 50, n = 150): one CTA per system,
// right-looking blocked LU with partial pivoting of the BORDERED matrix
//        [ A    0   B ]    rows 0..n-1    (identity padding up to P = 8 ceil(n/8))
//        [ 0    I   0 ]
//        [ B^T  0   0 ]    rows P..P+nch-1: one "z" row per channel (never a pivot candidate)
// whose corner block is left as -B^T A^-1 B = -a^2 R (rmatrix/core.py:15-22 without the inverse).  The matrix
// (PR x PC complex128, 410 KB for config 5 -- more than one SM's shared memory) lives in a per-CTA slab of
// the device workspace, i.e. in L2 (148 x 2 slabs = 121 MB of the 126 MB); what is reused stays on chip:
//   P  the 8-column panel in REGISTERS, one matrix row per thread; per column every warp publishes its best
//      candidate row (REDUX search, speculative reciprocal, packed key) and one barrier of the panel warps later
//      a second REDUX over the eight keys elects the pivot; rows are not moved, each thread tracks the position
//      of its row;
//   U  thread c owns trailing column c: gathers the eight pivot rows from their OLD positions, moves the
//      (at most eight) displaced rows to the positions the pivots left, forward-substitutes with L11 from
//      shared memory, leaves U12 in shared memory (and in the slab when coefficients are wanted);
//   T  A22 -= L21 U12 on DMMA m8n8k4 in the real representation of the complex product (fragment mapping of
//      lu_warp.cuh), L21 and U12 from shared memory, the C tiles streamed slab -> accumulators -> slab; tile
//      column 0 (the next panel) is produced first and stays in shared memory.
// Complex symmetric systems take the LDL^T path further down (diagonal pivots, lower triangle only, panel pairs).
// Wavefunction coefficients (core.py:63-82) add a blocked back substitution on the stored U.
// ---------------------------------------------------------------------------------------------
// single-channel systems without coefficients up to this padded order go warp-per-system; beyond it the blocked
// CTA-per-system kernel is faster (measured crossover between nbasis 48 and 56, tools/bench_general_sizes.py)
#ifndef JITR_WARP_SOLVE_MAXP
#define JITR_WARP_SOLVE_MAXP 48
#endif
// CTA sizes of the blocked kernel: one matrix row per thread in the panel phase, so systems with at most 128
// bordered rows run on 128 threads (four CTAs per SM), larger ones on 256 (two per SM)

constexpr int kBlkMaxRows = 256;  // one matrix row per thread of the first eight warps
constexpr int kBlkLpLd = 9;
constexpr int kBlkHeader = 16;  // complex elements (256 bytes) in front of the slabs

// -DJITR_BLK_PROFILE: per-phase clock accumulators (thread 0 of every CTA), read by jitr_b200_debug_block_phases
#ifdef JITR_BLK_PROFILE
__device__ unsigned long long g_blk_phase[16];
#define BLK_PH_INIT long long ph_t = clock64();
#define BLK_PH(k)                                                                  \
    if (threadIdx.x == 0) {                                                        \
        const long long now = clock64();                                           \
        atomicAdd(&g_blk_phase[k], (unsigned long long)(now - ph_t));              \
        ph_t = now;                                                                \
    }
#else
#define BLK_PH_INIT
#define BLK_PH(k)
#endif

struct BlkGeom {
    int P, PR, PC;       // padded order, rows (P + z rows), columns (P + RHS columns); PR == PC
    int ub_ld;           // odd leading dimension of the U12 buffer
    int off_l11, off_ub, off_pn, off_prow, off_small, off_xs, off_int;  // shared-memory offsets in double2 units
    int smem_elems;
    int fuse, off_lp2, off_wp2;  // symmetric path: two panels per pass over the trailing tiles (second L / W buffers)
};

__host__ __device__ inline BlkGeom blk_geom(int n, int nch) {
    BlkGeom g;
    g.P = (n + 7) & ~7;
    const int zb = (nch + 7) & ~7;
    g.PR = g.P + zb;
    g.PC = g.P + zb;
    g.ub_ld = (g.PC - 8) | 1;
    int o = (g.PR - 8) * kBlkLpLd;  // Lp
    g.off_l11 = o; o += 64;
    g.off_ub = o; o += (8 * g.ub_ld > (g.PR - 8) * kBlkLpLd) ? 8 * g.ub_ld : (g.PR - 8) * kBlkLpLd;  // U12, or W of the symmetric path
    g.off_pn = o; o += g.PR * kBlkLpLd;  // the next panel: rows by position, 8 columns
    g.off_prow = o; o += 2 * 8 * 8;  // [parity][warp][8]: 1/pivot, columns j+1..7 of the warp's candidate row
    g.off_small = o; o += 3 * nch * nch + nch + 1;  // R, Z+, Z- -> S, u'_ext
    g.off_xs = o; o += 8 * kMaxCh;
    g.off_int = o; o += 24;  // 96 ints
    // fused panel pairs: only for the 256-thread kernel (two CTAs per SM must still fit), two more row buffers
    g.fuse = 0;
    g.off_lp2 = g.off_wp2 = o;
    // (measured: a loss for the 128-thread kernel even where four CTAs per SM still fit)
    if (g.PR > 128 && (size_t)(o + 2 * (g.PR - 16) * kBlkLpLd) * 16 <= 115000) {
        g.fuse = 1;
        g.off_lp2 = o; o += (g.PR - 16) * kBlkLpLd;
        g.off_wp2 = o; o += (g.PR - 16) * kBlkLpLd;
    }
    g.smem_elems = o;
    return g;
}
static inline size_t blk_smem_bytes(const BlkGeom& g) { return (size_t)g.smem_elems * 16; }

template <int NQ>
__device__ __forceinline__ void blk_trailing_item(double2* __restrict__ M, const int ld, const int R0, const int tr,
                                                  const int tc0, const double* __restrict__ Lpd,
                                                  const double* __restrict__ Ubd, const int ub_ld, const int lane) {
    const int fr = lane >> 2, fk = lane & 3;
    const int rho = (fr >> 1) + 4 * (fr & 1);
    const int part = fk & 1, pn = fr & 1;
    const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
    double2* Crow = M + (size_t)(R0 + 8 * tr + rho) * ld + R0 + 8 * tc0 + fk;
    double2 c[NQ][2];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        c[q][0] = Crow[8 * q];
        c[q][1] = Crow[8 * q + 4];
    }
    double a[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) a[s] = dneg(Lpd[2 * ((8 * tr + rho) * kBlkLpLd + s + 4 * (fk >> 1)) + part]);
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double v = Ubd[2 * ((s + 4 * (fk >> 1)) * ub_ld + 8 * (tc0 + q) + 4 * h + (fr >> 1)) + (part ^ pn)];
                const double b = __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v));
                dmma884(c[q][h].x, c[q][h].y, a[s], b);
            }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        Crow[8 * q] = c[q][0];
        Crow[8 * q + 4] = c[q][1];
    }
}

// A22 -= L21 U12: items (kBlkNQ tile columns, one tile row) strided over the warps
#ifndef JITR_BLK_NQ
#define JITR_BLK_NQ 3
#endif
constexpr int kBlkNQ = JITR_BLK_NQ;
__device__ __forceinline__ void blk_trailing_simple(double2* __restrict__ M, const int ld, const int R0, const int ntr,
                                                    const int ntc_all, const int tc_begin, const double* __restrict__ Lpd,
                                                    const double* __restrict__ Ubd, const int ub_ld, const int warp,
                                                    const int lane, const int nwarps) {
    const int ntc = ntc_all - tc_begin;
    const int ncb = (ntc + kBlkNQ - 1) / kBlkNQ;
    for (int it = warp; it < ncb * ntr; it += nwarps) {
        const int cb = it / ntr, tr = it - cb * ntr;
        const int rem = ntc - kBlkNQ * cb;  // warp-uniform
        const int tc0 = tc_begin + kBlkNQ * cb;
        if (rem >= kBlkNQ) blk_trailing_item<kBlkNQ>(M, ld, R0, tr, tc0, Lpd, Ubd, ub_ld, lane);
        else if (kBlkNQ > 3 && rem == 3) blk_trailing_item<3>(M, ld, R0, tr, tc0, Lpd, Ubd, ub_ld, lane);
        else if (kBlkNQ > 2 && rem == 2) blk_trailing_item<2>(M, ld, R0, tr, tc0, Lpd, Ubd, ub_ld, lane);
        else blk_trailing_item<1>(M, ld, R0, tr, tc0, Lpd, Ubd, ub_ld, lane);
    }
}

// ---- symmetric path (complex symmetric A: couplings with V_ij = V_ji or a symmetric user matrix on the symmetric free
// matrix) ----------------------------------------------------------------------------------------------------------
// A = L D L^T with the pivots on the diagonal, accepted while every multiplier stays below kBlkSymMaxMult (the
// threshold-pivoting test |a_kk| >= max_i |a_ik| / 64 of the fused sweep, checked on the multipliers themselves).
// Only the lower triangle of the bordered matrix is assembled and touched: half the trailing tiles, no U phase, no
// row moves -- and two barriers per panel instead of one per column: warp 0 factorises the 8 x 8 diagonal block of the
// NEXT panel with shuffles inside the trailing phase (blk_sym_diag), publishes it, and every other row is
// forward-substituted against it (blk_sym_rows).  Panel pairs share one pass over the trailing tiles where the extra
// row buffers fit (BlkGeom::fuse).  A failed test hands the rest of the system to partial pivoting (kernel body).
constexpr double kBlkSymMaxMult = 64.0;

// tile (tr, tc0 .. tc0+NQ-1) -= L[tr] W[tc]^T with the B fragments taken from the rows of W
template <int NQ>
__device__ __forceinline__ void blk_sym_item(double2* __restrict__ M, const int ld, const int R0, const int tr, const int tc0,
                                             const double* __restrict__ Lpd, const double* __restrict__ Wpd, const int lane) {
    const int fr = lane >> 2, fk = lane & 3;
    const int rho = (fr >> 1) + 4 * (fr & 1);
    const int part = fk & 1, pn = fr & 1;
    const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
    double2* Crow = M + (size_t)(R0 + 8 * tr + rho) * ld + R0 + 8 * tc0 + fk;
    double2 c[NQ][2];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        c[q][0] = Crow[8 * q];
        c[q][1] = Crow[8 * q + 4];
    }
    double a[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) a[s] = dneg(Lpd[2 * ((8 * tr + rho) * kBlkLpLd + s + 4 * (fk >> 1)) + part]);
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double v = Wpd[2 * ((8 * (tc0 + q) + 4 * h + (fr >> 1)) * kBlkLpLd + s + 4 * (fk >> 1)) + (part ^ pn)];
                dmma884(c[q][h].x, c[q][h].y, a[s], __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v)));
            }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        Crow[8 * q] = c[q][0];
        Crow[8 * q + 4] = c[q][1];
    }
}

// tile column 0 of the symmetric trailing update (= the next panel) for tile rows tr0, tr0 + step, ... -> Pn
__device__ __forceinline__ void blk_sym_col0(const double2* __restrict__ M, const int ld, const int R0, const int ntr,
                                             const int tr0, const int step, const double* __restrict__ Lpd,
                                             const double* __restrict__ Wpd, double2* __restrict__ Pn, const int lane) {
    if (tr0 >= ntr) return;
    const int fr = lane >> 2, fk = lane & 3;
    const int rho = (fr >> 1) + 4 * (fr & 1);
    const int part = fk & 1, pn = fr & 1;
    const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
    double b[2][4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double v = Wpd[2 * ((4 * h + (fr >> 1)) * kBlkLpLd + s + 4 * (fk >> 1)) + (part ^ pn)];
            b[h][s] = __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v));
        }
    for (int tr = tr0; tr < ntr; tr += step) {
        const double2* Crow = M + (size_t)(R0 + 8 * tr + rho) * ld + R0 + fk;
        double2 c0 = Crow[0], c1 = Crow[4];
        double a[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) a[s] = dneg(Lpd[2 * ((8 * tr + rho) * kBlkLpLd + s + 4 * (fk >> 1)) + part]);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            dmma884(c0.x, c0.y, a[s], b[0][s]);
            dmma884(c1.x, c1.y, a[s], b[1][s]);
        }
        double2* dst = Pn + (8 * tr + rho) * kBlkLpLd + fk;
        dst[0] = c0;
        dst[4] = c1;
    }
}

// the remaining lower-triangle tiles (tc <= tr, tc >= tcb): items (tile row, block of three tile columns) handed out
// through a shared counter, so that a warp that was busy elsewhere (warp 0: the next diagonal block) joins late
__device__ __forceinline__ void blk_sym_rest(double2* __restrict__ M, const int ld, const int R0, const int ntr, const int tcb,
                                             const double* __restrict__ Lpd, const double* __restrict__ Wpd,
                                             int* __restrict__ counter, const int lane) {
    const int ncb = (ntr - tcb + 2) / 3;
    const int total = ntr * ncb;
    while (true) {
        int it = 0;
        if (lane == 0) it = atomicAdd(counter, 1);
        it = __shfl_sync(kFull, it, 0);
        if (it >= total) break;
        // large tile rows first: the last items handed out are the short ones
        const int tr = ntr - 1 - it / ncb, cb = it % ncb;
        const int tc0 = tcb + 3 * cb;
        const int nq = tr - tc0 + 1;  // tile columns tc0 .. min(tc0 + 2, tr); warp-uniform
        if (nq <= 0) continue;
        if (nq >= 3) blk_sym_item<3>(M, ld, R0, tr, tc0, Lpd, Wpd, lane);
        else if (nq == 2) blk_sym_item<2>(M, ld, R0, tr, tc0, Lpd, Wpd, lane);
        else blk_sym_item<1>(M, ld, R0, tr, tc0, Lpd, Wpd, lane);
    }
}

// ---- two panels per pass (A = columns k0.., B = columns k0+8..): C -= L_A W_A^T + L_B W_B^T for the tiles at and below
// R2 = k0 + 16.  L_A / W_A are indexed from R0 = k0 + 8 (one tile above R2), L_B / W_B from R2.
template <int NQ>
__device__ __forceinline__ void blk_sym_item2(double2* __restrict__ M, const int ld, const int R2, const int tr, const int tc0,
                                              const double* __restrict__ LAd, const double* __restrict__ WAd,
                                              const double* __restrict__ LBd, const double* __restrict__ WBd, const int lane) {
    const int fr = lane >> 2, fk = lane & 3;
    const int rho = (fr >> 1) + 4 * (fr & 1);
    const int part = fk & 1, pn = fr & 1;
    const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
    double2* Crow = M + (size_t)(R2 + 8 * tr + rho) * ld + R2 + 8 * tc0 + fk;
    double2 c[NQ][2];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        c[q][0] = Crow[8 * q];
        c[q][1] = Crow[8 * q + 4];
    }
#pragma unroll
    for (int ab = 0; ab < 2; ++ab) {
        const double* Ld = ab ? LBd : LAd + 2 * 8 * kBlkLpLd;  // L_A rows start one tile earlier
        const double* Wd = ab ? WBd : WAd + 2 * 8 * kBlkLpLd;
        double a[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) a[s] = dneg(Ld[2 * ((8 * tr + rho) * kBlkLpLd + s + 4 * (fk >> 1)) + part]);
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double v = Wd[2 * ((8 * (tc0 + q) + 4 * h + (fr >> 1)) * kBlkLpLd + s + 4 * (fk >> 1)) + (part ^ pn)];
                    dmma884(c[q][h].x, c[q][h].y, a[s], __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v)));
                }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        Crow[8 * q] = c[q][0];
        Crow[8 * q + 4] = c[q][1];
    }
}

// tile column 0 of the fused update (= the panel after the pair) for tile rows tr0, tr0 + step, ... -> Pn
__device__ __forceinline__ void blk_sym_col0_2(const double2* __restrict__ M, const int ld, const int R2, const int ntr,
                                               const int tr0, const int step, const double* __restrict__ LAd,
                                               const double* __restrict__ WAd, const double* __restrict__ LBd,
                                               const double* __restrict__ WBd, double2* __restrict__ Pn, const int lane) {
    if (tr0 >= ntr) return;
    const int fr = lane >> 2, fk = lane & 3;
    const int rho = (fr >> 1) + 4 * (fr & 1);
    const int part = fk & 1, pn = fr & 1;
    const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
    double b[2][2][4];
#pragma unroll
    for (int ab = 0; ab < 2; ++ab) {
        const double* Wd = ab ? WBd : WAd + 2 * 8 * kBlkLpLd;
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double v = Wd[2 * ((4 * h + (fr >> 1)) * kBlkLpLd + s + 4 * (fk >> 1)) + (part ^ pn)];
                b[ab][h][s] = __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v));
            }
    }
    for (int tr = tr0; tr < ntr; tr += step) {
        const double2* Crow = M + (size_t)(R2 + 8 * tr + rho) * ld + R2 + fk;
        double2 c0 = Crow[0], c1 = Crow[4];
#pragma unroll
        for (int ab = 0; ab < 2; ++ab) {
            const double* Ld = ab ? LBd : LAd + 2 * 8 * kBlkLpLd;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double a = dneg(Ld[2 * ((8 * tr + rho) * kBlkLpLd + s + 4 * (fk >> 1)) + part]);
                dmma884(c0.x, c0.y, a, b[ab][0][s]);
                dmma884(c1.x, c1.y, a, b[ab][1][s]);
            }
        }
        double2* dst = Pn + (8 * tr + rho) * kBlkLpLd + fk;
        dst[0] = c0;
        dst[4] = c1;
    }
}

__device__ __forceinline__ void blk_sym_rest2(double2* __restrict__ M, const int ld, const int R2, const int ntr, const int tcb,
                                              const double* __restrict__ LAd, const double* __restrict__ WAd,
                                              const double* __restrict__ LBd, const double* __restrict__ WBd,
                                              int* __restrict__ counter, const int lane) {
    const int ncb = (ntr - tcb + 2) / 3;
    const int total = ntr * ncb;
    while (true) {
        int it = 0;
        if (lane == 0) it = atomicAdd(counter, 1);
        it = __shfl_sync(kFull, it, 0);
        if (it >= total) break;
        const int tr = ntr - 1 - it / ncb, cb = it % ncb;
        const int tc0 = tcb + 3 * cb;
        const int nq = tr - tc0 + 1;
        if (nq <= 0) continue;
        if (nq >= 3) blk_sym_item2<3>(M, ld, R2, tr, tc0, LAd, WAd, LBd, WBd, lane);
        else if (nq == 2) blk_sym_item2<2>(M, ld, R2, tr, tc0, LAd, WAd, LBd, WBd, lane);
        else blk_sym_item2<1>(M, ld, R2, tr, tc0, LAd, WAd, LBd, WBd, lane);
    }
}

// rows below the diagonal block of the panel at kp (its columns are in Pn, the factorised block in L11 / slot): forward
// substitution, multipliers -> Lx (scaled) and Wx (unscaled), growth check, U rows for the coefficients
__device__ __forceinline__ void blk_sym_rows(const double2* __restrict__ Pn, const double2* __restrict__ L11,
                                             const double2* __restrict__ slot, double2* __restrict__ Lx, double2* __restrict__ Wx,
                                             double2* __restrict__ M, const int ld, const int kp, const int P, const int PR,
                                             const bool coeffs, int* __restrict__ fail, const int tid) {
    if (tid < 8 || tid >= PR - kp) return;
    double2 p[8], w[8];
    {
        const double2* rowp = Pn + tid * kBlkLpLd;
#pragma unroll
        for (int c = 0; c < 8; ++c) p[c] = rowp[c];
    }
    // the growth bound applies to the rows of A; the border rows only have to stay finite
    const double lim = (kp + tid < P) ? kBlkSymMaxMult : 1.0e300;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        w[j] = p[j];
        const double2 l = cmul(p[j], slot[j]);
        p[j] = l;
        bad |= !(fabs(l.x) <= lim && fabs(l.y) <= lim);
#pragma unroll
        for (int c = j + 1; c < 8; ++c) cfnma(p[c], l, L11[j * 8 + c]);
    }
    double2* lp = Lx + (tid - 8) * kBlkLpLd;
    double2* wp = Wx + (tid - 8) * kBlkLpLd;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        lp[c] = p[c];
        wp[c] = w[c];
    }
    if (bad) *fail = 1;
    if (coeffs) {
        // U = D L^T for the back substitution of the coefficients: row kp + c of U holds the unscaled multipliers w_c
        // of the rows below (the border rows give Y = L^-1 B in the RHS columns)
#pragma unroll
        for (int c = 0; c < 8; ++c) M[(size_t)(kp + c) * ld + kp + tid] = w[c];
    }
}

// 8 x 8 diagonal block of the panel whose columns are in Pn (rows 0..7), by ONE warp with shuffles: A11 = L11 D11 L11^T.
// Publishes row j as  [l_j0 .. l_j,j-1, d_j, d_j l_j+1,j .. d_j l_7j]  in L11 and 1 / d_j in slot[j]; flags a multiplier
// beyond the bound.  With coefficients the rows of U11 = D11 L11^T go to the slab (rows kp.., columns kp..).
__device__ __forceinline__ void blk_sym_diag(const double2* __restrict__ Pn, double2* __restrict__ L11, double2* __restrict__ slot,
                                             double2* __restrict__ M, const int ld, const int kp, const bool coeffs,
                                             int* __restrict__ fail, const int lane) {
    double2 p[8];
    {
        const double2* rowp = Pn + (lane < 8 ? lane : 0) * kBlkLpLd;
#pragma unroll
        for (int c = 0; c < 8; ++c) p[c] = rowp[c];
    }
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double2 rinv = crecip_fast(shfl2(p[j], j));
        double2 rowj[8];
#pragma unroll
        for (int c = j + 1; c < 8; ++c) rowj[c] = shfl2(p[c], j);
        if (lane == 0) slot[j] = rinv;
        if (lane > j && lane < 8) {
            const double2 l = cmul(p[j], rinv);
            p[j] = l;
            bad |= !(fabs(l.x) <= kBlkSymMaxMult && fabs(l.y) <= kBlkSymMaxMult);
#pragma unroll
            for (int c = j + 1; c < 8; ++c) cfnma(p[c], l, rowj[c]);
        }
        if (lane == j) bad |= !(isfinite(rinv.x) && isfinite(rinv.y));  // zero pivot
    }
    if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) L11[lane * 8 + c] = p[c];
        if (coeffs) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c >= lane) M[(size_t)(kp + lane) * ld + kp + c] = p[c];
        }
        if (bad) *fail = 1;
    }
}

// the same with symmetric 2 x 2 pivots (pair_mask from blk_sym_diag22): only called to RETRY a panel that failed the test,
// out of line so that the 1 x 1 code of the hot loop keeps its registers
// rows below the diagonal block of the panel at kp (its columns are in Pn, the factorised block in L11 / slot): forward
// substitution, multipliers -> Lx (scaled) and Wx (unscaled), growth check, U rows for the coefficients
__device__ __noinline__ void blk_sym_rows22(const double2* __restrict__ Pn, const double2* __restrict__ L11,
                                             const double2* __restrict__ slot, double2* __restrict__ Lx, double2* __restrict__ Wx,
                                             double2* __restrict__ M, const int ld, const int kp, const int P, const int PR,
                                             const bool coeffs, int* __restrict__ fail, const int pair_mask, const int tid) {
    if (tid < 8 || tid >= PR - kp) return;
    double2 p[8], w[8];
    {
        const double2* rowp = Pn + tid * kBlkLpLd;
#pragma unroll
        for (int c = 0; c < 8; ++c) p[c] = rowp[c];
    }
    // the growth bound applies to the rows of A; the border rows only have to stay finite
    const double lim = (kp + tid < P) ? kBlkSymMaxMult : 1.0e300;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j > 0 && ((pair_mask >> (j - 1)) & 1)) continue;  // second column of a 2 x 2 pivot: done with the first
        if (j < 7 && ((pair_mask >> j) & 1)) {
            // 2 x 2 pivot (columns j, j + 1):  [l0 l1] = [w0 w1] E,  E = D^-1 = [[slot[j], slot[8 + j]], [slot[8 + j], slot[j + 1]]]
            const double2 w0 = p[j], w1 = p[j + 1];
            w[j] = w0;
            w[j + 1] = w1;
            const double2 e00 = slot[j], e01 = slot[8 + j], e11 = slot[j + 1];
            double2 l0 = cmul(w0, e00), l1 = cmul(w0, e01);
            const double2 t0 = cmul(w1, e01), t1 = cmul(w1, e11);
            l0.x += t0.x; l0.y += t0.y;
            l1.x += t1.x; l1.y += t1.y;
            p[j] = l0;
            p[j + 1] = l1;
            bad |= !(fabs(l0.x) <= lim && fabs(l0.y) <= lim && fabs(l1.x) <= lim && fabs(l1.y) <= lim);
#pragma unroll
            for (int c = j + 2; c < 8; ++c) {
                cfnma(p[c], l0, L11[j * 8 + c]);
                cfnma(p[c], l1, L11[(j + 1) * 8 + c]);
            }
        } else {
            w[j] = p[j];
            const double2 l = cmul(p[j], slot[j]);
            p[j] = l;
            bad |= !(fabs(l.x) <= lim && fabs(l.y) <= lim);
#pragma unroll
            for (int c = j + 1; c < 8; ++c) cfnma(p[c], l, L11[j * 8 + c]);
        }
    }
    double2* lp = Lx + (tid - 8) * kBlkLpLd;
    double2* wp = Wx + (tid - 8) * kBlkLpLd;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        lp[c] = p[c];
        wp[c] = w[c];
    }
    if (bad) *fail = 1;
    if (coeffs) {
        // U = D L^T for the back substitution of the coefficients: row kp + c of U holds the unscaled multipliers w_c
        // of the rows below (the border rows give Y = L^-1 B in the RHS columns)
#pragma unroll
        for (int c = 0; c < 8; ++c) M[(size_t)(kp + c) * ld + kp + tid] = w[c];
    }
}

// blk_sym_diag with symmetric 2 x 2 pivots, for the retry of a panel that failed the 1 x 1 test.
// Symmetric 2 x 2 pivots (round 2).  With a purely real potential (examples/coupled.py) the matrix is real symmetric INDEFINITE and
// a diagonal entry of a Schur complement passes through zero now and then: 86 % of the l = 0 systems of BASELINE config 5 met one
// somewhere and left the symmetric path.  When |d_j| < |a_j+1,j| / 16 (and j < 7: the pair must not straddle the panel) the columns
// j, j + 1 are eliminated together with the 2 x 2 block D = [[d_j, a_j+1,j], [a_j+1,j, d_j+1]] as the pivot (Bunch-Kaufman without
// interchanges): [l0 l1] = [w0 w1] D^-1, trailing update  -= l0 w0^T + l1 w1^T  -- the same L W^T form, so the tile kernels do not
// change.  pair_mask bit j = columns j, j + 1 form a pair; D^-1 = [[slot[j], slot[8+j]], [slot[8+j], slot[j+1]]].  The multiplier
// bound stays the safety net (9 % of the l = 0 systems still fail it and finish with partial pivoting; 0 % from l = 3 on).
// Not used when coefficients are wanted (the back substitution assumes a triangular U).
__device__ __noinline__ void blk_sym_diag22(const double2* __restrict__ Pn, double2* __restrict__ L11, double2* __restrict__ slot,
                                             double2* __restrict__ M, const int ld, const int kp, const bool coeffs,
                                             int* __restrict__ fail, int* __restrict__ pair_mask_out, const int lane) {
    double2 p[8];
    {
        const double2* rowp = Pn + (lane < 8 ? lane : 0) * kBlkLpLd;
#pragma unroll
        for (int c = 0; c < 8; ++c) p[c] = rowp[c];
    }
    bool bad = false;
    int mask = 0;
    bool skip = false;  // warp-uniform: column j is the second of a pair
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (skip) {
            skip = false;
            continue;
        }
        bool pair = false;
        if (j < 7 && !coeffs) {
            // decided by the pivot's own lane on max(|re|, |im|), broadcast
            const double ma = fmax(fabs(p[j].x), fabs(p[j].y)), mb = fmax(fabs(p[j + 1].x), fabs(p[j + 1].y));
            pair = __shfl_sync(kFull, (int)(16.0 * ma < mb), j) != 0;
        }
        if (pair) {
            // D = [[a, b], [b, d]] : a, b on lane j (p[j], p[j+1]), d on lane j + 1 (p[j+1])
            const double2 a = shfl2(p[j], j), b = shfl2(p[j + 1], j), d = shfl2(p[j + 1], j + 1);
            double2 det = cmul(a, d);
            const double2 bb = cmul(b, b);
            det.x -= bb.x; det.y -= bb.y;
            const double2 dinv = crecip_fast(det);
            const double2 e00 = cmul(d, dinv), e11 = cmul(a, dinv);
            double2 e01 = cmul(b, dinv);
            e01.x = -e01.x; e01.y = -e01.y;
            double2 r0[8], r1[8];
#pragma unroll
            for (int c = j + 2; c < 8; ++c) {
                r0[c] = shfl2(p[c], j);
                r1[c] = shfl2(p[c], j + 1);
            }
            if (lane == 0) {
                slot[j] = e00;
                slot[j + 1] = e11;
                slot[8 + j] = e01;
            }
            if (lane > j + 1 && lane < 8) {
                const double2 w0 = p[j], w1 = p[j + 1];
                double2 l0 = cmul(w0, e00), l1 = cmul(w0, e01);
                const double2 t0 = cmul(w1, e01), t1 = cmul(w1, e11);
                l0.x += t0.x; l0.y += t0.y;
                l1.x += t1.x; l1.y += t1.y;
                p[j] = l0;
                p[j + 1] = l1;
                bad |= !(fabs(l0.x) <= kBlkSymMaxMult && fabs(l0.y) <= kBlkSymMaxMult && fabs(l1.x) <= kBlkSymMaxMult &&
                         fabs(l1.y) <= kBlkSymMaxMult);
#pragma unroll
                for (int c = j + 2; c < 8; ++c) {
                    cfnma(p[c], l0, r0[c]);
                    cfnma(p[c], l1, r1[c]);
                }
            }
            if (lane == j) bad |= !(isfinite(dinv.x) && isfinite(dinv.y));  // singular 2 x 2 block
            mask |= 1 << j;
            skip = true;
        } else {
            const double2 rinv = crecip_fast(shfl2(p[j], j));
            double2 rowj[8];
#pragma unroll
            for (int c = j + 1; c < 8; ++c) rowj[c] = shfl2(p[c], j);
            if (lane == 0) slot[j] = rinv;
            if (lane > j && lane < 8) {
                const double2 l = cmul(p[j], rinv);
                p[j] = l;
                bad |= !(fabs(l.x) <= kBlkSymMaxMult && fabs(l.y) <= kBlkSymMaxMult);
#pragma unroll
                for (int c = j + 1; c < 8; ++c) cfnma(p[c], l, rowj[c]);
            }
            if (lane == j) bad |= !(isfinite(rinv.x) && isfinite(rinv.y));  // zero pivot
        }
    }
    if (lane == 0) *pair_mask_out = mask;
    if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; ++c) L11[lane * 8 + c] = p[c];
        if (coeffs) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c >= lane) M[(size_t)(kp + lane) * ld + kp + c] = p[c];
        }
        if (bad) *fail = 1;
    }
}

// Wavefunction coefficients (core.py:63-82): X = U^-1 Y for the nch right-hand sides (columns P..P+nch-1 of the
// slab, which holds U and Y = L^-1 P B after the elimination), eight rows at a time from the bottom; then
// x = sum_i u'_ext,i X[:, i].  Out of line: keeps its registers out of the elimination loop.
__device__ __noinline__ void blk_coefficients(const GenArgs& A, const long long sysi, double2* __restrict__ M, const int ld,
                                              const int P, double2* __restrict__ xs, const double2* __restrict__ ue) {
    const int n = A.n, nch = A.nch, tid = threadIdx.x;
    const double2 zero2 = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int kb = P - 8; kb >= 0; kb -= 8) {
        if (tid < nch) {
            double2 x[8];
#pragma unroll
            for (int r = 7; r >= 0; --r) {
                double2 v = M[(size_t)(kb + r) * ld + P + tid];
#pragma unroll
                for (int c = r + 1; c < 8; ++c) cfnma(v, M[(size_t)(kb + r) * ld + kb + c], x[c]);
                x[r] = cmul(v, crecip(M[(size_t)(kb + r) * ld + kb + r]));
                xs[r * kMaxCh + tid] = x[r];
                M[(size_t)(kb + r) * ld + P + tid] = x[r];
            }
        }
        __syncthreads();
        for (int t = tid; t < kb * nch; t += blockDim.x) {
            const int r = t / nch, j = t - r * nch;
            double2 v = M[(size_t)r * ld + P + j];
#pragma unroll
            for (int c = 0; c < 8; ++c) cfnma(v, M[(size_t)r * ld + kb + c], xs[c * kMaxCh + j]);
            M[(size_t)r * ld + P + j] = v;
        }
        __syncthreads();
    }
    // x = A^-1 vec_i(b u'_ext,i) = sum_i u'_ext,i X[:, i]    (core.py:81-82)
    for (int r = tid; r < n; r += blockDim.x) {
        double2 acc = zero2;
        for (int i = 0; i < nch; ++i) {
            const double2 t = cmul(ue[i], M[(size_t)r * ld + P + i]);
            acc.x += t.x; acc.y += t.y;
        }
        A.coeffs[(size_t)sysi * n + r] = acc;
    }
}

// Tile column 0 of the trailing update = the columns of the NEXT panel: computed first, by all warps, and left
// in shared memory (Pn, rows by position): panels never reload from the slab.
__device__ __forceinline__ void blk_trailing_col0(const double2* __restrict__ M, const int ld, const int R0,
                                                  const int ntr, const double* __restrict__ Lpd,
                                                  const double* __restrict__ Ubd, const int ub_ld,
                                                  double2* __restrict__ Pn, const int warp, const int lane,
                                                  const int nwarps) {
    const int fr = lane >> 2, fk = lane & 3;
    const int rho = (fr >> 1) + 4 * (fr & 1);
    const int part = fk & 1, pn = fr & 1;
    const int bflip = (part & (pn ^ 1)) ? 0x80000000 : 0;
    const int a_off = 2 * (rho * kBlkLpLd + 4 * (fk >> 1)) + part;
    const int b_off = 2 * ((4 * (fk >> 1)) * ub_ld + (fr >> 1)) + (part ^ pn);
    if (warp >= ntr) return;
    double b[2][4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double v = Ubd[b_off + 2 * (s * ub_ld + 4 * h)];
            b[h][s] = __hiloint2double(__double2hiint(v) ^ bflip, __double2loint(v));
        }
#pragma unroll 1
    for (int tr = warp; tr < ntr; tr += nwarps) {
        const double2* Crow = M + (size_t)(R0 + 8 * tr + rho) * ld + R0 + fk;
        double2 c0 = Crow[0], c1 = Crow[4];
        double a[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) a[s] = dneg(Lpd[a_off + 2 * (8 * tr * kBlkLpLd + s)]);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            dmma884(c0.x, c0.y, a[s], b[0][s]);
            dmma884(c1.x, c1.y, a[s], b[1][s]);
        }
        double2* dst = Pn + (8 * tr + rho) * kBlkLpLd + fk;
        dst[0] = c0;
        dst[4] = c1;
    }
}

template <int kBlkThreads>
__global__ void __launch_bounds__(kBlkThreads, 512 / kBlkThreads) block_solve_kernel(const GenArgs A, const BlkGeom G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = A.n, N = A.N, nch = A.nch;
    const int P = G.P, PR = G.PR, PC = G.PC, ld = G.PC;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(kFull, tid >> 5, 0);
    constexpr int NW = kBlkThreads / 32;
    double2* sm = reinterpret_cast<double2*>(smem_raw);
    double2* Lp = sm;                    // [(PR - 8)][9]  multipliers of the rows below the panel, by position
    double2* L11 = sm + G.off_l11;       // [8][8]         multipliers inside the pivot rows
    double2* Ub = sm + G.off_ub;         // [8][ub_ld]     U12 (trailing columns incl. RHS)
    double2* Pn = sm + G.off_pn;         // [PR][9]        columns of the next panel, rows by position
    double2* slot = sm + G.off_prow;     // [2][NW][8] candidate rows: [0] = 1/pivot, [c] = column c (c > j)
    double2* small = sm + G.off_small;
    double2* xs = sm + G.off_xs;         // [8][kMaxCh] back-substitution block
    int* si = reinterpret_cast<int*>(sm + G.off_int);
    int* skey = si;                      // [2][8] pivot key of each warp's candidate (-2: warp holds no rows)
    int* spos = si + 32;                 // [2][8] its current position
    int* src = si + 48;                  // [8] old position of pivot row j
    int* dfrom = si + 56;                // [8]
    int* dto = si + 64;                  // [8]
    int* misc = si + 72;                 // [1] displaced count, [2] flags, [3] couplings not symmetric, [4] symmetric path failed, [5] at panel, [6] trailing item counter
    double2* M = A.workspace + kBlkHeader + (size_t)blockIdx.x * PR * ld;  // slabs follow a small header (symmetry flag)
    const double2 zero2 = make_double2(0.0, 0.0);

    BLK_PH_INIT
    const long long nsys = A.sys_list ? (long long)*A.sys_count : A.batch;
    for (long long isys = blockIdx.x; isys < nsys; isys += gridDim.x) {
        const long long sysi = A.sys_list ? (long long)A.sys_list[isys] : isys;
        const double* sy = A.sys + (size_t)sysi * A.sys_stride;
        const double a = sy[0], E0 = sy[1], k0s = sy[2];
        const double2* F = A.free_matrix + (A.free_index ? (size_t)A.free_index[sysi] * A.n * A.n : (size_t)sysi * A.free_stride);
        const double nl_scale = (a / k0s) / k0s / E0;  // kernel.py:202 (x radius), rmatrix.py:175-178 (/k0, /E0)
        // complex symmetric systems (local couplings on the symmetric free matrix) are eliminated with diagonal pivots;
        // from the first panel with a multiplier beyond the threshold on, partial pivoting takes over
        const bool try_sym = A.symflag != nullptr && *A.symflag == 0;
        int kstart = 0;     // first panel of the partial-pivoting loop
        bool solved = false;
        if (tid == 0) misc[2] = misc[3] = misc[4] = 0;
        const bool local_only = A.local && !A.nonlocal_ && !A.interaction;
        if (try_sym && local_only && nch > 1) {
            // V_ij = V_ji ?  (decides whether the lower triangle alone is assembled)
            __syncthreads();
            for (int t = tid; t < nch * nch * N; t += kBlkThreads) {
                const int cij = t / N, nn = t - cij * N;
                const int ci = cij / nch, cj = cij - ci * nch;
                if (ci < cj) {
                    const double2 w = A.local[((size_t)sysi * nch * nch + cij) * N + nn];
                    const double2 wt = A.local[(((size_t)sysi * nch + cj) * nch + ci) * N + nn];
                    if (wt.x != w.x || wt.y != w.y) misc[3] = 1;
                }
            }
        }
        __syncthreads();
        // the symmetric path never reads the tiles right of the diagonal tile (a fall-back rebuilds them by symmetry)
        const bool lower_only = try_sym && local_only && misc[3] == 0;
        // ---------------- assemble the bordered matrix into the slab
        for (int r = warp; r < PR; r += NW) {
            const int ci = r / N, nn = r - ci * N;
            double2* __restrict__ Mr = M + (size_t)r * ld;
            if (r < n) {
                const int ncopy = lower_only ? min(n, ((r >> 3) + 1) << 3) : n;  // columns of A to assemble in this row
                const double2* __restrict__ Fr = F + (size_t)r * n;
                const double2* __restrict__ Ir = A.interaction ? A.interaction + ((size_t)sysi * n + r) * n : nullptr;
                // four independent 16-byte loads in flight per lane (the free matrix comes from L2)
                for (int c0 = 0; c0 < ncopy; c0 += 128) {
                    double2 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = c0 + 32 * i + lane;
                        v[i] = c < ncopy ? Fr[c] : zero2;
                    }
                    if (Ir) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int c = c0 + 32 * i + lane;
                            if (c < ncopy) {
                                const double2 w = Ir[c];
                                v[i].x += w.x; v[i].y += w.y;
                            }
                        }
                    } else if (A.nonlocal_) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int c = c0 + 32 * i + lane;
                            if (c < ncopy) {
                                const int cj = c / N, mm = c - cj * N;
                                const double2 w = A.nonlocal_[((((size_t)sysi * nch + ci) * nch + cj) * N + nn) * N + mm];
                                const double sc = A.wsqrt[nn] * A.wsqrt[mm] * nl_scale;
                                v[i].x += w.x * sc; v[i].y += w.y * sc;
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = c0 + 32 * i + lane;
                        if (c < ncopy) {
                            Mr[c] = v[i];
                            if (c < 8) Pn[r * kBlkLpLd + c] = v[i];
                        }
                    }
                }
                // padding columns and right-hand sides (lower_only: only up to the end of the diagonal tile)
                const int cend = lower_only ? min(PC, ((r >> 3) + 1) << 3) : PC;
                for (int c = n + lane; c < cend; c += 32) {
                    const double2 v = make_double2((c >= P && c - P == ci) ? A.bvec[nn] : 0.0, 0.0);
                    Mr[c] = v;
                    if (c < 8) Pn[r * kBlkLpLd + c] = v;  // only when n < 8
                }
            } else if (r < P) {
                for (int c = lane; c < PC; c += 32) Mr[c] = make_double2(c == r ? 1.0 : 0.0, 0.0);
                if (lane < 8) Pn[r * kBlkLpLd + lane] = make_double2(lane == r ? 1.0 : 0.0, 0.0);
            } else {
                const int zi = r - P;  // z row of channel zi: b in the columns of that channel
                for (int c = lane; c < PC; c += 32) {
                    const int c0 = c - zi * N;
                    const double2 v = make_double2((zi < nch && c0 >= 0 && c0 < N) ? A.bvec[c0] : 0.0, 0.0);
                    Mr[c] = v;
                    if (c < 8) Pn[r * kBlkLpLd + c] = v;
                }
            }
        }
        if (!A.interaction && A.local) {
            // local (coupling) potentials sit on the diagonals of the channel blocks
            __syncthreads();
            for (int t = tid; t < nch * nch * N; t += kBlkThreads) {
                const int cij = t / N, nn = t - cij * N;
                const int ci = cij / nch, cj = cij - ci * nch;
                const int er = ci * N + nn, ec = cj * N + nn;
                if (lower_only && ec >= (((er >> 3) + 1) << 3)) continue;  // right of the diagonal tile
                const double2 w = A.local[((size_t)sysi * nch * nch + cij) * N + nn];
                double2* e = M + (size_t)er * ld + ec;
                double2 v = *e;
                v.x += w.x / E0; v.y += w.y / E0;
                *e = v;
                if (cj * N + nn < 8) Pn[(ci * N + nn) * kBlkLpLd + cj * N + nn] = v;
            }
        }
        __syncthreads();
        if (try_sym && (A.nonlocal_ || A.interaction)) {
            // user matrices: symmetric only if the assembled block says so (read back from the slab, i.e. from L2)
            for (int r = warp; r < n; r += NW)
                for (int c = lane; c < r; c += 32) {
                    const double2 x = M[(size_t)r * ld + c], y = M[(size_t)c * ld + r];
                    if (x.x != y.x || x.y != y.y) misc[3] = 1;
                }
            __syncthreads();
        }
        BLK_PH(0)
        if (try_sym && misc[3] == 0) {
            // ---------------- symmetric elimination: two barriers per panel; the diagonal block of the next panel is
            // factorised by warp 0 inside the trailing phase of the current one.  Where the extra row buffers fit
            // (G.fuse), two panels A, B share one pass over the trailing tiles (half the slab traffic).
            double2* Wp = Ub;  // [(PR - 8)][9] unscaled multipliers (= D L^T transposed), rows by position
            double2* Lp2 = sm + G.off_lp2;
            double2* Wp2 = sm + G.off_wp2;
            const double* Lpd = reinterpret_cast<const double*>(Lp);
            const double* Wpd = reinterpret_cast<const double*>(Wp);
            const double* Lp2d = reinterpret_cast<const double*>(Lp2);
            const double* Wp2d = reinterpret_cast<const double*>(Wp2);
            const bool coeffs = A.coeffs != nullptr;
            if (warp == 0) blk_sym_diag(Pn, L11, slot, M, ld, 0, coeffs, &misc[4], lane);
            __syncthreads();
            int k0 = 0;
#pragma unroll 1
            while (k0 < P) {
                blk_sym_rows(Pn, L11, slot, Lp, Wp, M, ld, k0, P, PR, coeffs, &misc[4], tid);
                if (tid == 0) misc[6] = 0;  // item counter of the trailing phase
                BLK_PH(10)
                __syncthreads();
                BLK_PH(11)
                if (misc[4] && !coeffs) {
                    // panel k0 failed the 1 x 1 test (in its diagonal block or in a row below): retry it with symmetric 2 x 2
                    // pivots -- Pn still holds its columns
                    __syncthreads();
                    if (tid == 0) misc[4] = 0;
                    __syncthreads();
                    if (warp == 0) blk_sym_diag22(Pn, L11, slot, M, ld, k0, coeffs, &misc[4], &misc[7], lane);
                    __syncthreads();
                    blk_sym_rows22(Pn, L11, slot, Lp, Wp, M, ld, k0, P, PR, coeffs, &misc[4], misc[7], tid);
                    __syncthreads();
                }
                if (misc[4]) {  // CTA-uniform: a multiplier beyond the threshold (or not finite) in panel k0
                    if (tid == 0) misc[5] = k0;
                    __syncthreads();
                    break;
                }
                const int R0 = k0 + 8, ntr = (PR - R0) >> 3;
                const bool pair = G.fuse && R0 + 8 <= P;
                if (!pair) {
                    const bool next_panel = R0 < P;
                    if (next_panel) {
                        if (warp == 0) {
                            blk_sym_col0(M, ld, R0, ntr, 0, ntr, Lpd, Wpd, Pn, lane);  // the next diagonal tile first
                            __syncwarp();
                            blk_sym_diag(Pn, L11, slot, M, ld, R0, coeffs, &misc[4], lane);
                            blk_sym_col0(M, ld, R0, ntr, NW, NW, Lpd, Wpd, Pn, lane);
                        } else {
                            blk_sym_col0(M, ld, R0, ntr, warp, NW, Lpd, Wpd, Pn, lane);
                        }
                    }
                    blk_sym_rest(M, ld, R0, ntr, next_panel ? 1 : 0, Lpd, Wpd, &misc[6], lane);
                    BLK_PH(12)
                    __syncthreads();
                    BLK_PH(13)
                    k0 += 8;
                    continue;
                }
                // ---- pair: panel B = the columns R0.. ; only they get panel A's update now
                if (warp == 0) {
                    blk_sym_col0(M, ld, R0, ntr, 0, ntr, Lpd, Wpd, Pn, lane);
                    __syncwarp();
                    blk_sym_diag(Pn, L11, slot, M, ld, R0, coeffs, &misc[4], lane);
                    blk_sym_col0(M, ld, R0, ntr, NW, NW, Lpd, Wpd, Pn, lane);
                } else {
                    blk_sym_col0(M, ld, R0, ntr, warp, NW, Lpd, Wpd, Pn, lane);
                }
                __syncthreads();
                blk_sym_rows(Pn, L11, slot, Lp2, Wp2, M, ld, R0, P, PR, coeffs, &misc[4], tid);
                __syncthreads();
                if (misc[4] && !coeffs) {  // retry panel B with symmetric 2 x 2 pivots
                    __syncthreads();
                    if (tid == 0) misc[4] = 0;
                    __syncthreads();
                    if (warp == 0) blk_sym_diag22(Pn, L11, slot, M, ld, R0, coeffs, &misc[4], &misc[7], lane);
                    __syncthreads();
                    blk_sym_rows22(Pn, L11, slot, Lp2, Wp2, M, ld, R0, P, PR, coeffs, &misc[4], misc[7], tid);
                    __syncthreads();
                }
                if (misc[4]) {
                    // panel B failed: complete panel A's update of the slab, then partial pivoting from R0
                    blk_sym_rest(M, ld, R0, ntr, 1, Lpd, Wpd, &misc[6], lane);
                    if (tid == 0) misc[5] = R0;
                    __syncthreads();
                    break;
                }
                {
                    const int R2 = R0 + 8, ntr2 = ntr - 1;
                    const bool next_panel = R2 < P;
                    if (next_panel) {
                        if (warp == 0) {
                            blk_sym_col0_2(M, ld, R2, ntr2, 0, ntr2, Lpd, Wpd, Lp2d, Wp2d, Pn, lane);
                            __syncwarp();
                            blk_sym_diag(Pn, L11, slot, M, ld, R2, coeffs, &misc[4], lane);
                            blk_sym_col0_2(M, ld, R2, ntr2, NW, NW, Lpd, Wpd, Lp2d, Wp2d, Pn, lane);
                        } else {
                            blk_sym_col0_2(M, ld, R2, ntr2, warp, NW, Lpd, Wpd, Lp2d, Wp2d, Pn, lane);
                        }
                    }
                    blk_sym_rest2(M, ld, R2, ntr2, next_panel ? 1 : 0, Lpd, Wpd, Lp2d, Wp2d, &misc[6], lane);
                }
                BLK_PH(12)
                __syncthreads();
                BLK_PH(13)
                k0 += 16;
            }
            solved = misc[4] == 0;
            if (!solved) {
                // panel kfail failed the test.  Everything eliminated so far stands; the trailing matrix is still
                // symmetric, so its upper triangle (incl. the right-hand-side columns = the border rows transposed) is
                // rebuilt from the lower one and from the panel in Pn, and partial pivoting continues from this panel.
                const int kfail = misc[5];
                const int R1 = kfail + 8;
                if (tid == 0 && A.fallbacks) atomicAdd(A.fallbacks, 1ULL);
                for (int r = R1 + warp; r < PR; r += NW) {
                    const int cfirst = R1 + ((((r - R1) >> 3) + 1) << 3);  // tiles right of the diagonal tile
                    for (int c = cfirst + lane; c < PC; c += 32) M[(size_t)r * ld + c] = M[(size_t)c * ld + r];
                }
                const int wt = PC - R1;
                for (int t = tid; t < 8 * wt; t += kBlkThreads) {
                    const int j = t / wt, c = R1 + (t - j * wt);
                    M[(size_t)(kfail + j) * ld + c] = Pn[(c - kfail) * kBlkLpLd + j];
                }
                kstart = kfail;
                __syncthreads();
            }
        }
        if (!solved) {
        // ---------------- blocked elimination with partial pivoting
#pragma unroll 1
        for (int k0 = kstart; k0 < P; k0 += 8) {
            const int m = PR - k0;
            const bool have = tid < m;
            const int start = k0 + tid;
            int pos = start;
            const bool cand = have && start < P;
            bool done = false;
            double2 p[8];
            const int nw = (m + 31) >> 5;  // warps that hold rows (<= 8)
            if (warp < nw) {
                {
                    const double2* rowp = Pn + (have ? tid : 0) * kBlkLpLd;
#pragma unroll
                    for (int c = 0; c < 8; ++c) p[c] = rowp[c];
                }
                if (tid == 0) misc[1] = 0;
                if (warp == 0 && lane >= nw && lane < 8) skey[lane] = skey[8 + lane] = -2;
                // pivot key (high word of |re| + |im|, LAPACK izamax) and speculative reciprocal of column 0
                // (sign bit cleared: the canonical FP64 NaN of the GPU is negative and must still win the search, so that a
                // non-finite column is eliminated like any other and flagged at the end)
                int key = cand ? (__double2hiint(fabs(p[0].x) + fabs(p[0].y)) & 0x7fffffff) : -1;
                double2 rinv = crecip_fast(p[0]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int buf = (j & 1) * 8;
                    BLK_PH(8)
                    // every warp publishes its best candidate row; one barrier (of the panel warps) per column
                    const int wmax = __reduce_max_sync(kFull, key);
                    const unsigned bal = __ballot_sync(kFull, key == wmax);
                    const bool mine = lane == __ffs(bal) - 1;
                    if (mine) {
                        skey[buf + warp] = (wmax & ~7) | (7 - warp);  // ties go to the lower warp
                        spos[buf + warp] = pos;
                        double2* sl = slot + (buf + warp) * 8;
                        sl[0] = rinv;
#pragma unroll
                        for (int c = j + 1; c < 8; ++c) sl[c] = p[c];
                    }
                    BLK_PH(9)
                    asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
                    BLK_PH(10)
                    const int best = __reduce_max_sync(kFull, skey[buf + (lane & 7)]);
                    const int bw = 7 - (best & 7);
                    const double2* pr = slot + (buf + bw) * 8;
                    BLK_PH(11)
                    int nkey = -1;
                    double2 nrinv = rinv;
                    if (mine && warp == bw) {
                        if ((best & ~7) < 0x00100000) atomicOr(&misc[2], 2);  // zero (or subnormal) pivot
                        pos = k0 + j;
                        done = true;
                    } else if (have && !done) {
                        if (pos == k0 + j) pos = spos[buf + bw];
                        const double2 l = cmul(p[j], pr[0]);
                        p[j] = l;
                        if (j < 7) {
                            // the next column first: its search starts while the other columns are updated
                            cfnma(p[j + 1], l, pr[j + 1]);
                            nkey = cand ? (__double2hiint(fabs(p[j + 1].x) + fabs(p[j + 1].y)) & 0x7fffffff) : -1;
                            nrinv = crecip_fast(p[j + 1]);
                        }
#pragma unroll
                        for (int c = j + 2; c < 8; ++c) cfnma(p[c], l, pr[c]);
                    }
                    key = nkey;
                    rinv = nrinv;
                }
                if (have) {
                    if (done) {
                        const int r = pos - k0;
                        src[r] = start;
    #pragma unroll
                        for (int c = 0; c < 8; ++c) L11[r * 8 + c] = p[c];
                        if (A.coeffs) {
                            double2* rowp = M + (size_t)pos * ld + k0;
    #pragma unroll
                            for (int c = 0; c < 8; ++c) rowp[c] = p[c];  // U11 on and right of the diagonal
                        }
                    } else {
                        double2* lp = Lp + (pos - k0 - 8) * kBlkLpLd;
    #pragma unroll
                        for (int c = 0; c < 8; ++c) lp[c] = p[c];
                        if (start < k0 + 8) {
                            const int i = atomicAdd(&misc[1], 1);
                            dfrom[i] = start;
                            dto[i] = pos;
                        }
                    }
                }
            }
            BLK_PH(1)
            __syncthreads();
            BLK_PH(2)
            // ------------ phase U (+ the row moves of the trailing columns)
            const int wtr = PC - k0 - 8;
            if (tid < wtr) {
                const int col = k0 + 8 + tid;
                double2 av[8];
                const int nd = misc[1];
#pragma unroll
                for (int j = 0; j < 8; ++j) av[j] = M[(size_t)src[j] * ld + col];
                // displaced rows (they started inside the panel) go where the pivots came from
#pragma unroll
                for (int i0 = 0; i0 < 8; i0 += 4) {
                    if (i0 < nd) {  // CTA-uniform
                        double2 dv[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (i0 + i < nd) dv[i] = M[(size_t)dfrom[i0 + i] * ld + col];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (i0 + i < nd) M[(size_t)dto[i0 + i] * ld + col] = dv[i];
                    }
                }
#pragma unroll
                for (int kk = 1; kk < 8; ++kk)
#pragma unroll
                    for (int k2 = 0; k2 < kk; ++k2) cfnma(av[kk], L11[kk * 8 + k2], av[k2]);
#pragma unroll
                for (int j = 0; j < 8; ++j) Ub[j * G.ub_ld + tid] = av[j];
                if (A.coeffs) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) M[(size_t)(k0 + j) * ld + col] = av[j];
                }
            }
            __syncthreads();
            BLK_PH(3)
            // ------------ phase T: the next panel's columns now (-> Pn), the rest beside the next panel phase
            if (k0 + 8 < P) {
                blk_trailing_col0(M, ld, k0 + 8, (PR - k0 - 8) >> 3, reinterpret_cast<const double*>(Lp),
                                  reinterpret_cast<const double*>(Ub), G.ub_ld, Pn, warp, lane, NW);
                blk_trailing_simple(M, ld, k0 + 8, (PR - k0 - 8) >> 3, wtr >> 3, 1, reinterpret_cast<const double*>(Lp),
                             reinterpret_cast<const double*>(Ub), G.ub_ld, warp, lane, NW);
            } else
                blk_trailing_simple(M, ld, k0 + 8, (PR - k0 - 8) >> 3, wtr >> 3, 0, reinterpret_cast<const double*>(Lp),
                             reinterpret_cast<const double*>(Ub), G.ub_ld, warp, lane, NW);
            BLK_PH(4)
            __syncthreads();
            BLK_PH(5)
        }
        }  // !solved
        // ---------------- R = -corner / a^2
        for (int ij = tid; ij < nch * nch; ij += kBlkThreads) {
            const int i = ij / nch, j = ij - i * nch;
            // the symmetric path maintains the lower triangle only (matters when the border spans two tiles, nch > 8)
            const double2 cv = (solved && i < j) ? M[(size_t)(P + j) * ld + P + i] : M[(size_t)(P + i) * ld + P + j];
            small[ij] = make_double2(-cv.x / (a * a), -cv.y / (a * a));
        }
        __syncthreads();
        finish_system(A, sysi, a, misc[2], small, tid, kBlkThreads);
        if (A.coeffs) blk_coefficients(A, sysi, M, ld, P, xs, small + 3 * nch * nch);
        __syncthreads();
        // (discard.global.L2 of the dead slab lines was measured here: DRAM write-back per system did not drop -- 343 KB instead of
        // 225 KB -- and the rate fell by 2 %; the write-back happens while the system is being eliminated, not afterwards.)
        BLK_PH(6)
    }
}

// 0 -> *flag stays 0 when the n x n matrix F equals its transpose bit for bit
__global__ void symcheck_kernel(const double2* __restrict__ F_all, int n, int* __restrict__ flag) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const double2* F = F_all + (size_t)blockIdx.y * n * n;  // blockIdx.y: which of the free matrices
    const int i = idx / n, j = idx - i * n;
    if (j >= i) return;
    const double2 x = F[idx], y = F[(size_t)j * n + i];
    if (x.x != y.x || x.y != y.y) atomicOr(flag, 1);
}

static int blk_threads(const BlkGeom& g) { return g.PR <= 128 ? 128 : 256; }
static int blk_ctas_per_sm(const BlkGeom& g) { return 512 / blk_threads(g); }

template <int T>
static int launch_block_solve_t(const GenArgs& a, const BlkGeom& g, cudaStream_t st) {
    const size_t smem = blk_smem_bytes(g);
    if (cudaFuncSetAttribute(block_solve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(block_solve)");
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, block_solve_kernel<T>, T, smem) != cudaSuccess || per_sm < 1)
        return set_cuda_error("block_solve occupancy");
    if (per_sm > blk_ctas_per_sm(g)) per_sm = blk_ctas_per_sm(g);  // the workspace holds that many slabs per SM
    long long blocks = (long long)sm_count() * per_sm;
    if (blocks > a.batch) blocks = a.batch;  // (with a deferred-systems list: the list is at most the batch)
    block_solve_kernel<T><<<(unsigned)blocks, T, smem, st>>>(a, g);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("block_solve launch");
    return JITR_B200_OK;
}

// *flag (first int of the workspace) = 0 when the shared free matrix is symmetric
static int launch_symcheck(const GenArgs& a, int* flag, cudaStream_t st) {
    if (cudaMemsetAsync(flag, 0, sizeof(int), st) != cudaSuccess) return set_cuda_error("solve_batched symmetry flag");
    const int nn = a.n * a.n;
    symcheck_kernel<<<dim3((nn + 255) / 256, a.free_index ? a.n_free : 1), 256, 0, st>>>(a.free_matrix, a.n, flag);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("symcheck launch");
    return JITR_B200_OK;
}

static int launch_block_solve(const GenArgs& a_in, cudaStream_t st, const bool flag_ready = false) {
    GenArgs a = a_in;
    const BlkGeom g = blk_geom(a.n, a.nch);
    // symmetric path: one free matrix for the whole batch, checked here; the potentials are checked per system in the kernel
    a.symflag = nullptr;
    a.fallbacks = nullptr;
    if (option(JITR_B200_OPT_SYMMETRIC) && a.free_stride == 0) {
        int* flag = reinterpret_cast<int*>(a.workspace);
        if (!flag_ready) {
            const int rc = launch_symcheck(a, flag, st);
            if (rc != JITR_B200_OK) return rc;
        }
        a.symflag = flag;
        a.fallbacks = fallback_counter();
    }
    return blk_threads(g) == 128 ? launch_block_solve_t<128>(a, g, st) : launch_block_solve_t<256>(a, g, st);
}

// bytes of the slab part of the workspace (header + one slab per resident CTA)
static long long blk_slab_bytes(const BlkGeom& g, long long batch) {
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    long long ctas = (long long)sms * blk_ctas_per_sm(g);
    if (ctas > batch) ctas = batch;
    return (ctas * g.PR * g.PC + kBlkHeader) * 16;
}
static bool packed_path(int nbasis, int nch, int /*coeffs*/) {
    return nch == 1 && packed_solve_supported(nbasis) && option(JITR_B200_OPT_SYMMETRIC) != 0;
}

template <int MAXNS>
static int launch_warp_solve(const GenArgs& a, cudaStream_t st) {
    const int P = (a.N + 7) & ~7, msz = (P + 1) * (P + 1);
    int warps = 232448 / (msz * 16);
    if (warps > 8) warps = 8;
    if (warps < 1) return set_error(JITR_B200_ERR_UNSUPPORTED, "warp_solve: system does not fit in shared memory");
    const int smem = warps * msz * 16;
    if (cudaFuncSetAttribute(warp_solve_kernel<MAXNS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(warp_solve)");
    long long blocks = (a.batch + warps - 1) / warps;
    if (blocks > sm_count()) blocks = sm_count();
    warp_solve_kernel<MAXNS><<<(unsigned)blocks, warps * 32, smem, st>>>(a);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("warp_solve launch");
    return JITR_B200_OK;
}

}  // namespace jitr

using namespace jitr;

#ifdef JITR_BLK_PROFILE
extern "C" int jitr_b200_debug_block_phases(unsigned long long* out16, int reset) {
    if (out16 && cudaMemcpyFromSymbol(out16, g_blk_phase, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(g_blk_phase, z, sizeof(z)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

extern "C" long long jitr_b200_solve_workspace_bytes(int nbasis, int nch, long long batch, int coeffs) {
    if (nbasis < 1 || nch < 1 || nch > kMaxCh || batch < 0) return -1;
    const long long n = (long long)nbasis * nch;
    const BlkGeom g = blk_geom((int)n, nch);
    // packed symmetric path: the slabs of the partial-pivoting kernel (for the systems it defers) + the deferred list
    if (packed_path(nbasis, nch, coeffs)) return blk_slab_bytes(g, batch) + 16 + ((batch * 4 + 15) & ~15LL);
    if (nch == 1 && !coeffs && ((nbasis + 7) & ~7) <= JITR_WARP_SOLVE_MAXP) return 0;  // warp-per-system: shared memory only
    if (g.PR <= kBlkMaxRows) return blk_slab_bytes(g, batch);  // one slab per resident CTA
    const long long in_smem = (n * (n + nch) + 2 * n + nch) * 16 + 20000;
    return in_smem > 232448 ? batch * n * (n + nch) * 16 : 0;
}

static int solve_batched_impl(int nbasis, int nch, long long batch, const double* free_matrix, long long free_stride,
                              const int* free_index, int n_free, const double* local, const double* nonlocal_,
                              const double* interaction, const double* wsqrt, const double* sys, int sys_stride,
                              const double* bvec, const double* asym, long long asym_stride, const double* weights, double* R,
                              double* S, double* uext, double* coeffs, double* workspace, int* status, void* stream) {
    if (batch == 0) return JITR_B200_OK;  // empty batch (its tensors have null data pointers)
    if (nbasis < 1 || nch < 1 || nch > kMaxCh || batch < 0 || !free_matrix || !sys || !bvec || !asym || !weights || !R ||
        !S || !uext || !status)
        return set_error(JITR_B200_ERR_BAD_ARG, "solve_batched: null pointer or bad size (nch <= 16)");
    if (nonlocal_ && !wsqrt) return set_error(JITR_B200_ERR_BAD_ARG, "solve_batched: nonlocal potential needs wsqrt");
    if (batch == 0) return JITR_B200_OK;
    GenArgs a{};
    a.N = nbasis; a.nch = nch; a.n = nbasis * nch; a.batch = batch;
    a.free_matrix = reinterpret_cast<const double2*>(free_matrix); a.free_stride = free_stride;
    a.free_index = free_index; a.n_free = n_free;
    a.local = reinterpret_cast<const double2*>(local);
    a.nonlocal_ = reinterpret_cast<const double2*>(nonlocal_);
    a.interaction = reinterpret_cast<const double2*>(interaction);
    a.wsqrt = wsqrt; a.sys = sys; a.sys_stride = sys_stride; a.bvec = bvec;
    a.asym = reinterpret_cast<const double2*>(asym); a.asym_stride = asym_stride; a.weights = weights;
    a.R = reinterpret_cast<double2*>(R); a.S = reinterpret_cast<double2*>(S);
    a.uext = reinterpret_cast<double2*>(uext); a.coeffs = reinterpret_cast<double2*>(coeffs);
    a.workspace = reinterpret_cast<double2*>(workspace); a.status = status;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // single channel, no coefficients, one shared free matrix, P <= 64: packed symmetric LDL^T, one warp per system, the
    // nonlocal kernel staged by bulk-async copies; what it defers goes to the partial-pivoting CTA-per-system kernel
    if (packed_path(nbasis, nch, coeffs != nullptr) && free_stride == 0 && workspace) {
        const BlkGeom g = blk_geom(a.n, nch);
        int* flag = reinterpret_cast<int*>(workspace);
        int* defer_count = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + blk_slab_bytes(g, batch));
        int* defer_list = defer_count + 4;
        int rc = launch_symcheck(a, flag, st);
        if (rc != JITR_B200_OK) return rc;
        if (cudaMemsetAsync(defer_count, 0, 16, st) != cudaSuccess) return set_cuda_error("solve_batched deferred-systems counter");
        a.fallbacks = fallback_counter();
        rc = launch_packed_solve(a, flag, option(JITR_B200_OPT_ASSUME_SYMMETRIC_KERNEL) ? 0 : 1, defer_list, defer_count, st);
        if (rc != JITR_B200_OK) return rc;
        a.sys_list = defer_list;
        a.sys_count = defer_count;
        return launch_block_solve(a, st, true);
    }
    // single channel, no coefficients, small P: warp-per-system blocked LU
    if (nch == 1 && !coeffs && ((nbasis + 7) & ~7) <= JITR_WARP_SOLVE_MAXP) {
        static_assert(JITR_WARP_SOLVE_MAXP + 1 <= 64, "warp_solve_kernel<2>: at most two row slots per lane");
        return launch_warp_solve<2>(a, st);
    }
    if (blk_geom(a.n, nch).PR <= kBlkMaxRows) {
        if (!workspace) return set_error(JITR_B200_ERR_BAD_ARG, "solve_batched: workspace of jitr_b200_solve_workspace_bytes() bytes required");
        return launch_block_solve(a, st);
    }
    const int n = a.n, ld = n + nch;
    const size_t scratch = (size_t)(n + ld) * 16 + 32 * 8 + 36 * 4 + (size_t)4 * kMaxCh * kMaxCh * 16;
    const size_t scratch_al = (scratch + 15) & ~(size_t)15;
    const size_t in_smem = scratch_al + (size_t)n * ld * 16;
    const bool use_smem = in_smem <= 232448;
    if (!use_smem && !workspace) return set_error(JITR_B200_ERR_BAD_ARG, "solve_batched: system too large for shared memory and no workspace given");
    const size_t smem = use_smem ? in_smem : scratch_al;
    const int sms = sm_count();
    int per_sm = use_smem ? (int)(232448 / (smem + 1024)) : 4;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long blocks = batch < (long long)sms * per_sm ? batch : (long long)sms * per_sm;
    if (use_smem) {
        if (cudaFuncSetAttribute(general_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return set_cuda_error("cudaFuncSetAttribute(general_solve<smem>)");
        general_solve_kernel<true><<<(unsigned)blocks, kGenThreads, smem, st>>>(a);
    } else {
        if (cudaFuncSetAttribute(general_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return set_cuda_error("cudaFuncSetAttribute(general_solve<global>)");
        general_solve_kernel<false><<<(unsigned)blocks, kGenThreads, smem, st>>>(a);
    }
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("general_solve launch");
    return JITR_B200_OK;
}

extern "C" int jitr_b200_solve_batched(int nbasis, int nch, long long batch, const double* free_matrix,
                                       long long free_stride, const double* local, const double* nonlocal_,
                                       const double* interaction, const double* wsqrt, const double* sys,
                                       int sys_stride, const double* bvec, const double* asym, long long asym_stride,
                                       const double* weights, double* R, double* S, double* uext, double* coeffs,
                                       double* workspace, int* status, void* stream) {
    return solve_batched_impl(nbasis, nch, batch, free_matrix, free_stride, nullptr, 1, local, nonlocal_, interaction, wsqrt, sys,
                              sys_stride, bvec, asym, asym_stride, weights, R, S, uext, coeffs, workspace, status, stream);
}

extern "C" int jitr_b200_solve_batched_indexed(int nbasis, int nch, long long batch, const double* free_matrices, int n_free,
                                               const int* free_index, const double* local, const double* nonlocal_,
                                               const double* interaction, const double* wsqrt, const double* sys,
                                               int sys_stride, const double* bvec, const double* asym, long long asym_stride,
                                               const double* weights, double* R, double* S, double* uext, double* coeffs,
                                               double* workspace, int* status, void* stream) {
    if (batch > 0 && (n_free < 1 || !free_index))
        return set_error(JITR_B200_ERR_BAD_ARG, "solve_batched_indexed: free_index and n_free >= 1 required");
    return solve_batched_impl(nbasis, nch, batch, free_matrices, 0, free_index, n_free, local, nonlocal_, interaction, wsqrt, sys,
                              sys_stride, bvec, asym, asym_stride, weights, R, S, uext, coeffs, workspace, status, stream);
}
