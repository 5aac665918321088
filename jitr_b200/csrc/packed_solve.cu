// Single-channel batched R-matrix solve on a packed lower triangle in shared memory (BASELINE config 4: N = 60, one full
// N x N complex nonlocal kernel per system).  Replaces, for complex-symmetric systems, the L2-slab CTA-per-system kernel of
// general.cu:  rmatrix/rmatrix.py:126-246 (Solver.interaction_matrix / Solver.solve), quadrature/kernel.py:197-214
// (nonlocal weighting), rmatrix/core.py:7-60.
//
// One WARP per system, PackedGeom<P>::WARP_ELEMS * 16 B of shared memory each (38 KB at P = 64: six systems per SM):
//   1. the lower triangle of the system's nonlocal kernel (or interaction matrix) is fetched row by row with bulk-async
//      copies (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes: row r = (r + 1) * 16 bytes straight into
//      its packed position), completion on the warp's mbarrier -- the other warps of the SM keep eliminating meanwhile, so
//      the HBM stream (half of N^2 x 16 B per system) hides behind their FP64 work without a second buffer;
//   2. one pass forms  A = F + sqrt(w_n w_m) V (a / k0) / k0 / E0 (+ diag V_loc / E0)  in place and, unless the caller
//      vouches for it, checks V[n][m] == V[m][n] against the upper triangle read straight from global memory;
//   3. threshold-pivoted LDL^T in place (lu_packed.cuh), R = b^T A^-1 b / a^2, S, u'_ext.
// Systems that are not symmetric or fail the threshold test are appended to a list and solved by the partial-pivoting
// kernel of general.cu afterwards (same launch sequence, no host round trip).
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"
#include "general_decl.cuh"
#include "jitr_b200.h"
#include "lu_packed.cuh"

namespace jitr {

namespace {

constexpr int kPackedMaxSmem = 232448;

// FP ("first panel" mode, P > 32, no coefficients): the first eight columns are eliminated in registers and only the Schur
// complement of order P - 8 is stored -- seven systems per SM at P = 64 instead of six, and no staging pass over the region
template <int P, bool FP>
struct PackedPlan {
    using G = PackedGeom<FP ? P - 8 : P>;
    static constexpr int kTail = 2 * P * 8 + 16 * 8;  // b, wsqrt (doubles) + up to 16 mbarriers
    static constexpr int kFit = (kPackedMaxSmem - kTail) / (G::WARP_ELEMS * 16);
    // first-panel mode keeps all B fragments of the panel in registers: at most eight warps, 255 registers per thread
    static constexpr int kCap = FP ? 8 : 16;
    static constexpr int kWarps = kFit > kCap ? kCap : kFit;
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kSmem = kWarps * G::WARP_ELEMS * 16 + kTail;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!done);
}

// A = F + (scaled V | interaction) (+ V_loc / E0 on the diagonal), in place on the packed lower triangle whose V part the bulk
// copies have delivered; identity padding beyond N; optionally the bit-wise comparison of V with its transpose read from
// global memory.  One block row (8 rows) per step: left of the diagonal tile a lane owns column c = lane (+32) and has up to
// sixteen independent global loads in flight (rows of F coalesced; the upper triangle of V is walked by columns, eight
// consecutive elements = one 128-byte line per lane); the diagonal tile is spread over all 32 lanes.  Returns the OR of the
// XORs of the compared words (0: symmetric).
template <int P, bool HAS_V, bool INTER, bool CHECK>
__device__ __forceinline__ unsigned long long stage_system(double2* __restrict__ M, const double2* __restrict__ F,
                                                           const double2* __restrict__ Vsrc, const double2* __restrict__ Vloc,
                                                           const double* __restrict__ ws, const double nl_scale, const double E0,
                                                           const int N, const int lane) {
    using G = PackedGeom<P>;
    const double2 zero2 = make_double2(0.0, 0.0);
    unsigned long long diff = 0ULL;
    double2 vl[2];  // local potential / E0 of the rows this lane "owns" (row = lane + 32 h)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        vl[h] = zero2;
        if (Vloc && c < N) {
            const double2 w = Vloc[c];
            vl[h] = make_double2(w.x / E0, w.y / E0);
        }
    }
#pragma unroll 1
    for (int tr = 0; tr < G::NB; ++tr) {
        const int r0 = 8 * tr;
        const int nrow = min(8, N - r0);  // rows of this block row inside the system (<= 0: all padding); warp-uniform
        double2* Mb = M + G::off(tr);
        const int ldt = G::ld(tr);
        // ---- tiles left of the diagonal tile
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            if (c < r0) {
                const double2* Fp = F + (size_t)r0 * N + c;
                const double2* Vp = HAS_V ? Vsrc + (size_t)c * N + r0 : nullptr;
                double2 f[8], t[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    f[i] = i < nrow ? Fp[(size_t)i * N] : zero2;
                    if (HAS_V && CHECK) t[i] = i < nrow ? Vp[i] : zero2;
                }
                const double wc = HAS_V && !INTER ? ws[c] * nl_scale : 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    double2* e = Mb + i * ldt + c;
                    double2 v = f[i];
                    if (HAS_V && i < nrow) {
                        const double2 w = *e;
                        if (CHECK) diff |= (unsigned long long)(__double_as_longlong(w.x) ^ __double_as_longlong(t[i].x)) |
                                           (unsigned long long)(__double_as_longlong(w.y) ^ __double_as_longlong(t[i].y));
                        if (INTER) {
                            v.x += w.x; v.y += w.y;
                        } else {
                            const double sc = wc * ws[r0 + i];
                            v.x = fma(w.x, sc, v.x); v.y = fma(w.y, sc, v.y);
                        }
                    }
                    *e = v;
                }
            }
        }
        // ---- the diagonal tile: element (r0 + i, r0 + j), j <= i, i = (lane >> 3) + 4 k, j = lane & 7
        const int hrow = r0 >> 5;  // the rows r0 .. r0 + 7 are owned by slot hrow of the lanes (r & 31)
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = (lane >> 3) + 4 * k, j = lane & 7;
            const int r = r0 + i, c = r0 + j;
            const double2 vd = shfl2(hrow ? vl[1] : vl[0], r & 31);
            if (j <= i) {
                double2* e = Mb + i * ldt + c;
                double2 v = make_double2(j == i ? 1.0 : 0.0, 0.0);  // identity padding
                if (r < N) {
                    v = F[(size_t)r * N + c];
                    if (HAS_V) {
                        const double2 w = *e;
                        if (CHECK && j < i) {
                            const double2 t = Vsrc[(size_t)c * N + r];
                            diff |= (unsigned long long)(__double_as_longlong(w.x) ^ __double_as_longlong(t.x)) |
                                    (unsigned long long)(__double_as_longlong(w.y) ^ __double_as_longlong(t.y));
                        }
                        if (INTER) {
                            v.x += w.x; v.y += w.y;
                        } else {
                            const double sc = ws[c] * nl_scale * ws[r];
                            v.x = fma(w.x, sc, v.x); v.y = fma(w.y, sc, v.y);
                        }
                    }
                    if (j == i) {
                        v.x += vd.x; v.y += vd.y;
                    }
                }
                *e = v;
            }
        }
    }
    return diff;
}

// First-panel mode: A = F + (scaled V | interaction) (+ V_loc / E0) never exists as a whole.  The eight pivot columns of V arrive by
// bulk-async copies (one 128-byte row segment per row, completion on the warp's mbarrier) into the scratch the panel is about to
// use, are combined with F's columns in registers and eliminated there; the Schur complement is accumulated on DMMA straight from
// global memory -- every accumulator starts from its own entries of F and V (four consecutive lanes read 64 contiguous bytes of a row)
// -- and only its lower triangle of order P - 8 is written, into the packed region.  With CHECK every entry of V that is read is
// compared bit for bit with its transpose.  Returns the OR of the XORs (0: symmetric); cacc / fail as packed_first_panel.
template <int PT, bool HAS_V, bool INTER, bool CHECK>
__device__ __forceinline__ unsigned long long first_panel_global(double2* __restrict__ W, const double2* __restrict__ F,
                                                                 const double2* __restrict__ Vsrc, const double2* __restrict__ Vloc,
                                                                 const double* __restrict__ ws, const double* __restrict__ bs,
                                                                 const double nl_scale, const double E0, const int N, const int lane,
                                                                 const unsigned bar, unsigned& parity, double2& cacc, int& fail) {
    using G = PackedGeom<PT - 8>;
    constexpr int n = PT - 8, NT = n / 8, LB_LD = 9;
    constexpr int WB_OFF = 0, LB_OFF = G::SIZE - n * LB_LD;
    static_assert(PT > 32 && PT <= 64 && LB_OFF >= PT * LB_LD, "first-panel mode: two row slots, scratch inside the region");
    const double2 zero2 = make_double2(0.0, 0.0);
    unsigned long long diff = 0ULL;
    auto xorw = [](const double2 a, const double2 b) {
        return (unsigned long long)(__double_as_longlong(a.x) ^ __double_as_longlong(b.x)) |
               (unsigned long long)(__double_as_longlong(a.y) ^ __double_as_longlong(b.y));
    };
    if (HAS_V) {
        // ---- the panel columns of V: row i, columns 0..7 = 128 contiguous bytes -> W[i * 9 .. i * 9 + 7]
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (lane == 0) {
            const unsigned bytes = (unsigned)N * 8u * 16u;
            asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
        }
        __syncwarp();
        for (int r = lane; r < N; r += 32) {
            const unsigned dst = smem_u32(W + WB_OFF + r * LB_LD);
            const unsigned long long src = (unsigned long long)(Vsrc + (size_t)r * N);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                         "r"(128u), "r"(bar)
                         : "memory");
        }
    }
    double2 p[2][9], d[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int i = 32 * s + lane;
        const bool in = i < N;
#pragma unroll
        for (int j = 0; j < 8; ++j) p[s][j] = (in && j < N) ? F[(size_t)i * N + j] : make_double2(i == j ? 1.0 : 0.0, 0.0);
        p[s][8] = make_double2(i < PT ? bs[i] : 0.0, 0.0);
        d[s] = make_double2(1.0, 0.0);
        if (in) {
            d[s] = F[(size_t)i * N + i];
            if (Vloc) {
                const double2 w = Vloc[i];
                d[s].x += w.x / E0; d[s].y += w.y / E0;
            }
        }
    }
    if (HAS_V) {
        double2 vt[2][8];
        if (CHECK) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int i = 32 * s + lane;
#pragma unroll
                for (int j = 0; j < 8; ++j) vt[s][j] = (i < N && j < N) ? Vsrc[(size_t)j * N + i] : zero2;  // V[j][i]: coalesced across the lanes
            }
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int i = 32 * s + lane;
            if (i < N) {
                const double wi = INTER ? 1.0 : ws[i] * nl_scale;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < N) {
                        const double2 v = W[WB_OFF + i * LB_LD + j];
                        if (CHECK) diff |= xorw(v, vt[s][j]);
                        const double sc = INTER ? 1.0 : wi * ws[j];
                        p[s][j].x = fma(v.x, sc, p[s][j].x);
                        p[s][j].y = fma(v.y, sc, p[s][j].y);
                    }
                }
            }
        }
        // the diagonal entries of V (rows >= 8: their column is not in the panel)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int i = 32 * s + lane;
            if (i < N) {
                const double2 v = i < 8 ? zero2 : Vsrc[(size_t)i * N + i];
                const double sc = INTER ? 1.0 : ws[i] * nl_scale * ws[i];
                if (i >= 8) {
                    d[s].x = fma(v.x, sc, d[s].x);
                    d[s].y = fma(v.y, sc, d[s].y);
                }
            }
        }
        __syncwarp();
    }
    // local potential on the diagonal of the panel block (rows < 8: their diagonal entry is p[0][lane])
    if (Vloc && lane < 8 && lane < N) {
        const double2 w = Vloc[lane];
        p[0][lane].x += w.x / E0;
        p[0][lane].y += w.y / E0;
    }
    sym_panel_eliminate<2>(
        p, PT, lane, cacc, fail,
        [&](int s, int kk, double2 v) {
            const int i = 32 * s + lane;
            if (i > kk && i < PT) W[WB_OFF + i * LB_LD + kk] = v;
        },
        [&](int kk, int j) { return W[WB_OFF + j * LB_LD + kk]; });
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
    // ---- Schur complement, lower tiles only; accumulators initialised from F and V in global memory
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
        const int row = 8 + 8 * tr + rho;
        const bool rin = row < N;
        const double wr = (HAS_V && !INTER && rin) ? ws[row] * nl_scale : 1.0;
        const double2 dv = shfl2(row >= 32 ? d[1] : d[0], row & 31);  // the assembled diagonal entry of this lane's accumulator row
        double2 c[NT][2];
#pragma unroll
        for (int q = 0; q <= tr; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = 8 + 8 * q + 4 * h + fk;
                double2 v = make_double2(row == col ? 1.0 : 0.0, 0.0);  // identity padding
                if (rin && col < N) {
                    if (row == col) {
                        v = dv;
                    } else {
                        v = F[(size_t)row * N + col];
                        if (HAS_V) {
                            const double2 w = Vsrc[(size_t)row * N + col];
                            if (CHECK) diff |= xorw(w, Vsrc[(size_t)col * N + row]);
                            const double sc = INTER ? 1.0 : wr * ws[col];
                            v.x = fma(w.x, sc, v.x);
                            v.y = fma(w.y, sc, v.y);
                        }
                    }
                }
                c[q][h] = v;
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
    return diff;
}

template <int P, bool FP>
__global__ void __launch_bounds__(PackedPlan<P, FP>::kThreads, 1)
packed_solve_kernel(const GenArgs A, const int* __restrict__ symflag, const int check_symmetry,
                    unsigned long long* __restrict__ work_counter, int* __restrict__ defer_list, int* __restrict__ defer_count) {
    using Plan = PackedPlan<P, FP>;
    using G = typename Plan::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = A.N;
    const int warp = __shfl_sync(kFull, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    double2* M = reinterpret_cast<double2*>(smem_raw) + (size_t)warp * G::WARP_ELEMS;
    double* bs = reinterpret_cast<double*>(smem_raw + (size_t)Plan::kWarps * G::WARP_ELEMS * 16);
    double* ws = bs + P;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(ws + P);
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        bs[i] = i < N ? A.bvec[i] : 0.0;
        ws[i] = (i < N && A.wsqrt) ? A.wsqrt[i] : 0.0;
    }
    const unsigned bar = smem_u32(bars + warp);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const bool free_ok = *symflag == 0;  // the shared free matrix is symmetric
    const double2* Vsrc_base = A.interaction ? A.interaction : A.nonlocal_;
    const double2 zero2 = make_double2(0.0, 0.0);
    unsigned parity = 0;
    // (A cp.async.bulk.prefetch.L2 of the next system's matrix one system ahead was measured and removed: ncu showed 97 KB of DRAM
    // reads per system instead of the algorithmic 57.6 KB -- the prefetched lines were fetched again -- and 7 % less throughput.)
    auto grab = [&]() -> long long {
        long long v = 0;
        if (lane == 0) v = (long long)atomicAdd(work_counter, 1ULL);
        return __shfl_sync(kFull, v, 0);
    };
    while (true) {
        const long long sysi = grab();
        if (sysi >= A.batch) break;
        bool defer = !free_ok;
        const double* sy = A.sys + (size_t)sysi * A.sys_stride;
        const double a = sy[0], E0 = sy[1], k0 = sy[2];
        const double nl_scale = A.interaction ? 1.0 : (a / k0) / k0 / E0;  // kernel.py:202 (x radius), rmatrix.py:175-178 (/k0, /E0)
        const double2* Vsrc = Vsrc_base ? Vsrc_base + (size_t)sysi * N * N : nullptr;
        double2 corner = zero2;
        if constexpr (FP) {
            if (!defer) {
                const double2* F = A.free_matrix + (A.free_index ? (size_t)A.free_index[sysi] * N * N : 0);
                const double2* Vloc = (!A.interaction && A.local) ? A.local + (size_t)sysi * N : nullptr;
                unsigned long long diff = 0ULL;
                int fail = 0, bad = 0;
                if (!Vsrc)
                    diff = first_panel_global<P, false, false, false>(M, F, nullptr, Vloc, ws, bs, nl_scale, E0, N, lane, bar, parity, corner, fail);
                else if (A.interaction)
                    diff = check_symmetry
                               ? first_panel_global<P, true, true, true>(M, F, Vsrc, Vloc, ws, bs, nl_scale, E0, N, lane, bar, parity, corner, fail)
                               : first_panel_global<P, true, true, false>(M, F, Vsrc, Vloc, ws, bs, nl_scale, E0, N, lane, bar, parity, corner, fail);
                else
                    diff = check_symmetry
                               ? first_panel_global<P, true, false, true>(M, F, Vsrc, Vloc, ws, bs, nl_scale, E0, N, lane, bar, parity, corner, fail)
                               : first_panel_global<P, true, false, false>(M, F, Vsrc, Vloc, ws, bs, nl_scale, E0, N, lane, bar, parity, corner, fail);
                defer = __any_sync(kFull, diff != 0ULL);
                if (!defer) {
                    corner = packed_ldlt<P - 8>(M, lane, bad, corner);
                    bad |= fail;
                    defer = bad != 0;
                    if (defer && lane == 0 && A.fallbacks) atomicAdd(A.fallbacks, 1ULL);
                }
            }
        } else {
        if (!defer && Vsrc) {
            // ---- stage the lower triangle: one bulk-async copy per row, all of them in flight at once
            __syncwarp();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the warp's earlier accesses to its region come first
            if (lane == 0) {
                const unsigned bytes = (unsigned)(N * (N + 1) / 2) * 16u;
                asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
            }
            __syncwarp();
            for (int r = lane; r < N; r += 32) {
                const unsigned dst = smem_u32(M + G::row(r));
                const unsigned long long src = (unsigned long long)(Vsrc + (size_t)r * N);
                const unsigned bytes = (unsigned)(r + 1) * 16u;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                             "r"(bytes), "r"(bar)
                             : "memory");
            }
#ifndef JITR_PACKED_NO_UPPER_PREFETCH
            // the symmetry check reads the UPPER triangle from global memory block row by block row (stage_system): ask the L2 for
            // those lines now, while the bulk copies of the lower triangle are in flight -- the check then waits for L2 hits instead
            // of eight DRAM round trips in sequence.  Every line is still read from DRAM once (this system's data, used within
            // microseconds; unlike a prefetch of the NEXT system, which doubled the DRAM reads and was removed).
            if (check_symmetry) {
                for (int r = lane; r < N - 1; r += 32) {
                    const char* p0 = reinterpret_cast<const char*>(Vsrc + (size_t)r * N + r + 1);
                    const char* p1 = reinterpret_cast<const char*>(Vsrc + (size_t)r * N + N) - 1;
                    for (const char* q = p0; q <= p1; q += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p1));
                }
            }
#endif
            mbar_wait(bar, parity);
            parity ^= 1u;
        }
        if (!defer) {
            const double2* F = A.free_matrix + (A.free_index ? (size_t)A.free_index[sysi] * N * N : 0);
            const double2* Vloc = (!A.interaction && A.local) ? A.local + (size_t)sysi * N : nullptr;
            unsigned long long diff = 0ULL;
            if (!Vsrc) diff = stage_system<P, false, false, false>(M, F, nullptr, Vloc, ws, nl_scale, E0, N, lane);
            else if (A.interaction)
                diff = check_symmetry ? stage_system<P, true, true, true>(M, F, Vsrc, Vloc, ws, nl_scale, E0, N, lane)
                                      : stage_system<P, true, true, false>(M, F, Vsrc, Vloc, ws, nl_scale, E0, N, lane);
            else
                diff = check_symmetry ? stage_system<P, true, false, true>(M, F, Vsrc, Vloc, ws, nl_scale, E0, N, lane)
                                      : stage_system<P, true, false, false>(M, F, Vsrc, Vloc, ws, nl_scale, E0, N, lane);
            const bool asym = diff != 0ULL;
            for (int r = lane; r < P; r += 32) M[G::rhs(r)] = make_double2(bs[r], 0.0);
            defer = __any_sync(kFull, asym);
            __syncwarp();
        }
        if (!defer) {
            int bad = 0;
            corner = packed_ldlt<P>(M, lane, bad);
            defer = bad != 0;
            if (defer && lane == 0 && A.fallbacks) atomicAdd(A.fallbacks, 1ULL);
        }
        }
        if (defer) {
            if (lane == 0) defer_list[atomicAdd(defer_count, 1)] = (int)sysi;
            continue;
        }
        double2 ue = zero2;
        {
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
            ue = make_double2(-0.5 * v.y, 0.5 * v.x);
            if (lane == 0) {
                A.R[sysi] = R;
                A.S[sysi] = S;
                A.uext[sysi] = ue;
                const bool nonfinite = !(isfinite(S.x) && isfinite(S.y));
                A.status[sysi] = nonfinite ? JITR_B200_STATUS_NONFINITE : JITR_B200_STATUS_OK;
            }
        }
        if (A.coeffs) {
            // x = A^-1 (b u'_ext) = u'_ext A^-1 b   (core.py:81-82, one channel)
            double2 z[2];
            packed_back_substitute<P>(M, lane, z);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = lane + 32 * h;
                if (i < N) A.coeffs[(size_t)sysi * N + i] = cmul(ue, z[h]);
            }
        }
        __syncwarp();
    }
}

template <int P, bool FP>
int launch_packed_fp(const GenArgs& a, const int* symflag, int check_symmetry, int* defer_list, int* defer_count, cudaStream_t st) {
    using Plan = PackedPlan<P, FP>;
    auto kern = packed_solve_kernel<P, FP>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan::kSmem) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(packed_solve)");
    long long blocks = (a.batch + Plan::kWarps - 1) / Plan::kWarps;
    if (blocks > sm_count()) blocks = sm_count();
    unsigned long long* wc = next_work_counter(st);
    if (!wc) return set_cuda_error("packed_solve work counter");
    kern<<<(unsigned)blocks, Plan::kThreads, Plan::kSmem, st>>>(a, symflag, check_symmetry, wc, defer_list, defer_count);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("packed_solve launch");
    return JITR_B200_OK;
}

// first-panel mode whenever it exists (P > 32) and no coefficients are wanted (their back substitution needs all columns)
template <int P>
int launch_packed(const GenArgs& a, const int* symflag, int check_symmetry, int* defer_list, int* defer_count, cudaStream_t st) {
    if constexpr (P > 32) {
        if (!a.coeffs && option(JITR_B200_OPT_PACKED_FIRST_PANEL))
            return launch_packed_fp<P, true>(a, symflag, check_symmetry, defer_list, defer_count, st);
    }
    return launch_packed_fp<P, false>(a, symflag, check_symmetry, defer_list, defer_count, st);
}

}  // namespace

bool packed_solve_supported(int nbasis) { return nbasis >= 1 && nbasis <= 64; }

int launch_packed_solve(const GenArgs& a, const int* symflag, int check_symmetry, int* defer_list, int* defer_count, cudaStream_t st) {
    const int P = (a.N + 7) & ~7;
    switch (P) {
        case 8: return launch_packed<8>(a, symflag, check_symmetry, defer_list, defer_count, st);
        case 16: return launch_packed<16>(a, symflag, check_symmetry, defer_list, defer_count, st);
        case 24: return launch_packed<24>(a, symflag, check_symmetry, defer_list, defer_count, st);
        case 32: return launch_packed<32>(a, symflag, check_symmetry, defer_list, defer_count, st);
        case 40: return launch_packed<40>(a, symflag, check_symmetry, defer_list, defer_count, st);
        case 48: return launch_packed<48>(a, symflag, check_symmetry, defer_list, defer_count, st);
        case 56: return launch_packed<56>(a, symflag, check_symmetry, defer_list, defer_count, st);
        case 64: return launch_packed<64>(a, symflag, check_symmetry, defer_list, defer_count, st);
        default: return set_error(JITR_B200_ERR_UNSUPPORTED, "packed_solve: nbasis must be in 1..64");
    }
}

}  // namespace jitr
