// Fused spin-1/2 elastic sweep kernel:
// potential evaluation -> system assembly -> bordered LU -> R -> S, one WARP per (sample, energy)
// pair walking l = 0..lmax, j = l +- 1/2 with the reference's data-dependent early break
// (xs/elastic.py:104-183).  nbasis is a runtime value (<= 64); the kernel is instantiated per solver MODE
// (fast paths for P = 8 ceil(N/8) <= 40, generic paths above), one translation unit per MODE.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"
#include "jitr_b200.h"
#include "lu_fast.cuh"
#include "lu_packed.cuh"
#include "lu_sym.cuh"
#include "lu_warp.cuh"
#include "potentials.cuh"
#include "sweep_decl.cuh"

namespace jitr {

constexpr int kSweepMaxSmem = 232448 - 64;  // max dynamic shared memory per block on sm_100, less the static part

// ---------------------------------------------------------------------------------------------
// Per-warp solver of one system  a^2 A = That + diag(d):  returns -b^T (a^2 A)^-1 b.
// MODE > 0: fast path for P = MODE (lu_fast.cuh, the system never exists as a whole; 12 warps per
//           SM for P = 40).   MODE = -2 / -3: generic path, the bordered matrix is assembled in the
//           warp's shared-memory region (lu_warp.cuh) with at most 2 / 3 row slots per lane.
// ---------------------------------------------------------------------------------------------
template <int MODE>
struct WarpSolver;

template <int MODE>
struct SweepPlan;

template <int MAXNS>
struct GenericPlan {
    static constexpr int kMaxWarps = 8;
    static constexpr int kThreads = kMaxWarps * 32;
    static bool supports(int N) { return N >= 1 && ((N + 7) & ~7) + 1 <= 32 * MAXNS && N <= 64; }
    static int warps(int N) {
        const int P = (N + 7) & ~7, msz = (P + 1) * (P + 1);
        const int fit = (kSweepMaxSmem - N * N * 8 - N * 8) / (msz * 16);
        return fit > kMaxWarps ? kMaxWarps : fit;
    }
    static int smem(int N, int w) {
        const int P = (N + 7) & ~7, msz = (P + 1) * (P + 1);
        return w * msz * 16 + N * N * 8 + N * 8;
    }
};
template <>
struct SweepPlan<-2> : GenericPlan<2> {};
template <>
struct SweepPlan<-3> : GenericPlan<3> {};

template <int MAXNS>
struct GenericSolver {
    double2* M;
    const double* Ts;
    const double* bs;
    LuGeom g;
    int N, lane;
    const double* mx;
    __device__ __forceinline__ void centrifugal(double& c0, double& c1) const {
        const double x0 = mx[lane < N ? lane : 0], x1 = mx[lane + 32 < N ? lane + 32 : 0];
        c0 = 1.0 / (x0 * x0);
        c1 = 1.0 / (x1 * x1);
    }
    __device__ __forceinline__ void init(unsigned char* smem, const double* That, const double* bvec,
                                         const double* mesh_x, int N_, int nwarps, int warp, int lane_) {
        N = N_;
        lane = lane_;
        mx = mesh_x;
        g.N = N;
        g.P = (N + 7) & ~7;
        g.LD = g.P + 1;
        const int msz = g.LD * g.LD;
        M = reinterpret_cast<double2*>(smem) + (size_t)warp * msz;
        double* T = reinterpret_cast<double*>(smem + (size_t)nwarps * msz * 16);
        double* b = T + N * N;
        for (int i = threadIdx.x; i < N * N; i += blockDim.x) T[i] = That[i];
        for (int i = threadIdx.x; i < N; i += blockDim.x) b[i] = bvec[i];
        Ts = T;
        bs = b;
    }
    // [[That + diag, 0, b], [0, I, 0], [b^T, 0, 0]] in the warp's region
    __device__ __forceinline__ void assemble(const double2 d0, const double2 d1) {
        const int P = g.P, LD = g.LD;
        const int n0 = lane, n1 = lane + 32;
        const bool v0 = n0 < N, v1 = n1 < N;
        const double2 zero2 = make_double2(0.0, 0.0);
        for (int i = 0; i < N; ++i) {
            const double* Trow = Ts + i * N;
            double2* Mrow = M + i * LD;
            if (lane < P) Mrow[lane] = make_double2(v0 ? Trow[n0] : 0.0, 0.0);
            if (n1 < P) Mrow[n1] = make_double2(v1 ? Trow[n1] : 0.0, 0.0);
        }
        for (int i = N; i < P; ++i)  // identity padding rows
            for (int j = lane; j < LD; j += 32) M[i * LD + j] = make_double2(i == j ? 1.0 : 0.0, 0.0);
        __syncwarp();
        if (v0) {
            const double2 bb = make_double2(bs[n0], 0.0);
            M[n0 * LD + n0] = d0;
            M[n0 * LD + P] = bb;
            M[P * LD + n0] = bb;
        }
        if (v1) {
            const double2 bb = make_double2(bs[n1], 0.0);
            M[n1 * LD + n1] = d1;
            M[n1 * LD + P] = bb;
            M[P * LD + n1] = bb;
        }
        for (int j = N + lane; j <= P; j += 32) M[P * LD + j] = zero2;  // z row: padding + corner
        __syncwarp();
    }
    // d0, d1: diagonal entries of rows lane and lane + 32.  P <= 64: threshold-pivoted symmetric elimination
    // first (lu_sym.cuh), partial pivoting (lu_warp.cuh) when a diagonal entry fails the test or P > 64.
    __device__ __forceinline__ double2 solve(const double2 d0, const double2 d1, int& singular, const bool symmetric,
                                             unsigned long long* fallbacks, int& fellback, int& /*deferred*/) {
        assemble(d0, d1);
        if (symmetric && g.P <= 64) {
            int fail = 0;
            const double2 corner = sym_schur_inplace(M, g.P, g.LD, lane, fail);
            __syncwarp();
            if (!fail) return corner;
            if (lane == 0 && fallbacks) atomicAdd(fallbacks, 1ULL);
            fellback = 1;
            assemble(d0, d1);
        }
        const double2 corner = bordered_schur<MAXNS>(M, g, lane, singular);
        __syncwarp();
        return corner;
    }
};
template <>
struct WarpSolver<-2> : GenericSolver<2> {};
template <>
struct WarpSolver<-3> : GenericSolver<3> {};

// per-warp shared-memory layout of the fast paths: W (large enough for both the symmetric and the
// partial-pivoting variant), the diagonal d, the row permutation of the pivoting variant
template <int PT>
struct FastLayout {
    static constexpr int WALLOC = FastGeom<PT>::WALLOC > SymGeom<PT>::WALLOC ? FastGeom<PT>::WALLOC : SymGeom<PT>::WALLOC;
    static constexpr int WARP_BYTES = WALLOC * 16 + PT * 16 + FastGeom<PT>::ORIG_BYTES;
};

template <int PT>
struct SweepPlan {
    using G = FastLayout<PT>;
    static constexpr int kShared = (PT * PT + PT + PT) * 8;
    static constexpr int kFit = (kSweepMaxSmem - kShared) / G::WARP_BYTES;
#ifndef JITR_FAST_WARPS
#define JITR_FAST_WARPS 16
#endif
    static constexpr int kMaxWarps = kFit > JITR_FAST_WARPS ? JITR_FAST_WARPS : kFit;
    static constexpr int kThreads = kMaxWarps * 32;
    static bool supports(int N) { return N > PT - 8 && N <= PT; }
    static int warps(int) { return kMaxWarps; }
    static int smem(int, int w) { return kShared + w * G::WARP_BYTES; }
};

// The partial-pivoting re-solve of a system that failed the threshold test: rare (1e-4 of the systems),
// deliberately out of line so that its code stays out of the hot loop's instruction footprint and out of
// the compiler's convergence analysis of the main kernel (with it inlined the analysis gave up and wrapped
// every warp-collective instruction of the hot path in divergence guards: 15.5k instead of 7k instructions).
template <int PT>
static __device__ __noinline__ double2 pivoting_fallback(unsigned wofs, double2 d0, double2 d1, int lane, int* singular) {
    using G = FastLayout<PT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const double* Tp = reinterpret_cast<const double*>(smem_raw);
    const double* bsh = Tp + PT * PT;
    double2* W = reinterpret_cast<double2*>(smem_raw + wofs);
    double2* dvec = W + G::WALLOC;
    unsigned char* origv = reinterpret_cast<unsigned char*>(dvec + PT);
    constexpr int NS = FastGeom<PT>::NS0;
    double2 d[NS];
    d[0] = d0;
    if (NS > 1) d[NS - 1] = d1;
    int sing = 0;
    const double2 corner = fast_schur<PT>(W, dvec, origv, Tp, bsh, d, lane, sing);
    if (sing) *singular = 1;
    return corner;
}

template <int PT>
struct WarpSolver {
    using G = FastLayout<PT>;
    // only a byte offset is kept in registers; the pointers are re-derived from the dynamic shared
    // memory base at every solve (register pressure is what limits this kernel to 168 registers)
    unsigned wofs;
    int N, lane;
    static constexpr int kShared = (PT * PT + PT + PT) * 8;  // That (padded), b, 1/x^2
    __device__ __forceinline__ void init(unsigned char* smem, const double* That, const double* bvec,
                                         const double* mesh_x, int N_, int nwarps, int warp, int lane_) {
        N = N_;
        lane = lane_;
        double* T = reinterpret_cast<double*>(smem);
        double* b = T + PT * PT;
        double* c = b + PT;
        wofs = kShared + warp * G::WARP_BYTES;
        for (int i = threadIdx.x; i < PT * PT; i += blockDim.x) {  // That padded with the identity
            const int r = i / PT, cc = i - r * PT;
            T[i] = (r < N && cc < N) ? That[r * N + cc] : (r == cc ? 1.0 : 0.0);
        }
        for (int i = threadIdx.x; i < PT; i += blockDim.x) {
            b[i] = i < N ? bvec[i] : 0.0;
            c[i] = i < N ? 1.0 / (mesh_x[i] * mesh_x[i]) : 0.0;
        }
    }
    // centrifugal factors 1/x_n^2 of this lane's two mesh points
    __device__ __forceinline__ void centrifugal(double& c0, double& c1) const {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        const double* c = reinterpret_cast<const double*>(smem_raw) + PT * PT + PT;
        c0 = c[lane];
        c1 = c[lane + 32 < PT ? lane + 32 : 0];
    }
    // Threshold-pivoted symmetric elimination first (lu_sym.cuh); a system whose diagonal fails the
    // threshold test is re-solved with partial pivoting (lu_fast.cuh).  `fallbacks` counts those.
    __device__ __forceinline__ double2 solve(const double2 d0, const double2 d1, int& singular, const bool symmetric,
                                             unsigned long long* fallbacks, int& fellback, int& /*deferred*/) {
        const double2 one = make_double2(1.0, 0.0);
        const double2 da = lane < N ? d0 : one, db = lane + 32 < N ? d1 : one;
        if (symmetric) {
            extern __shared__ __align__(16) unsigned char smem_raw[];
            const double* Tp = reinterpret_cast<const double*>(smem_raw);
            const double* bsh = Tp + PT * PT;
            double2* W = reinterpret_cast<double2*>(smem_raw + wofs);
            double2* dvec = W + G::WALLOC;
            constexpr int NS = SymGeom<PT>::NS0;
            double2 d[NS];
            d[0] = da;
            if (NS > 1) d[NS - 1] = db;
            int fail = 0;
            const double2 corner = sym_schur<PT>(W, dvec, Tp, bsh, d, lane, fail);
            __syncwarp();
            if (!fail) return corner;
            if (lane == 0 && fallbacks) atomicAdd(fallbacks, 1ULL);
            fellback = 1;
        }
        int sing = 0;
        const double2 corner = pivoting_fallback<PT>(wofs, da, db, lane, &sing);
        __syncwarp();
        if (__any_sync(0xffffffffu, sing)) singular = 1;
        return corner;
    }
};

// ---------------------------------------------------------------------------------------------
// Packed modes, P = 48 / 56 / 64 (BASELINE config 3: N = 60): only the LOWER TRIANGLE of the Schur complement of the first
// panel is ever stored (lu_packed.cuh: 29.6 KB at P = 64 instead of the 67.6 KB of the assembled bordered matrix -> seven
// systems per SM instead of three).  The first panel is eliminated in registers straight from That (global memory, 28.8 KB,
// read-only, L1/L2 resident), the rest in place by the threshold-pivoted packed LDL^T.  A system that fails the threshold
// test defers its whole (sample, energy) pair to the partial-pivoting modes.
// ---------------------------------------------------------------------------------------------
template <int PT>
struct PackedSweepPlan {
    using G = PackedGeom<PT - 8>;  // the Schur complement of the first panel is all that is stored (packed_first_panel)
    static constexpr int kShared = 2 * PT * 8;  // b, 1/x^2
    static constexpr int kFit = (kSweepMaxSmem - kShared) / (G::WARP_ELEMS * 16);
    // at most JITR_PACKED_WARPS warps (12: 168 registers per thread -- the first panel keeps all B fragments in registers)
#ifndef JITR_PACKED_WARPS
#define JITR_PACKED_WARPS 12
#endif
    static constexpr int kMaxWarps = kFit > JITR_PACKED_WARPS ? JITR_PACKED_WARPS : kFit;
    static constexpr int kThreads = kMaxWarps * 32;
    static bool supports(int N) { return N > PT - 8 && N <= PT; }
    static int warps(int) { return kMaxWarps; }
    static int smem(int, int w) { return kShared + w * G::WARP_ELEMS * 16; }
};
template <int PT>
struct PackedSolver {
    using G = PackedGeom<PT - 8>;
    unsigned wofs;
    int N, lane;
    const double* That;
    __device__ __forceinline__ void init(unsigned char* smem, const double* That_, const double* bvec, const double* mesh_x,
                                         int N_, int nwarps, int warp, int lane_) {
        N = N_;
        lane = lane_;
        That = That_;
        wofs = (unsigned)warp * G::WARP_ELEMS * 16;
        double* b = reinterpret_cast<double*>(smem + (size_t)nwarps * G::WARP_ELEMS * 16);
        double* c = b + PT;
        for (int i = threadIdx.x; i < PT; i += blockDim.x) {
            b[i] = i < N ? bvec[i] : 0.0;
            c[i] = i < N ? 1.0 / (mesh_x[i] * mesh_x[i]) : 0.0;
        }
    }
    __device__ __forceinline__ const double* tail() const {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        return reinterpret_cast<const double*>(smem_raw + (size_t)(blockDim.x >> 5) * G::WARP_ELEMS * 16);
    }
    __device__ __forceinline__ void centrifugal(double& c0, double& c1) const {
        const double* c = tail() + PT;
        c0 = c[lane];
        c1 = c[lane + 32 < PT ? lane + 32 : 0];
    }
    __device__ __forceinline__ double2 solve(const double2 d0, const double2 d1, int& /*singular*/, const bool /*symmetric*/,
                                             unsigned long long* fallbacks, int& /*fellback*/, int& deferred) {
        extern __shared__ __align__(16) unsigned char smem_raw[];
        double2* M = reinterpret_cast<double2*>(smem_raw + wofs);
        const double* b = tail();
        const double2 one = make_double2(1.0, 0.0);
        const double2 dd[2] = {lane < N ? d0 : one, lane + 32 < N ? d1 : one};  // identity padding beyond N
        // first panel straight from That (no assembly pass), its Schur complement into the packed region, then in place
        double2 corner = make_double2(0.0, 0.0);
        int fail = 0, bad = 0;
        packed_first_panel<PT>(M, That, b, dd, N, lane, corner, fail);
        corner = packed_ldlt<PT - 8>(M, lane, bad, corner);
        __syncwarp();
        bad |= fail;
        if (bad) {
            if (lane == 0 && fallbacks) atomicAdd(fallbacks, 1ULL);
            deferred = 1;
        }
        return corner;
    }
};
#ifdef JITR_PACKED40
template <> struct SweepPlan<40> : PackedSweepPlan<40> {};
template <> struct WarpSolver<40> : PackedSolver<40> {};
#endif
template <> struct SweepPlan<48> : PackedSweepPlan<48> {};
template <> struct SweepPlan<56> : PackedSweepPlan<56> {};
template <> struct SweepPlan<64> : PackedSweepPlan<64> {};
template <> struct WarpSolver<48> : PackedSolver<48> {};
template <> struct WarpSolver<56> : PackedSolver<56> {};
template <> struct WarpSolver<64> : PackedSolver<64> {};

// ---------------------------------------------------------------------------------------------
// Loose lock step of the warps of a CTA: warp w may start iteration it only after EVERY warp has
// finished iteration it - K (K mbarriers used round robin, one arrival per warp and iteration).
// The warps stay within one solve of each other -- close enough to share the instruction cache --
// but a warp that needs longer for one iteration (potential evaluation at the start of a pair,
// the partial-pivoting fallback) does not stall the other eleven the way a full barrier did
// (ncu before: 15 % of all warp cycles in the barrier).
// ---------------------------------------------------------------------------------------------
#ifndef JITR_LOCKSTEP_SLACK
#define JITR_LOCKSTEP_SLACK 2
#endif
struct LockStep {
    static constexpr int K = JITR_LOCKSTEP_SLACK;  // iterations a warp may run ahead of the slowest one
    unsigned bar0;  // shared-memory address of the first of the K mbarriers
    __device__ __forceinline__ void init(unsigned long long* smem_bars, int nwarps) {
        bar0 = (unsigned)__cvta_generic_to_shared(smem_bars);
        if (threadIdx.x == 0) {
            for (int k = 0; k < K; ++k)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * k), "r"(nwarps));
        }
    }
    // block until the phase `parity` of barrier b has completed
    __device__ __forceinline__ void wait(int b, unsigned parity) const {
        unsigned done;
        do {
            asm volatile(
                "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                : "=r"(done)
                : "r"(bar0 + 8 * b), "r"(parity)
                : "memory");
        } while (!done);
    }
    __device__ __forceinline__ void arrive(int b, int lane) const {
        __syncwarp();
        if (lane == 0) {
            asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar0 + 8 * b) : "memory");
        }
    }
};

// Built-in potential forms at this lane's two mesh points.  Deliberately NOT inlined: it runs once per
// (sample, energy) pair, its code is large (parameter maps of four models, ten exponentials), and
// inlined into the solve loop it evicted the LU code from the instruction cache (ncu: 2.7
// no-instruction stall cycles per issued instruction, profiles/r01_sweep_v4_icache.txt).
struct PairPotentials {
    double2 vc0, vs0, vc1, vs1;
};
static __device__ __noinline__ PairPotentials eval_pair_potentials(int model, const double* __restrict__ prow,
                                                            const double* __restrict__ et, double r0, double r1) {
    const Shape sh = shape_from_model(model, prow, et);
    PairPotentials o;
    double co;
    const double rr[2] = {r0, r1};
    double2 vc[2], vs[2];
#pragma unroll  // both mesh points in one instruction stream: the chains (exp, divisions) are latency bound
    for (int s = 0; s < 2; ++s) {
        eval_shape(sh, rr[s], vc[s], vs[s], co);
        vc[s].x += co;
    }
    o.vc0 = vc[0]; o.vs0 = vs[0]; o.vc1 = vc[1]; o.vs1 = vs[1];
    return o;
}

// PARAMS: built-in potential forms from parameter rows; otherwise mesh values from tensors.
template <bool PARAMS, int MODE>
__global__ void __launch_bounds__(SweepPlan<MODE>::kThreads, 1) elastic_sweep_kernel(const SweepArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = A.N;
    const int nwarps = blockDim.x >> 5;
    // broadcast from lane 0: tells the compiler the warp index (hence the whole work loop) is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    WarpSolver<MODE> solver;
    solver.init(smem_raw, A.That, A.bvec, A.mesh_x, N, nwarps, warp, lane);
    __syncthreads();

    // mesh points owned by this lane (two slots, N <= 64)
    const int n0 = lane, n1 = lane + 32;
    const bool v1 = n1 < N;
    const bool v0 = n0 < N;
    const int L1 = A.lmax + 1;
    const double2 zero2 = make_double2(0.0, 0.0);

    // The warps of the CTA advance in (loose) lock step, ONE SOLVE PER ITERATION.  Every solve executes
    // the same ~45 KB of unrolled code; left alone, the early break lets the 12 warps drift to unrelated
    // code regions and the SM's instruction cache thrashes (ncu: 2.5 no-instruction stall cycles per
    // issued instruction, icc hit rate 63 %); aligned, one warp's fetch serves the other eleven.
    // Pairs are handed out dynamically (one atomic per pair): the number of partial waves per pair varies
    // threefold with the energy, and a static round-robin (stride = #warps, a multiple of 16) gave every
    // warp a fixed subset of 4 of the 64 energies -- up to 21 % of a warp's time was an idle tail.
    auto next_pair = [&]() -> long long {
        unsigned long long v = 0;
        if (lane == 0) {
            v = atomicAdd(A.work_counter, 1ULL);
            if (A.pair_list) v = v < (unsigned long long)*A.pair_count ? (unsigned long long)A.pair_list[v] : (unsigned long long)A.n_pairs;
        }
        return (long long)__shfl_sync(0xffffffffu, v, 0);
    };
    // l-split mode: a work item is one partial wave l of one pair (both j), item = pair * (lmax + 1) + l
    const long long n_items = A.lsplit ? A.n_pairs * L1 : A.n_pairs;
    long long item = next_pair();
    long long pair = A.lsplit ? item / L1 : item;
    bool active = item < n_items;  // warp-uniform
    bool fresh = true;               // the pair's potentials have not been evaluated yet
    int e = 0, l = 0, jj = 0, nl = 0, singular = 0, nonfinite = 0;
    bool plus_small = false;  // |1 - S+| < tol at the current l
    double2 d0 = zero2, d1 = zero2, so0 = zero2, so1 = zero2;
    __shared__ unsigned long long lock_bars[LockStep::K];
    // Termination: every warp must leave the loop at the SAME iteration, or a warp that went on would wait
    // for an arrival that never comes.  A warp records the iteration of its last solve (last_iter, a
    // maximum) BEFORE it gives up its count in warps_active; warps_active == 0 then implies last_iter is
    // final, and every warp leaves at iteration last_iter + K -- which each of them reaches, because the
    // wait at that iteration needs only arrivals up to iteration last_iter.
    __shared__ int warps_active, last_iter;
    LockStep lock;
    lock.init(lock_bars, nwarps);
    if (threadIdx.x == 0) {
        warps_active = nwarps;
        last_iter = -1;
    }
    __syncthreads();
    if (!active && lane == 0) atomicSub(&warps_active, 1);
#ifdef JITR_SWEEP_STAGGER
    // tuning experiment: odd warps start late, so that the panel phases (shared-memory / latency bound) of one half of
    // the CTA overlap the trailing updates (FP64-pipe bound) of the other half
    if (warp & 1) {
        const long long t0 = clock64();
        while (clock64() - t0 < JITR_SWEEP_STAGGER) __nanosleep(200);
    }
#endif
    // The packed modes run their panels in a runtime loop: their elimination code is a few KB and fits the instruction cache
    // whatever the warps are doing, so they run free (JITR_PACKED_LOCKSTEP=1 puts them in lock step again).
#ifndef JITR_PACKED_LOCKSTEP
#define JITR_PACKED_LOCKSTEP 0
#endif
#ifdef JITR_PACKED40
    constexpr bool kLock = MODE < 40 || JITR_PACKED_LOCKSTEP != 0;
#else
    constexpr bool kLock = MODE < 48 || JITR_PACKED_LOCKSTEP != 0;
#endif
    for (unsigned it = 0;; ++it) {
        const int b = it % LockStep::K;
        if (kLock) {
            if (it >= LockStep::K) lock.wait(b, (it / LockStep::K - 1) & 1);  // every warp has finished iteration it - K
            const int na = *(volatile int*)&warps_active;
            __threadfence_block();
            const int li = *(volatile int*)&last_iter;
            if (__all_sync(0xffffffffu, na == 0 && (int)it >= li + LockStep::K)) break;
        } else if (!active) {
            break;
        }
        if (!active) {
            lock.arrive(b, lane);
            continue;
        }
        if (fresh) {
            const long long smp = pair / A.n_energies;
            e = (int)(pair - smp * A.n_energies);
            const double* et = A.etab + (size_t)e * JITR_B200_ET_STRIDE;
            const double a = et[JITR_B200_ET_A], E0 = et[JITR_B200_ET_ECM], radius = et[JITR_B200_ET_RCH];
            const double a2 = a * a;
            const double x0 = A.mesh_x[v0 ? n0 : 0], x1 = A.mesh_x[v1 ? n1 : 0];
            const double t0 = A.That[(v0 ? n0 : 0) * (N + 1)], t1 = A.That[(v1 ? n1 : 0) * (N + 1)];
            // ---- potentials on this lane's mesh points, in the a^2-scaled form of the diagonal
            double2 vc0, vs0, vc1, vs1;
            if (PARAMS) {
                const PairPotentials pp =
                    eval_pair_potentials(A.model, A.params + (size_t)(A.model == JITR_B200_MODEL_SHAPE_PAIR ? pair : smp) * A.n_params,
                                         et, x0 * radius, x1 * radius);
                vc0 = pp.vc0; vs0 = pp.vs0; vc1 = pp.vc1; vs1 = pp.vs1;
            } else {
                const size_t off = (size_t)pair * N;
                vc0 = A.central[off + (v0 ? n0 : 0)];
                vc1 = A.central[off + (v1 ? n1 : 0)];
                vs0 = vs1 = zero2;
                if (A.spin_orbit) {
                    vs0 = A.spin_orbit[off + (v0 ? n0 : 0)];
                    vs1 = A.spin_orbit[off + (v1 ? n1 : 0)];
                }
                if (A.coulomb) {
                    const double2 c0 = A.coulomb[off + (v0 ? n0 : 0)], c1 = A.coulomb[off + (v1 ? n1 : 0)];
                    vc0.x += c0.x; vc0.y += c0.y;
                    vc1.x += c1.x; vc1.y += c1.y;
                }
            }
            // diag(a^2 A) = That_nn + l(l+1)/x_n^2 + a^2(-1 + V/E0) + (l.s) a^2 Vso/E0
            d0 = make_double2(t0 + a2 * (vc0.x / E0 - 1.0), a2 * (vc0.y / E0));
            d1 = make_double2(t1 + a2 * (vc1.x / E0 - 1.0), a2 * (vc1.y / E0));
            so0 = make_double2(a2 * (vs0.x / E0), a2 * (vs0.y / E0));
            so1 = make_double2(a2 * (vs1.x / E0), a2 * (vs1.y / E0));
            l = A.lsplit ? (int)(item - pair * L1) : 0;
            jj = 0; nl = 0; singular = 0; nonfinite = 0;
            fresh = false;
        }
        // ---- one system: partial wave l, j = l + 1/2 (jj = 0) or l - 1/2 (jj = 1)
        const double ll1 = (double)l * (double)(l + 1);
        const double ls = (jj == 0) ? (double)l : -(double)(l + 1);
        double cent0, cent1;
        solver.centrifugal(cent0, cent1);
        const double2 dd0 = make_double2(d0.x + ll1 * cent0 + ls * so0.x, d0.y + ls * so0.y);
        const double2 dd1 = make_double2(d1.x + ll1 * cent1 + ls * so1.x, d1.y + ls * so1.y);
        // corner = -b^T (a^2 A)^-1 b = -R
        int fellback = 0, deferred = 0;
        const double2 corner = solver.solve(dd0, dd1, singular, A.symmetric != 0, A.fallbacks, fellback, deferred);
        if ((fellback | deferred) && A.fallback_mask && lane == 0) atomicOr(A.fallback_mask + pair, 1ULL << (2 * l + jj));
        bool pair_done = false;
        if (deferred) {  // warp-uniform: the pair is abandoned here and redone by the partial-pivoting modes
            if (lane == 0 && (!A.lsplit || atomicExch(A.lsplit_seen + pair, 1) == 0)) A.defer_list[atomicAdd(A.defer_count, 1)] = (int)pair;
            pair_done = true;
        } else {
        const double2* as = A.asym + ((size_t)e * L1 + l) * 4;  // (after the solve: fewer live registers)
        const double2 Hp = as[0], Hm = as[1], Hpp = as[2], Hmp = as[3];
        const double a = A.etab[(size_t)e * JITR_B200_ET_STRIDE + JITR_B200_ET_A];
        const double2 aR = make_double2(-a * corner.x, -a * corner.y);
        // S = (H- - a R H-') / (H+ - a R H+')   (rmatrix/core.py:51-56 with nch = 1)
        const double2 t_m = cmul(aR, Hmp), t_p = cmul(aR, Hpp);
        const double2 num = make_double2(Hm.x - t_m.x, Hm.y - t_m.y);
        const double2 den = make_double2(Hp.x - t_p.x, Hp.y - t_p.y);
        const double2 S = cmul(num, crecip_fast(den));
        if (!(isfinite(S.x) && isfinite(S.y))) nonfinite = 1;
        double2* Sp_out = A.S_out + (size_t)pair * 2 * L1;
        double2* Sm_out = Sp_out + L1;
        // |1 - S| < tol  (xs/elastic.py:178-181), compared in squares
        const double e1 = 1.0 - S.x;
        const bool small = fma(e1, e1, S.y * S.y) < A.tol * A.tol;
        if (jj == 0) {
            plus_small = small;
            if (lane == 0) Sp_out[l] = S;
            if (l == 0) {  // l = 0 has the single j = 1/2 system; S- [0] = 0 (xs/elastic.py:112,152)
                if (lane == 0) Sm_out[0] = zero2;
                nl = 1;
                pair_done = A.lmax == 0 || A.lsplit;
                l = 1;
            } else {
                jj = 1;
            }
        } else {
            if (lane == 0) Sm_out[l] = S;
            nl = l + 1;
            pair_done = l == A.lmax || A.lsplit;
            if (A.tol > 0.0 && !A.lsplit) {
                // (the vote keeps every branch of this loop provably warp-uniform for the compiler)
                if (__all_sync(0xffffffffu, small && plus_small)) pair_done = true;
            }
            l += 1;
            jj = 0;
        }
        }  // !deferred
        if (__any_sync(0xffffffffu, pair_done)) {
            if (A.lsplit) {
                if (lane == 0) A.lsplit_flags[item] = nonfinite ? JITR_B200_STATUS_NONFINITE : (singular ? JITR_B200_STATUS_SINGULAR : JITR_B200_STATUS_OK);
            } else if (!deferred) {
                double2* Sp_out = A.S_out + (size_t)pair * 2 * L1;
                double2* Sm_out = Sp_out + L1;
                for (int q = nl + lane; q <= A.lmax; q += 32) {
                    Sp_out[q] = zero2;
                    Sm_out[q] = zero2;
                }
                if (lane == 0) {
                    A.nl_out[pair] = nl;
                    A.status[pair] =
                        nonfinite ? JITR_B200_STATUS_NONFINITE : (singular ? JITR_B200_STATUS_SINGULAR : JITR_B200_STATUS_OK);
                }
            }
            item = next_pair();
            pair = A.lsplit ? item / L1 : item;
            active = item < n_items;
            fresh = true;
            if (kLock && !active && lane == 0) {
                atomicMax(&last_iter, (int)it);
                __threadfence_block();
                atomicSub(&warps_active, 1);
            }
        }
        if (kLock) lock.arrive(b, lane);
    }
}

// l-split mode, second step: the reference's data-dependent break (xs/elastic.py:178-181) applied to the finished S-matrices
// of a pair -- same comparison, same order -- then zeros beyond the kept partial waves, nl and the status of the kept ones.
static __global__ void sweep_finalize_kernel(const SweepArgs A) {
    const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= A.n_pairs) return;
    const int L1 = A.lmax + 1;
    double2* Sp = A.S_out + (size_t)pair * 2 * L1;
    double2* Sm = Sp + L1;
    int nl = 1;
    for (int l = 1; l <= A.lmax; ++l) {
        nl = l + 1;
        if (A.tol > 0.0) {
            const double ep = 1.0 - Sp[l].x, em = 1.0 - Sm[l].x;
            if (fma(ep, ep, Sp[l].y * Sp[l].y) < A.tol * A.tol && fma(em, em, Sm[l].y * Sm[l].y) < A.tol * A.tol) break;
        }
    }
    int nonfinite = 0, singular = 0;
    for (int l = 0; l < nl; ++l) {
        const int f = A.lsplit_flags[pair * L1 + l];
        nonfinite |= f == JITR_B200_STATUS_NONFINITE;
        singular |= f == JITR_B200_STATUS_SINGULAR;
    }
    const double2 zero2 = make_double2(0.0, 0.0);
    for (int l = nl; l <= A.lmax; ++l) Sp[l] = Sm[l] = zero2;
    A.nl_out[pair] = nl;
    A.status[pair] = nonfinite ? JITR_B200_STATUS_NONFINITE : (singular ? JITR_B200_STATUS_SINGULAR : JITR_B200_STATUS_OK);
}

template <bool PARAMS, int MODE>
int launch_sweep_mode(const SweepArgs& a, cudaStream_t st) {
    using Plan = SweepPlan<MODE>;
    const int warps = Plan::warps(a.N);
    if (warps < 1) return set_error(JITR_B200_ERR_UNSUPPORTED, "elastic sweep: system does not fit in shared memory");
    const int smem = Plan::smem(a.N, warps);
    auto kern = elastic_sweep_kernel<PARAMS, MODE>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
        return set_cuda_error("cudaFuncSetAttribute(elastic_sweep)");
    const int sms = sm_count();
    SweepArgs args = a;
    // small batches: fewer pairs than half the resident warps -> one work item per partial wave (a packed mode defers a pair
    // at most once: lsplit_seen)
    args.lsplit = 0;
    if (MODE > 0 && !a.pair_list && a.lmax > 0 && 2 * a.n_pairs < (long long)sms * warps) {
        args.lsplit_flags = defer_buffer(a.n_pairs * (a.lmax + 2), st);
        if (!args.lsplit_flags) return set_cuda_error("elastic sweep: l-split scratch");
        args.lsplit_flags += 4;
        args.lsplit_seen = args.lsplit_flags + a.n_pairs * (a.lmax + 1);
        if (cudaMemsetAsync(args.lsplit_seen, 0, (size_t)a.n_pairs * sizeof(int), st) != cudaSuccess)
            return set_cuda_error("elastic sweep: l-split scratch");
        args.lsplit = 1;
    }
    const long long items = args.lsplit ? a.n_pairs * (a.lmax + 1) : a.n_pairs;
    long long blocks = (items + warps - 1) / warps;
    if (blocks > sms) blocks = sms;
    if (blocks < 1) return JITR_B200_OK;
    args.work_counter = next_work_counter(st);
    if (!args.work_counter) return set_cuda_error("elastic_sweep work counter");
    kern<<<(unsigned)blocks, warps * 32, smem, st>>>(args);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("elastic_sweep launch");
    if (args.lsplit) {
        sweep_finalize_kernel<<<(unsigned)((a.n_pairs + 127) / 128), 128, 0, st>>>(args);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) return set_cuda_error("elastic_sweep finalize launch");
    }
    return JITR_B200_OK;
}

}  // namespace jitr
