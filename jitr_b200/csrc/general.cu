// General batched R-matrix solve: any number of channels, local and/or nonlocal interaction,
// optional wavefunction coefficients.  One CTA per system; the augmented matrix [A | B] lives in
// shared memory when it fits (n <= ~115) and in the caller's device workspace otherwise.
// Replaces Solver.solve (rmatrix/rmatrix.py:180-246) + rmatrix/core.py:7-82 for one system per
// call in the reference.  The fused warp-per-system kernel in elastic.cu is the fast path for the
// single-channel elastic sweep; this kernel is the fully general one.
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"
#include "jitr_b200.h"
#include "lu_warp.cuh"

namespace jitr {

struct GenArgs {
    int N, nch, n;  // n = N * nch
    long long batch;
    const double2* free_matrix;
    long long free_stride;
    const double2* local;
    const double2* nonlocal_;
    const double2* interaction;
    const double* wsqrt;
    const double* sys;
    int sys_stride;
    const double* bvec;
    const double2* asym;
    long long asym_stride;
    const double* weights;
    double2* R;
    double2* S;
    double2* uext;
    double2* coeffs;
    double2* workspace;
    int* status;
};

constexpr int kGenThreads = 256;
constexpr int kMaxCh = 16;

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
        const double2* F = A.free_matrix + (size_t)sysi * A.free_stride;
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
        double2* Zp = small + kMaxCh * kMaxCh;        // [nch][nch]
        double2* Zm = small + 2 * kMaxCh * kMaxCh;    // [nch][nch] -> S
        double2* ue = small + 3 * kMaxCh * kMaxCh;    // [nch]
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
            int bad = redi[33];
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
// Single-channel systems (nch = 1, no coefficients) up to P = 8 ceil(N/8) <= 64: one WARP per system, the
// bordered matrix [[F + V, b], [b^T, 0]] assembled in the warp's shared-memory region straight from the
// global arrays (the nonlocal kernel is the HBM stream of this path: N^2 x 16 B per system) and eliminated
// by the blocked partial-pivoting LU of lu_warp.cuh (register panel + DMMA trailing update).  This is
// BASELINE config 4 (nonlocal Perey-Buck, N = 60): 7x the CTA-per-system kernel above.
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
        const double2* F = A.free_matrix + (size_t)sysi * A.free_stride;
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

template <int MAXNS>
static int launch_warp_solve(const GenArgs& a, cudaStream_t st) {
    const int P = (a.N + 7) & ~7, msz = (P + 1) * (P + 1);
    int warps = 232448 / (msz * 16);
    if (warps > 8) warps = 8;
    if (warps < 1) return set_error(JITR_B200_ERR_UNSUPPORTED, "warp_solve: system does not fit in shared memory");
    const int smem = warps * msz * 16;
    static int conf = 0;
    if (smem > conf) {
        if (cudaFuncSetAttribute(warp_solve_kernel<MAXNS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
            return set_cuda_error("cudaFuncSetAttribute(warp_solve)");
        conf = smem;
    }
    long long blocks = (a.batch + warps - 1) / warps;
    if (blocks > sm_count()) blocks = sm_count();
    warp_solve_kernel<MAXNS><<<(unsigned)blocks, warps * 32, smem, st>>>(a);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("warp_solve launch");
    return JITR_B200_OK;
}

}  // namespace jitr

using namespace jitr;

extern "C" int jitr_b200_solve_batched(int nbasis, int nch, long long batch, const double* free_matrix,
                                       long long free_stride, const double* local, const double* nonlocal_,
                                       const double* interaction, const double* wsqrt, const double* sys,
                                       int sys_stride, const double* bvec, const double* asym, long long asym_stride,
                                       const double* weights, double* R, double* S, double* uext, double* coeffs,
                                       double* workspace, int* status, void* stream) {
    if (nbasis < 1 || nch < 1 || nch > kMaxCh || batch < 0 || !free_matrix || !sys || !bvec || !asym || !weights || !R ||
        !S || !uext || !status)
        return set_error(JITR_B200_ERR_BAD_ARG, "solve_batched: null pointer or bad size (nch <= 16)");
    if (nonlocal_ && !wsqrt) return set_error(JITR_B200_ERR_BAD_ARG, "solve_batched: nonlocal potential needs wsqrt");
    if (batch == 0) return JITR_B200_OK;
    GenArgs a{};
    a.N = nbasis; a.nch = nch; a.n = nbasis * nch; a.batch = batch;
    a.free_matrix = reinterpret_cast<const double2*>(free_matrix); a.free_stride = free_stride;
    a.local = reinterpret_cast<const double2*>(local);
    a.nonlocal_ = reinterpret_cast<const double2*>(nonlocal_);
    a.interaction = reinterpret_cast<const double2*>(interaction);
    a.wsqrt = wsqrt; a.sys = sys; a.sys_stride = sys_stride; a.bvec = bvec;
    a.asym = reinterpret_cast<const double2*>(asym); a.asym_stride = asym_stride; a.weights = weights;
    a.R = reinterpret_cast<double2*>(R); a.S = reinterpret_cast<double2*>(S);
    a.uext = reinterpret_cast<double2*>(uext); a.coeffs = reinterpret_cast<double2*>(coeffs);
    a.workspace = reinterpret_cast<double2*>(workspace); a.status = status;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // single channel, no coefficients, P <= 64: warp-per-system blocked LU
    if (nch == 1 && !coeffs && ((nbasis + 7) & ~7) <= 64) {
        if (((nbasis + 7) & ~7) + 1 <= 64) return launch_warp_solve<2>(a, st);
        return launch_warp_solve<3>(a, st);
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
        static size_t conf = 0;
        if (smem > conf) {
            if (cudaFuncSetAttribute(general_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return set_cuda_error("cudaFuncSetAttribute(general_solve<smem>)");
            conf = smem;
        }
        general_solve_kernel<true><<<(unsigned)blocks, kGenThreads, smem, st>>>(a);
    } else {
        static size_t conf = 0;
        if (smem > conf) {
            if (cudaFuncSetAttribute(general_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return set_cuda_error("cudaFuncSetAttribute(general_solve<global>)");
            conf = smem;
        }
        general_solve_kernel<false><<<(unsigned)blocks, kGenThreads, smem, st>>>(a);
    }
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("general_solve launch");
    return JITR_B200_OK;
}
