// Percentile bands over the sample axis, on the device (the np.percentile(xs, q, axis=samples) of the reference's
// kduq_uq_demo notebook, cell 18), so that only O(n_quantiles x n_angles) numbers leave the GPU after a sweep.
//
// values [S][M] (sample-major: M = energies x angles columns), exact order statistics by MSB radix selection on the
// order-preserving 64-bit image of the doubles: eight passes of 8-bit digits with shared-memory histograms, one CTA
// per 16 adjacent columns (a row of the tile is one 128-byte line), all requested quantiles at once; a ninth pass
// finds the upper neighbour of each order statistic for numpy's default linear interpolation.
#include <cuda_runtime.h>

#include "common.cuh"
#include "jitr_b200.h"

namespace jitr {

constexpr int kPctCols = 16;     // columns per CTA
constexpr int kPctRows = 16;     // rows per pass of the 256 threads
constexpr int kPctMaxQ = 8;      // quantiles per call

__device__ __forceinline__ unsigned long long pct_key(double x) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);  // total order of the finite doubles (and +-inf)
}
__device__ __forceinline__ double pct_val(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)u);
}

struct PctQ {
    double q[kPctMaxQ];
};

// numpy's virtual index for method="linear": (n - 1) * (percent / 100), without FMA contraction
__device__ __forceinline__ double pct_virtual_index(long long n, double percent) {
    return __dmul_rn((double)(n - 1), __ddiv_rn(percent, 100.0));
}

// NQ = number of quantiles, a compile-time value: the per-thread arrays over the quantiles stay in registers (the runtime-nq
// version of round 1 indexed them dynamically and kept them in local memory: 534 LDL/STL in its SASS)
template <int NQ>
__global__ void __launch_bounds__(kPctCols* kPctRows) percentile_kernel(const double* __restrict__ x, long long S, long long M,
                                                                        const PctQ Q, double* __restrict__ out) {
    constexpr int nq = NQ;
    extern __shared__ __align__(16) unsigned char pct_smem[];
    typedef unsigned long long (*ull_rows)[kPctCols];
    ull_rows prefix = reinterpret_cast<ull_rows>(pct_smem);                                 // [nq][16] selected key so far
    long long(*rank)[kPctCols] = reinterpret_cast<long long(*)[kPctCols]>(prefix + nq);     // rank left inside the class
    ull_rows above = reinterpret_cast<ull_rows>(rank + nq);                                 // smallest key above the selection
    ull_rows n_le = above + nq;                                                             // elements <= the selection
    unsigned(*hist)[kPctCols][256] = reinterpret_cast<unsigned(*)[kPctCols][256]>(n_le + nq);  // [nq][16][256]
    __shared__ unsigned nan_col[kPctCols];  // the column holds a NaN: numpy.percentile returns NaN for it
    const double* q = Q.q;
    const int tc = threadIdx.x % kPctCols, tr = threadIdx.x / kPctCols;
    const long long col = (long long)blockIdx.x * kPctCols + tc;
    const bool live = col < M;
    if (threadIdx.x < kPctCols) nan_col[threadIdx.x] = 0u;
    for (int i = threadIdx.x; i < nq * kPctCols; i += blockDim.x) {
        const int t = i / kPctCols, c = i % kPctCols;
        prefix[t][c] = 0ULL;
        const double h = pct_virtual_index(S, q[t]);
        long long k = (long long)floor(h);
        if (k < 0) k = 0;
        if (k > S - 1) k = S - 1;
        rank[t][c] = k;
    }
    __syncthreads();
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        for (int i = threadIdx.x; i < nq * kPctCols * 256; i += blockDim.x) (&hist[0][0][0])[i] = 0u;
        __syncthreads();
        if (live) {
            unsigned long long want[NQ];  // leading digits of the class that holds the wanted rank
#pragma unroll
            for (int t = 0; t < nq; ++t) want[t] = pass == 0 ? 0ULL : prefix[t][tc] >> (shift + 8);
            // four independent rows in flight per thread
            for (long long r0 = tr; r0 < S; r0 += 4 * kPctRows) {
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long r = r0 + (long long)u * kPctRows;
                    v[u] = r < S ? x[r * M + col] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (r0 + (long long)u * kPctRows >= S) break;
                    const unsigned long long k = pct_key(v[u]);
                    if (pass == 0 && v[u] != v[u]) nan_col[tc] = 1u;  // (benign race: every writer stores 1)
                    const unsigned long long lead = pass == 0 ? 0ULL : k >> (shift + 8);
#pragma unroll
                    for (int t = 0; t < nq; ++t)
                        if (lead == want[t]) atomicAdd(&hist[t][tc][(unsigned)(k >> shift) & 255u], 1u);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nq * kPctCols; i += blockDim.x) {
            const int t = i / kPctCols, c = i % kPctCols;
            long long k = rank[t][c];
            int d = 0;
            for (; d < 255; ++d) {
                const long long n = hist[t][c][d];
                if (k < n) break;
                k -= n;
            }
            rank[t][c] = k;
            prefix[t][c] |= (unsigned long long)d << shift;
        }
        __syncthreads();
    }
    // upper neighbour: the same value if enough duplicates follow, else the smallest larger element
    for (int i = threadIdx.x; i < nq * kPctCols; i += blockDim.x) {
        (&above[0][0])[i] = ~0ULL;
        (&n_le[0][0])[i] = 0ULL;
    }
    __syncthreads();
    if (live) {
        // per-thread partial results first: one shared-memory atomic per thread and target at the end
        unsigned long long sel[NQ], cnt[NQ], mn[NQ];
#pragma unroll
        for (int t = 0; t < nq; ++t) {
            sel[t] = prefix[t][tc];
            cnt[t] = 0ULL;
            mn[t] = ~0ULL;
        }
        for (long long r0 = tr; r0 < S; r0 += 4 * kPctRows) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long r = r0 + (long long)u * kPctRows;
                v[u] = r < S ? x[r * M + col] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (r0 + (long long)u * kPctRows >= S) break;
                const unsigned long long k = pct_key(v[u]);
#pragma unroll
                for (int t = 0; t < nq; ++t) {
                    if (k <= sel[t]) ++cnt[t];
                    else if (k < mn[t]) mn[t] = k;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < nq; ++t) {
            atomicAdd(&n_le[t][tc], cnt[t]);
            atomicMin(&above[t][tc], mn[t]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nq * kPctCols; i += blockDim.x) {
        const int t = i / kPctCols, c = i % kPctCols;
        const long long cc = (long long)blockIdx.x * kPctCols + c;
        if (cc >= M) continue;
        const double h = pct_virtual_index(S, q[t]);
        long long k = (long long)floor(h);
        if (k < 0) k = 0;
        if (k > S - 1) k = S - 1;
        const double g = __dadd_rn(h, -(double)k);
        const double a = pct_val(prefix[t][c]);
        // rank k + 1 (0-based) is still `a` if at least k + 2 elements are <= a
        const bool dup = (long long)n_le[t][c] >= k + 2 || k + 1 > S - 1;
        const double b = dup ? a : pct_val(above[t][c]);
        // numpy's _lerp: a + (b - a) g, taken from the other end for g >= 0.5
        const double diff = __dadd_rn(b, -a);
        double v = __dadd_rn(a, __dmul_rn(diff, g));
        if (g >= 0.5) v = __dadd_rn(b, -__dmul_rn(diff, __dadd_rn(1.0, -g)));
        if (nan_col[c]) v = __longlong_as_double(0x7ff8000000000000LL);
        out[(size_t)t * M + cc] = v;
    }
}

}  // namespace jitr

using namespace jitr;

extern "C" int jitr_b200_percentiles(const double* values, long long n_samples, long long n_columns, const double* q_host, int nq,
                                     double* out, void* stream) {
    if (n_samples < 1 || n_columns < 0 || nq < 1 || nq > kPctMaxQ || !values || !q_host || !out)
        return set_error(JITR_B200_ERR_BAD_ARG, "percentiles: null pointer or bad size (1..8 quantiles)");
    if (n_columns == 0) return JITR_B200_OK;
    PctQ Q{};
    for (int t = 0; t < nq; ++t) {
        if (!(q_host[t] >= 0.0 && q_host[t] <= 100.0)) return set_error(JITR_B200_ERR_BAD_ARG, "percentiles: q must be in [0, 100]");
        Q.q[t] = q_host[t];
    }
    const size_t smem = (size_t)nq * kPctCols * (4 * 8 + 256 * 4);
    const unsigned blocks = (unsigned)((n_columns + kPctCols - 1) / kPctCols);
    bool ok = true;
    auto launch = [&](auto kern) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            ok = false;
            return;
        }
        kern<<<blocks, kPctCols * kPctRows, smem, static_cast<cudaStream_t>(stream)>>>(values, n_samples, n_columns, Q, out);
    };
    switch (nq) {
        case 1: launch(percentile_kernel<1>); break;
        case 2: launch(percentile_kernel<2>); break;
        case 3: launch(percentile_kernel<3>); break;
        case 4: launch(percentile_kernel<4>); break;
        case 5: launch(percentile_kernel<5>); break;
        case 6: launch(percentile_kernel<6>); break;
        case 7: launch(percentile_kernel<7>); break;
        default: launch(percentile_kernel<8>); break;
    }
    if (!ok) return set_cuda_error("cudaFuncSetAttribute(percentiles)");
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("percentiles launch");
    return JITR_B200_OK;
}
