// lu_lab: stand-alone timing / correctness harness for the warp-level bordered LU (csrc/lu_warp.cuh).
// Development tool, not product code.   nvcc ... -DLU_HEADER='"path/to/lu_warp.cuh"' [-DJITR_PHASE_TIMERS]
//   ./lu_lab <lab_data_N40.bin> [systems_per_warp=64] [warps_per_cta=0(auto)]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#ifdef JITR_PHASE_TIMERS
__device__ unsigned long long g_phase[8];
#define JITR_PH_INIT long long ph_last = clock64();
#define JITR_PH(k)                                                              \
    {                                                                           \
        const long long ph_t = clock64();                                       \
        if (lane == 0) atomicAdd(&g_phase[k], (unsigned long long)(ph_t - ph_last)); \
        ph_last = clock64();                                                    \
    }
#endif

#include LU_HEADER
#ifdef LAB_FAST
#include "lu_fast.cuh"
#endif
#ifdef LAB_SYM
#include "lu_sym.cuh"
#endif

using namespace jitr;

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                               \
        }                                                                          \
    } while (0)

struct LabArgs {
    int N, K;
    long long B;
    const double* That;
    const double* bvec;
    const double2* diag;
    double2* out;
    int* sing;
};

__global__ void __launch_bounds__(256, 1) lab_kernel(const LabArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = A.N;
    LuGeom g;
    g.N = N;
    g.P = (N + 7) & ~7;
    g.LD = g.P + 1;
    const int P = g.P, LD = g.LD, msz = LD * LD;
    const int nwarps = blockDim.x >> 5;
    double2* Mall = reinterpret_cast<double2*>(smem_raw);
    double* Ts = reinterpret_cast<double*>(smem_raw + (size_t)nwarps * msz * 16);
    // broadcast from lane 0: tells the compiler the warp index (hence the whole work loop) is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    double2* M = Mall + (size_t)warp * msz;
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) Ts[i] = A.That[i];
    __syncthreads();
    const int n0 = lane, n1 = lane + 32;
    const bool v0 = n0 < N, v1 = n1 < N;
    const double b0 = A.bvec[v0 ? n0 : 0], b1 = A.bvec[v1 ? n1 : 0];
    const double2 zero2 = make_double2(0.0, 0.0);
    int singular = 0;
    for (long long sys = (long long)blockIdx.x * nwarps + warp; sys < A.B; sys += (long long)gridDim.x * nwarps) {
        const double2* dg = A.diag + (size_t)(sys % A.K) * N;
        const double2 d0 = dg[v0 ? n0 : 0], d1 = dg[v1 ? n1 : 0];
#ifdef JITR_PHASE_TIMERS
        long long ph_last = clock64();
#endif
        for (int i = 0; i < N; ++i) {
            const double* Trow = Ts + i * N;
            double2* Mrow = M + i * LD;
            if (lane < P) Mrow[lane] = make_double2(v0 ? Trow[n0] : 0.0, 0.0);
            if (n1 < P) Mrow[n1] = make_double2(v1 ? Trow[n1] : 0.0, 0.0);
        }
        for (int i = N; i < P; ++i)
            for (int j = lane; j < LD; j += 32) M[i * LD + j] = make_double2(i == j ? 1.0 : 0.0, 0.0);
        __syncwarp();
        if (v0) {
            M[n0 * LD + n0] = d0;
            M[n0 * LD + P] = make_double2(b0, 0.0);
            M[P * LD + n0] = make_double2(b0, 0.0);
        }
        if (v1) {
            M[n1 * LD + n1] = d1;
            M[n1 * LD + P] = make_double2(b1, 0.0);
            M[P * LD + n1] = make_double2(b1, 0.0);
        }
        for (int j = N + lane; j <= P; j += 32) M[P * LD + j] = zero2;
        __syncwarp();
#ifdef JITR_PHASE_TIMERS
        JITR_PH(3)
#endif
        #ifdef LAB_OLD_API
        const double2 corner = bordered_schur(M, g, lane, singular);
#else
        const double2 corner = bordered_schur<2>(M, g, lane, singular);
#endif
        __syncwarp();
        if (lane == 0) A.out[sys] = corner;
    }
    if (lane == 0 && singular) atomicAdd(A.sing, 1);
}

#ifdef LAB_FAST
template <int PT>
__global__ void __launch_bounds__(384, 1) lab_fast_kernel(const LabArgs A) {
    using G = FastGeom<PT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = A.N;
    const int nwarps = blockDim.x >> 5;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    double* Tp = reinterpret_cast<double*>(smem_raw);
    double* bsh = Tp + PT * PT;
    unsigned char* wbase = smem_raw + (PT * PT + PT) * 8 + (size_t)warp * G::WARP_BYTES;
    double2* W = reinterpret_cast<double2*>(wbase);
    double2* dvec = W + G::WALLOC;
    unsigned char* origv = reinterpret_cast<unsigned char*>(dvec + PT);
    for (int i = threadIdx.x; i < PT * PT; i += blockDim.x) {
        const int r = i / PT, c = i - r * PT;
        Tp[i] = (r < N && c < N) ? A.That[r * N + c] : (r == c ? 1.0 : 0.0);
    }
    for (int i = threadIdx.x; i < PT; i += blockDim.x) bsh[i] = i < N ? A.bvec[i] : 0.0;
    __syncthreads();
    int singular = 0;
    for (long long sys = (long long)blockIdx.x * nwarps + warp; sys < A.B; sys += (long long)gridDim.x * nwarps) {
        const double2* dg = A.diag + (size_t)(sys % A.K) * N;
        double2 d[G::NS0];
#pragma unroll
        for (int s = 0; s < G::NS0; ++s) {
            const int i = 32 * s + lane;
            d[s] = i < N ? dg[i] : make_double2(1.0, 0.0);
        }
#ifdef LAB_ALIGN
        __syncthreads();  // keep the warps of the CTA in the same code region (instruction-cache experiment)
#endif
        const double2 corner = fast_schur<PT>(W, dvec, origv, Tp, bsh, d, lane, singular);
        __syncwarp();
        if (lane == 0) A.out[sys] = corner;
    }
    if (lane == 0 && singular) atomicAdd(A.sing, 1);
}
#endif

#ifdef LAB_SYM
template <int PT>
__global__ void __launch_bounds__(384, 1) lab_sym_kernel(const LabArgs A) {
    using G = SymGeom<PT>;
    constexpr int WARP_BYTES = G::WALLOC * 16 + PT * 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = A.N;
    const int nwarps = blockDim.x >> 5;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    double* Tp = reinterpret_cast<double*>(smem_raw);
    double* bsh = Tp + PT * PT;
    unsigned char* wbase = smem_raw + (PT * PT + PT) * 8 + (size_t)warp * WARP_BYTES;
    double2* W = reinterpret_cast<double2*>(wbase);
    double2* dvec = W + G::WALLOC;
    for (int i = threadIdx.x; i < PT * PT; i += blockDim.x) {
        const int r = i / PT, c = i - r * PT;
        Tp[i] = (r < N && c < N) ? A.That[r * N + c] : (r == c ? 1.0 : 0.0);
    }
    for (int i = threadIdx.x; i < PT; i += blockDim.x) bsh[i] = i < N ? A.bvec[i] : 0.0;
    __syncthreads();
    int nfail = 0;
    for (long long sys = (long long)blockIdx.x * nwarps + warp; sys < A.B; sys += (long long)gridDim.x * nwarps) {
        const double2* dg = A.diag + (size_t)(sys % A.K) * N;
        double2 d[G::NS0];
#pragma unroll
        for (int s = 0; s < G::NS0; ++s) {
            const int i = 32 * s + lane;
            d[s] = i < N ? dg[i] : make_double2(1.0, 0.0);
        }
#ifdef LAB_ALIGN
        __syncthreads();
#endif
        int fail = 0;
        const double2 corner = sym_schur<PT>(W, dvec, Tp, bsh, d, lane, fail);
        __syncwarp();
        if (lane == 0) A.out[sys] = corner;
        nfail += fail;
    }
    if (lane == 0 && nfail) atomicAdd(A.sing, nfail);
}
#endif

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: lu_lab data.bin [systems_per_warp] [warps_per_cta]\n");
        return 1;
    }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 1; }
    int hdr[2];
    if (fread(hdr, 4, 2, f) != 2) return 1;
    const int N = hdr[0], K = hdr[1];
    std::vector<double> That(N * N), b(N);
    std::vector<double2> diag((size_t)K * N), ref(K);
    if (fread(That.data(), 8, N * N, f) != (size_t)N * N) return 1;
    if (fread(b.data(), 8, N, f) != (size_t)N) return 1;
    if (fread(diag.data(), 16, (size_t)K * N, f) != (size_t)K * N) return 1;
    if (fread(ref.data(), 16, K, f) != (size_t)K) return 1;
    fclose(f);
    const int spw = argc > 2 ? atoi(argv[2]) : 64;
    int warps = argc > 3 ? atoi(argv[3]) : 0;
    const int P = (N + 7) & ~7, LD = P + 1, msz = LD * LD;
    const int maxs = 232448;
    const int fit = (maxs - N * N * 8) / (msz * 16);
    if (warps <= 0) warps = fit > 8 ? 8 : fit;
    if (warps > fit) warps = fit;
#ifdef LAB_FAST
    constexpr int PT = 40;
    if (P != PT) { fprintf(stderr, "LAB_FAST is built for P = %d\n", PT); return 1; }
    warps = argc > 3 ? atoi(argv[3]) : 12;
    const int smem = (PT * PT + PT) * 8 + warps * FastGeom<PT>::WARP_BYTES;
    if (smem > maxs) { fprintf(stderr, "smem %d > %d\n", smem, maxs); return 1; }
#define lab_kernel lab_fast_kernel<PT>
#elif defined(LAB_SYM)
    constexpr int PT = 40;
    if (P != PT) { fprintf(stderr, "LAB_SYM is built for P = %d\n", PT); return 1; }
    warps = argc > 3 ? atoi(argv[3]) : 12;
    const int smem = (PT * PT + PT) * 8 + warps * (SymGeom<PT>::WALLOC * 16 + PT * 16);
    if (smem > maxs) { fprintf(stderr, "smem %d > %d\n", smem, maxs); return 1; }
#define lab_kernel lab_sym_kernel<PT>
#else
    const int smem = warps * msz * 16 + N * N * 8;
#endif
    int dev = 0, sms = 0, clk = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    const long long B = (long long)sms * warps * spw;
    double *dT, *db;
    double2 *dd, *dout;
    int* dsing;
    CK(cudaMalloc(&dT, N * N * 8));
    CK(cudaMalloc(&db, N * 8));
    CK(cudaMalloc(&dd, (size_t)K * N * 16));
    CK(cudaMalloc(&dout, (size_t)B * 16));
    CK(cudaMalloc(&dsing, 4));
    CK(cudaMemset(dsing, 0, 4));
    CK(cudaMemcpy(dT, That.data(), N * N * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b.data(), N * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dd, diag.data(), (size_t)K * N * 16, cudaMemcpyHostToDevice));
    LabArgs a{N, K, B, dT, db, dd, dout, dsing};
    CK(cudaFuncSetAttribute(lab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) lab_kernel<<<sms, warps * 32, smem>>>(a);
    CK(cudaDeviceSynchronize());
#ifdef JITR_PHASE_TIMERS
    unsigned long long zero[8] = {0};
    CK(cudaMemcpyToSymbol(g_phase, zero, sizeof(zero)));
#endif
    const int reps = 3;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) lab_kernel<<<sms, warps * 32, smem>>>(a);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    std::vector<double2> out(B);
    int sing = 0;
    CK(cudaMemcpy(out.data(), dout, (size_t)B * 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&sing, dsing, 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (long long i = 0; i < B; ++i) {
        const double2 r = ref[i % K], o = out[i];
        const double e = hypot(o.x - r.x, o.y - r.y) / hypot(r.x, r.y);
        if (!(e <= maxerr)) maxerr = e;  // NaN propagates
    }
    const double flops = 8.0 * N * N * N / 3 + 8.0 * N * N;
    const double solves_s = B / (ms * 1e-3);
    const double cyc_per_solve_warp = (double)ms * 1e-3 * 1.965e9 / spw;
    printf("{\"N\": %d, \"warps_per_cta\": %d, \"B\": %lld, \"ms\": %.3f, \"solves_per_s\": %.4g, \"tflops\": %.3f, "
           "\"cycles_per_solve_per_warp_at_1965MHz\": %.0f, \"max_rel_err\": %.3g, \"singular_flags\": %d",
           N, warps, B, ms, solves_s, solves_s * flops / 1e12, cyc_per_solve_warp, maxerr, sing);
#ifdef JITR_PHASE_TIMERS
    unsigned long long ph[8];
    CK(cudaMemcpyFromSymbol(ph, g_phase, sizeof(ph)));
    const double den = (double)B * reps;
    printf(", \"phase_cycles_per_solve\": {\"P\": %.0f, \"U\": %.0f, \"T\": %.0f, \"assemble\": %.0f}", ph[0] / den, ph[1] / den,
           ph[2] / den, ph[3] / den);
#endif
    printf("}\n");
    return (maxerr < 1e-9) ? 0 : 3;
}
