// mio_probe: throughput of shared-memory / shuffle instruction kinds on sm_100a (development tool).
// Each of W warps per SM issues ITERS x 8 instructions of one kind; prints SM-cycles per warp-instruction.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds128(unsigned a) { double2 r; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a)); return r; }
__device__ __forceinline__ double lds64(unsigned a) { double r; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a)); return r; }
__device__ __forceinline__ void sts128(unsigned a, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts64_pred(unsigned a, double v, int p) {
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.f64 [%0], %1; }" ::"r"(a), "d"(v), "r"(p) : "memory");
}
__device__ __forceinline__ void sts128_pred(unsigned a, double2 v, int p) {
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q st.shared.v2.f64 [%0], {%1,%2}; }" ::"r"(a), "d"(v.x), "d"(v.y), "r"(p) : "memory");
}

enum { K_LDS128_CONTIG, K_LDS128_ROW41, K_LDS128_BCAST, K_STS128_CONTIG, K_STS128_ONE, K_SHFL, K_LDS64_CONTIG,
       K_SHFL_PLUS_LDS, K_DFMA_PLUS_SHFL, K_DFMA_PLUS_LDSB, K_STS128_ROW41, K_LDS64_BCAST, K_DFMA, K_SHFL_IND, K_STS64_CONTIG, K_STS64_ONE, K_STS128_HALF, K_DFMA4_SHFL_IND, K_SHFL_IND_LDS, K_NK };

template <int KIND>
__global__ void __launch_bounds__(512) k(double* out, int iters, double seed, long long* cyc) {
    extern __shared__ __align__(16) double2 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* my = sm + warp * 41 * 41;  // per-warp 41x41 region
    for (int i = lane; i < 41 * 41; i += 32) my[i] = make_double2(seed + i, seed - i);
    __syncthreads();
    const unsigned base = saddr(my);
    double2 v[8];
    double acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = make_double2(seed + j, seed); acc[j] = seed + j; }
    int sh = lane;
    int shv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) shv[j] = lane * 3 + j;
    const int src = (lane * 5 + 3) & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int r = it & 31;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = (r + j) & 31;
            if (KIND == K_LDS128_CONTIG) { double2 t = lds128(base + (c * 41 + lane) * 16); v[j].x += t.x; v[j].y += t.y; }
            if (KIND == K_LDS128_ROW41) { double2 t = lds128(base + (lane * 41 + c) * 16); v[j].x += t.x; v[j].y += t.y; }
            if (KIND == K_LDS128_BCAST) { double2 t = lds128(base + (c * 41 + j) * 16); v[j].x += t.x; v[j].y += t.y; }
            if (KIND == K_LDS64_BCAST) { double t = lds64(base + (c * 41 + j) * 16); v[j].x += t; }
            if (KIND == K_LDS64_CONTIG) { double t = lds64(base + (c * 41) * 16 + lane * 8); v[j].x += t; }
            if (KIND == K_STS128_CONTIG) sts128(base + (c * 41 + lane) * 16, v[j]);
            if (KIND == K_STS128_ROW41) sts128(base + (lane * 41 + c) * 16, v[j]);
            if (KIND == K_STS128_ONE) sts128_pred(base + (c * 41 + j) * 16, v[j], lane == c);
            if (KIND == K_SHFL) { sh = __shfl_sync(0xffffffffu, sh + j, src); }
            if (KIND == K_SHFL_PLUS_LDS) { sh = __shfl_sync(0xffffffffu, sh + j, src); double2 t = lds128(base + (c * 41 + j) * 16); v[j].x += t.x; v[j].y += t.y; }
            if (KIND == K_DFMA_PLUS_SHFL) { sh = __shfl_sync(0xffffffffu, sh + j, src); acc[j] = fma(acc[j], 0.99999, seed); v[j].x = fma(v[j].x, 0.99999, seed); v[j].y = fma(v[j].y, 0.99999, seed); acc[(j + 4) & 7] = fma(acc[(j + 4) & 7], 0.99999, seed); }
            if (KIND == K_DFMA_PLUS_LDSB) { double2 t = lds128(base + (c * 41 + j) * 16); acc[j] = fma(acc[j], 0.99999, t.x); v[j].x = fma(v[j].x, 0.99999, seed); v[j].y = fma(v[j].y, 0.99999, seed); acc[(j + 4) & 7] = fma(acc[(j + 4) & 7], 0.99999, seed); }
            if (KIND == K_SHFL_IND) { shv[j] = __shfl_sync(0xffffffffu, shv[j], src); }
            if (KIND == K_STS64_CONTIG) sts64(base + (c * 41) * 16 + lane * 8, v[j].x);
            if (KIND == K_STS64_ONE) sts64_pred(base + (c * 41 + j) * 16, v[j].x, lane == c);
            if (KIND == K_STS128_HALF) sts128_pred(base + (c * 41 + lane) * 16, v[j], lane < 16);
            if (KIND == K_DFMA4_SHFL_IND) { shv[j] = __shfl_sync(0xffffffffu, shv[j], src); acc[j] = fma(acc[j], 0.99999, seed); v[j].x = fma(v[j].x, 0.99999, seed); v[j].y = fma(v[j].y, 0.99999, seed); acc[(j + 4) & 7] = fma(acc[(j + 4) & 7], 0.99999, seed); }
            if (KIND == K_SHFL_IND_LDS) { shv[j] = __shfl_sync(0xffffffffu, shv[j], src); double2 t = lds128(base + (c * 41 + lane) * 16); v[j].x += t.x; v[j].y += t.y; }
            if (KIND == K_DFMA) { acc[j] = fma(acc[j], 0.99999, seed); v[j].x = fma(v[j].x, 0.99999, seed); v[j].y = fma(v[j].y, 0.99999, seed); }
        }
    }
    long long t1 = clock64();
    double s = sh;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += shv[j];
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j].x + v[j].y + acc[j];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int KIND>
static double run(int sms, int warps) {
    double* d; long long* c;
    CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&c, 64));
    const int iters = 2048;
    const int smem = warps * 41 * 41 * 16;
    CK(cudaFuncSetAttribute(k<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int w = 0; w < 2; ++w) { k<KIND><<<sms, warps * 32, smem>>>(d, iters, 1.25, c); CK(cudaDeviceSynchronize()); }
    long long h; CK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost));
    CK(cudaFree(d)); CK(cudaFree(c));
    return (double)h / ((double)warps * 8.0 * iters);  // SM-cycles per (warp, j-step)
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    for (int warps : {8, 4}) {
        printf("{\"warps_per_sm\": %d, \"sm_cycles_per_step\": {", warps);
        printf("\"lds128_contig\": %.2f, ", run<K_LDS128_CONTIG>(sms, warps));
        printf("\"lds128_row41\": %.2f, ", run<K_LDS128_ROW41>(sms, warps));
        printf("\"lds128_bcast\": %.2f, ", run<K_LDS128_BCAST>(sms, warps));
        printf("\"lds64_bcast\": %.2f, ", run<K_LDS64_BCAST>(sms, warps));
        printf("\"lds64_contig\": %.2f, ", run<K_LDS64_CONTIG>(sms, warps));
        printf("\"sts128_contig\": %.2f, ", run<K_STS128_CONTIG>(sms, warps));
        printf("\"sts128_row41\": %.2f, ", run<K_STS128_ROW41>(sms, warps));
        printf("\"sts128_one_lane_pred\": %.2f, ", run<K_STS128_ONE>(sms, warps));
        printf("\"shfl\": %.2f, ", run<K_SHFL>(sms, warps));
        printf("\"shfl+lds128_bcast\": %.2f, ", run<K_SHFL_PLUS_LDS>(sms, warps));
        printf("\"shfl_indep\": %.2f, ", run<K_SHFL_IND>(sms, warps));
        printf("\"sts64_contig\": %.2f, ", run<K_STS64_CONTIG>(sms, warps));
        printf("\"sts64_one_lane\": %.2f, ", run<K_STS64_ONE>(sms, warps));
        printf("\"sts128_half_warp\": %.2f, ", run<K_STS128_HALF>(sms, warps));
        printf("\"dfma4+shfl_indep\": %.2f, ", run<K_DFMA4_SHFL_IND>(sms, warps));
        printf("\"shfl_indep+lds128_contig\": %.2f, ", run<K_SHFL_IND_LDS>(sms, warps));
        printf("\"dfma3\": %.2f, ", run<K_DFMA>(sms, warps));
        printf("\"dfma4+shfl\": %.2f, ", run<K_DFMA_PLUS_SHFL>(sms, warps));
        printf("\"dfma4+lds128_bcast\": %.2f}}\n", run<K_DFMA_PLUS_LDSB>(sms, warps));
    }
    return 0;
}
