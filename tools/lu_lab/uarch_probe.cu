// uarch_probe: micro-benchmarks that decide the design of the warp-level LU kernel (sm_100a).
// Development tool.  Prints one JSON object.
//  * instruction-cache capacity: straight-line FP32 code of growing size in a loop, 8 warps/SM
//  * issue throughput per SM (warp-instructions / clk) of SHFL, LDS.128 broadcast, LDS.128 strided,
//    STS.128 one-lane, DFMA, DMMA, and DFMA while LDS broadcast streams
//  * dependent latencies: DFMA, DMMA (accumulator chain), MUFU.RCP64H+Newton, LDS, SHFL, REDUX
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ double2 lds128(const double2* p) {
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return r;
}
__device__ __forceinline__ double lds64(const double* p) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return r;
}
__device__ __forceinline__ void sts128(double2* p, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"((unsigned)__cvta_generic_to_shared(p)), "d"(v.x), "d"(v.y) : "memory");
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s @%d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

// ---------------- i-cache: KI thousands of FFMA instructions, straight line ----------------
template <int KI>
__global__ void __launch_bounds__(256) icache_kernel(float* out, int iters, float s) {
    float a0 = s, a1 = s + 1, a2 = s + 2, a3 = s + 3, a4 = s + 4, a5 = s + 5, a6 = s + 6, a7 = s + 7;
    const float m = 1.0001f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < KI * 128; ++k) {  // 8 FFMA per k -> KI*1024 instructions
            a0 = fmaf(a0, m, s); a1 = fmaf(a1, m, s); a2 = fmaf(a2, m, s); a3 = fmaf(a3, m, s);
            a4 = fmaf(a4, m, s); a5 = fmaf(a5, m, s); a6 = fmaf(a6, m, s); a7 = fmaf(a7, m, s);
        }
    }
    long long t1 = clock64();
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0 && blockIdx.x == 0) ((long long*)out)[1] = t1 - t0;
    if (r == 123.456f) out[0] = r;
}

// ---------------- throughput kernels: every warp issues `iters` x UNR instructions ----------------
enum { T_SHFL, T_LDS_BCAST, T_LDS_ROW, T_STS_ONE, T_DFMA, T_DMMA, T_DFMA_LDS, T_LDS64_BCAST, T_DFMA_SHFL, T_NK };

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int KIND>
__global__ void __launch_bounds__(256) tput_kernel(double* out, int iters, double seed, long long* cyc) {
    __shared__ __align__(16) double2 sm[32 * 41];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32 * 41; i += blockDim.x) sm[i] = make_double2(seed + i, seed - i);
    __syncthreads();
    double acc[8];
    double2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[k] = seed + k; v[k] = make_double2(seed, seed); }
    const double y = 0.999999, x = seed;
    int idx = (lane * 7) & 31;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (KIND == T_SHFL) {
                int lo = __double2loint(acc[k]);
                lo = __shfl_sync(0xffffffffu, lo, idx);
                acc[k] = __hiloint2double(__double2hiint(acc[k]), lo);
            } else if (KIND == T_LDS_BCAST) {
                const double2 t = lds128(&sm[(it + k) & 1023]);
                v[k].x += t.x; v[k].y = t.y;
            } else if (KIND == T_LDS64_BCAST) {
                const double t = lds64(&sm[(it + k) & 1023].x);
                v[k].x = t;
            } else if (KIND == T_LDS_ROW) {
                const double2 t = lds128(&sm[lane * 41 + ((it + k) & 31)]);
                v[k].x += t.x; v[k].y = t.y;
            } else if (KIND == T_STS_ONE) {
                if (lane == (it & 31)) sts128(&sm[(it + k) & 1023], v[k]);
            } else if (KIND == T_DFMA) {
                acc[k] = fma(acc[k], y, x);
            } else if (KIND == T_DMMA) {
                dmma884(acc[k], v[k].x, x, y);
            } else if (KIND == T_DFMA_LDS) {   // 4 DFMA : 1 broadcast LDS.128
                acc[k] = fma(acc[k], y, x);
                if ((k & 3) == 0) { const double2 t = lds128(&sm[(it + k) & 1023]); v[k].y = t.y; }
            } else if (KIND == T_DFMA_SHFL) {  // 4 DFMA : 2 SHFL
                acc[k] = fma(acc[k], y, x);
                if ((k & 1) == 0) { int lo = __double2loint(v[k].y); lo = __shfl_sync(0xffffffffu, lo, idx); v[k].y = __hiloint2double(0, lo); }
            }
        }
    }
    long long t1 = clock64();
    double r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += acc[k] + v[k].x + v[k].y;
    if (r == 123.456) out[0] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// ---------------- latency kernels (one warp) ----------------
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__global__ void latency_kernel(long long* out, double seed) {
    __shared__ __align__(16) double2 sm[64];
    const int lane = threadIdx.x;
    sm[lane] = make_double2(0.0, 0.0);
    sm[lane + 32] = make_double2(0.0, 0.0);
    __syncwarp();
    const int n = 1024;
    double a = seed, b = 0.9999999;
    long long t0, t1;
    t0 = clock64();
    for (int i = 0; i < n; ++i) a = fma(a, b, seed);
    t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    double c0 = seed, c1 = seed;
    t0 = clock64();
    for (int i = 0; i < n; ++i) dmma884(c0, c1, b, b);
    t1 = clock64();
    if (lane == 0) out[1] = t1 - t0;
    double r = seed + 2.0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) r = rcp_fast(r) + 1.5;
    t1 = clock64();
    if (lane == 0) out[2] = t1 - t0;
    double d = seed + 2.0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) d = 1.0 / d + 1.5;
    t1 = clock64();
    if (lane == 0) out[3] = t1 - t0;
    int p = 0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) p = __double2loint(lds128(&sm[p & 63]).x);
    t1 = clock64();
    if (lane == 0) out[4] = t1 - t0;
    int q = lane;
    t0 = clock64();
    for (int i = 0; i < n; ++i) q = __shfl_sync(0xffffffffu, q, (q + 1) & 31);
    t1 = clock64();
    if (lane == 0) out[5] = t1 - t0;
    int m = lane;
    t0 = clock64();
    for (int i = 0; i < n; ++i) m = __reduce_max_sync(0xffffffffu, m) - lane;
    t1 = clock64();
    if (lane == 0) out[6] = t1 - t0;
    // STS (one lane) -> syncwarp -> LDS broadcast round trip
    int w = 0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (lane == (w & 31)) sts128(&sm[i & 63], make_double2(__hiloint2double(0, w + 1), 0.0));
        __syncwarp();
        w = __double2loint(lds128(&sm[i & 63]).x);
    }
    t1 = clock64();
    if (lane == 0) out[7] = t1 - t0;
    if (a + c0 + c1 + r + d + p + q + m + w == 123.456) out[15] = 1;
}

template <int KI>
static double run_icache(int sms) {
    float* d;
    CK(cudaMalloc(&d, 64));
    const int iters = 16384 / KI < 8 ? 8 : 16384 / KI / 8;
    icache_kernel<KI><<<sms, 256>>>(d, iters, 0.5f);
    CK(cudaDeviceSynchronize());
    icache_kernel<KI><<<sms, 256>>>(d, iters, 0.5f);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    // warp-instructions per clk per SM: 8 warps * KI*1024*iters / cycles
    return 8.0 * KI * 1024.0 * iters / (double)h[1];
}

template <int KIND>
static double run_tput(int sms, int warps) {
    double* d;
    long long* c;
    CK(cudaMalloc(&d, 64));
    CK(cudaMalloc(&c, 64));
    const int iters = 4096;
    tput_kernel<KIND><<<sms, warps * 32>>>(d, iters, 1.25, c);
    CK(cudaDeviceSynchronize());
    tput_kernel<KIND><<<sms, warps * 32>>>(d, iters, 1.25, c);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    CK(cudaFree(c));
    return (double)warps * 8.0 * iters / (double)h;  // "k-steps" per clk per SM
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("{\"icache_ffma_warp_inst_per_clk_per_sm\": {");
    printf("\"16KB\": %.3f, ", run_icache<1>(sms));
    printf("\"32KB\": %.3f, ", run_icache<2>(sms));
    printf("\"48KB\": %.3f, ", run_icache<3>(sms));
    printf("\"64KB\": %.3f, ", run_icache<4>(sms));
    printf("\"96KB\": %.3f, ", run_icache<6>(sms));
    printf("\"128KB\": %.3f, ", run_icache<8>(sms));
    printf("\"192KB\": %.3f, ", run_icache<12>(sms));
    printf("\"256KB\": %.3f}, ", run_icache<16>(sms));
    for (int warps : {8, 16}) {
        printf("\"ksteps_per_clk_per_sm_%dwarps\": {", warps);
        printf("\"shfl32\": %.3f, ", run_tput<T_SHFL>(sms, warps));
        printf("\"lds128_bcast\": %.3f, ", run_tput<T_LDS_BCAST>(sms, warps));
        printf("\"lds64_bcast\": %.3f, ", run_tput<T_LDS64_BCAST>(sms, warps));
        printf("\"lds128_row_per_lane\": %.3f, ", run_tput<T_LDS_ROW>(sms, warps));
        printf("\"sts128_one_lane\": %.3f, ", run_tput<T_STS_ONE>(sms, warps));
        printf("\"dfma\": %.3f, ", run_tput<T_DFMA>(sms, warps));
        printf("\"dmma\": %.3f, ", run_tput<T_DMMA>(sms, warps));
        printf("\"dfma_with_1lds_per_4\": %.3f, ", run_tput<T_DFMA_LDS>(sms, warps));
        printf("\"dfma_with_2shfl_per_4\": %.3f}, ", run_tput<T_DFMA_SHFL>(sms, warps));
    }
    long long* dl;
    CK(cudaMalloc(&dl, 128));
    latency_kernel<<<1, 32>>>(dl, 1.25);
    CK(cudaDeviceSynchronize());
    long long h[16];
    CK(cudaMemcpy(h, dl, 128, cudaMemcpyDeviceToHost));
    printf("\"latency_cycles\": {\"dfma\": %.1f, \"dmma_acc_chain\": %.1f, \"rcp_fast+dadd\": %.1f, \"ddiv+dadd\": %.1f, "
           "\"lds\": %.1f, \"shfl\": %.1f, \"redux+iadd\": %.1f, \"sts_sync_lds\": %.1f}}\n",
           h[0] / 1024.0, h[1] / 1024.0, h[2] / 1024.0, h[3] / 1024.0, h[4] / 1024.0, h[5] / 1024.0, h[6] / 1024.0, h[7] / 1024.0);
    return 0;
}
