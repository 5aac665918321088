// FP64 pipe probes: measured DFMA / DMMA peak of the device (roofline denominators).
// MEASURED_PEAKS.json carries HBM and bf16 figures only; the R-matrix solve is FP64-bound,
// so bench.py measures the FP64 peak live with these kernels.
#include <cuda_runtime.h>
#include <stdint.h>
#include "jitr_b200.h"

namespace {

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_probe(double* out, int iters, double seed) {
    double acc[CHAINS];
    double x = seed + 1e-9 * threadIdx.x, y = 0.999999 + 1e-12 * threadIdx.x;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = c * 0.125 + x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], y, x);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) dmma_probe(double* out, int iters, double seed) {
    double c0[CHAINS], c1[CHAINS];
    double a = seed + 1e-9 * threadIdx.x, b = 1e-3 + 1e-12 * threadIdx.x;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { c0[c] = c; c1[c] = -c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) dmma884(c0[c], c1[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += c0[c] + c1[c];
    if (s == 123.456) out[0] = s;
}

// warps alternate: even warps DFMA chains, odd warps DMMA chains -> do the two share a pipe?
template <int CHAINS>
__global__ void __launch_bounds__(256) mixed_probe(double* out, int iters, double seed, int dmma_warps_of_8) {
    int w = threadIdx.x >> 5;
    double s = 0;
    if (w < dmma_warps_of_8) {
        double c0[CHAINS], c1[CHAINS];
        double a = seed + 1e-9 * threadIdx.x, b = 1e-3;
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) { c0[c] = c; c1[c] = -c; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) dmma884(c0[c], c1[c], a, b);
        }
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) s += c0[c] + c1[c];
    } else {
        double acc[CHAINS];
        double x = seed + 1e-9 * threadIdx.x, y = 0.999999;
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = c * 0.125 + x;
        for (int it = 0; it < 8 * iters; ++it) {   // 8 DFMA-instr ~ flops of one DMMA-instr
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], y, x);
        }
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) s += acc[c];
    }
    if (s == 123.456) out[0] = s;
}

// latency probes (single warp): dependent chains, cycles per op
__global__ void latency_probe(long long* out, double seed) {
    double x = seed, y = 0.9999;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) x = fma(x, y, seed);
    long long t1 = clock64();
    unsigned v = threadIdx.x * 2654435761u + (unsigned)(x);
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) v = __reduce_max_sync(0xffffffffu, v + i) ^ threadIdx.x;
    long long t2 = clock64();
    double z = x;
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) z = __shfl_sync(0xffffffffu, z, (threadIdx.x + 1) & 31) + 1.0;
    long long t3 = clock64();
    double r = z;
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) r = 1.0 / (r + 3.0);
    long long t4 = clock64();
    if (threadIdx.x == 0) {
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3;
        out[4] = (long long)(x + v + z + r);
    }
}

}  // namespace

extern "C" int jitr_b200_fp64_probe(int which, int iters, int warps_dmma, double* tflops_out, long long* lat_out) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return JITR_B200_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* d = nullptr;
    if (cudaMalloc(&d, 64) != cudaSuccess) return JITR_B200_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int CH = 8;
    int blocks = sms * 8, threads = 256;
    float ms = 0;
    double flops = 0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) {
            dfma_probe<CH><<<blocks, threads>>>(d, iters, 1.0);
            flops = 2.0 * CH * (double)iters * blocks * threads;
        } else if (which == 1) {
            dmma_probe<CH><<<blocks, threads>>>(d, iters, 1.0);
            flops = 2.0 * 8 * 8 * 4 * CH * (double)iters * blocks * (threads / 32);
        } else if (which == 2) {
            mixed_probe<CH><<<blocks, threads>>>(d, iters, 1.0, warps_dmma);
            double fd = 2.0 * 8 * 8 * 4 * CH * (double)iters * blocks * warps_dmma;
            double fv = 2.0 * CH * 8.0 * (double)iters * blocks * (threads - 32 * warps_dmma);
            flops = fd + fv;
        } else if (which == 3) {
            long long* dl = nullptr;
            cudaMalloc(&dl, 64);
            latency_probe<<<1, 32>>>(dl, 1.0);
            cudaMemcpy(lat_out, dl, 5 * sizeof(long long), cudaMemcpyDeviceToHost);
            cudaFree(dl);
        }
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return JITR_B200_ERR_CUDA; }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    if (tflops_out) *tflops_out = flops / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return cudaGetLastError() == cudaSuccess ? JITR_B200_OK : JITR_B200_ERR_CUDA;
}
