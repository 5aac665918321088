// Library-level entry points: version, device check, error text, launch counter.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "common.cuh"
#include "jitr_b200.h"

namespace jitr {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int set_cuda_error(const char* where) {
    cudaError_t e = cudaGetLastError();
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return JITR_B200_ERR_CUDA;
}
int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 1;
    }
    return sms;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace jitr

extern "C" const char* jitr_b200_last_error(void) { return jitr::g_err; }
extern "C" int jitr_b200_version(void) { return 100; }
extern "C" long long jitr_b200_launch_count(void) { return jitr::g_launches.load(); }
extern "C" int jitr_b200_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
        cudaGetLastError();
        return 0;
    }
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    return major == 10 ? 1 : 0;
}
