// Library-level entry points: version, device check, error text, launch counter.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>

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
// Per-device state (one process may drive several devices from several threads): indexed by the current device.
constexpr int kMaxDevices = 64;
static std::mutex g_dev_mutex;
static int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
    return dev;
}
int sm_count() {
    static std::atomic<int> sms[kMaxDevices];
    const int dev = current_device();
    int v = sms[dev].load();
    if (!v) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 1;
        sms[dev].store(v);
    }
    return v;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static std::atomic<int> g_options[JITR_B200_OPT_COUNT] = {{1}, {0}, {0}, {1}};
int option(int key) { return (key >= 0 && key < JITR_B200_OPT_COUNT) ? g_options[key].load() : 0; }
unsigned long long* fallback_counter() {
    static unsigned long long* ctr[kMaxDevices] = {nullptr};
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    if (!ctr[dev]) {
        if (cudaMalloc(&ctr[dev], sizeof(unsigned long long)) != cudaSuccess) {
            cudaGetLastError();
            ctr[dev] = nullptr;
            return nullptr;
        }
        cudaMemset(ctr[dev], 0, sizeof(unsigned long long));
    }
    return ctr[dev];
}
static std::atomic<unsigned long long*> g_fallback_mask{nullptr};
unsigned long long* fallback_mask() { return g_fallback_mask.load(); }
unsigned long long* next_work_counter(cudaStream_t st) {
    constexpr int kRing = 256;
    static unsigned long long* ring[kMaxDevices] = {nullptr};
    static std::atomic<unsigned> seq{0};
    const int dev = current_device();
    {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        if (!ring[dev]) {
            if (cudaMalloc(&ring[dev], kRing * sizeof(unsigned long long)) != cudaSuccess) {
                ring[dev] = nullptr;
                return nullptr;
            }
        }
    }
    unsigned long long* c = ring[dev] + (seq.fetch_add(1) % kRing);
    if (cudaMemsetAsync(c, 0, sizeof(unsigned long long), st) != cudaSuccess) return nullptr;
    return c;
}
int* defer_buffer(long long n_entries, cudaStream_t st) {
    constexpr int kRing = 4;
    static int* buf[kMaxDevices][kRing] = {};
    static long long cap[kMaxDevices][kRing] = {};
    static std::atomic<unsigned> seq{0};
    const int dev = current_device();
    const int slot = (int)(seq.fetch_add(1) % kRing);
    int* p = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        if (cap[dev][slot] < n_entries) {
            if (buf[dev][slot]) cudaFree(buf[dev][slot]);  // (synchronises: nothing in flight uses the old block any more)
            buf[dev][slot] = nullptr;
            cap[dev][slot] = 0;
            const long long want = n_entries + n_entries / 4 + 1024;
            if (cudaMalloc(&buf[dev][slot], (size_t)(want + 4) * sizeof(int)) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
            cap[dev][slot] = want;
        }
        p = buf[dev][slot];
    }
    if (cudaMemsetAsync(p, 0, 4 * sizeof(int), st) != cudaSuccess) return nullptr;
    return p;
}
}  // namespace jitr

extern "C" const char* jitr_b200_last_error(void) { return jitr::g_err; }
extern "C" int jitr_b200_version(void) { return 100; }
extern "C" long long jitr_b200_launch_count(void) { return jitr::g_launches.load(); }
extern "C" int jitr_b200_set_option(int key, int value) {
    if (key < 0 || key >= JITR_B200_OPT_COUNT) return jitr::set_error(JITR_B200_ERR_BAD_ARG, "set_option: unknown key");
    jitr::g_options[key].store(value);
    return JITR_B200_OK;
}
extern "C" int jitr_b200_debug_set_fallback_mask(unsigned long long* mask) {
    jitr::g_fallback_mask.store(mask);
    return JITR_B200_OK;
}
extern "C" long long jitr_b200_sym_fallback_count(void) {
    unsigned long long* c = jitr::fallback_counter();
    if (!c) return -1;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, c, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (long long)v;
}
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
