// Shared host-side helpers of the C-ABI layer (error text, SM count, launch counter).
#pragma once
#include <cuda_runtime.h>

// nbasis values the fused warp-per-system sweep is compiled for (N <= 63: two row slots per lane)
#define JITR_SWEEP_SIZES(X) X(20) X(30) X(40) X(50) X(60)

namespace jitr {
int set_error(int code, const char* msg);
int set_cuda_error(const char* where);
int sm_count();
void count_launch();
}  // namespace jitr
