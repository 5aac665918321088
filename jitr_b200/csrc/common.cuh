// Shared host-side helpers of the C-ABI layer (error text, SM count, launch counter).
#pragma once
#include <cuda_runtime.h>

namespace jitr {
int set_error(int code, const char* msg);
int set_cuda_error(const char* where);
int sm_count();
void count_launch();
int option(int key);                    // jitr_b200_set_option values
unsigned long long* fallback_counter();
unsigned long long* fallback_mask();    // jitr_b200_debug_set_fallback_mask (nullptr: off)
// a zeroed (stream-ordered) device counter for one launch; a ring of 256, so launches in flight on
// different streams do not share one
unsigned long long* next_work_counter(cudaStream_t st);
// per-device scratch for the pairs a packed sweep mode defers: [0] = count (zeroed stream-ordered here), [4 ..] = list of
// at least n_entries ints; grown on demand (a ring of 4 so that sweeps in flight on different streams do not share one)
int* defer_buffer(long long n_entries, cudaStream_t st);  // device counter (lazily allocated), nullptr on failure
}  // namespace jitr
