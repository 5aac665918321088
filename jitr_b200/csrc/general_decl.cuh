// Arguments of the general batched solve, shared by its kernels (general.cu, packed_solve.cu).
#pragma once
#include <cuda_runtime.h>

namespace jitr {

constexpr int kMaxCh = 16;

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
    const int* symflag;            // device int: 0 when the shared free matrix is symmetric (NULL: do not try the symmetric path)
    unsigned long long* fallbacks;  // systems the symmetric path handed to partial pivoting
    // indirection for the systems a faster kernel deferred (packed_solve.cu): when sys_list is given the kernel solves
    // the systems sys_list[0 .. *sys_count) instead of 0 .. batch
    const int* sys_list;
    const int* sys_count;
    // jitr_b200_solve_batched_indexed: n_free free matrices, system b uses number free_index[b] (free_stride is 0 then)
    const int* free_index;
    int n_free;
};


// Single channel, no coefficients, P = 8 ceil(N/8) <= 64, complex symmetric system: warp-per-system LDL^T on a packed lower
// triangle in shared memory, the nonlocal kernel / interaction matrix staged by bulk-async copies (packed_solve.cu).
// Systems that turn out not to be symmetric, or fail the threshold test, are appended to defer_list / defer_count.
bool packed_solve_supported(int nbasis);
int launch_packed_solve(const GenArgs& a, const int* symflag, int check_symmetry, int* defer_list, int* defer_count, cudaStream_t st);

}  // namespace jitr
