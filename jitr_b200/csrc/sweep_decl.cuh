// Declarations shared by the dispatcher (elastic.cu) and the per-N instantiation units.
#pragma once
#include <cuda_runtime.h>

namespace jitr {

struct SweepArgs {
    int N, lmax, model, n_energies, n_params;
    int symmetric;                  // try the threshold-pivoted symmetric elimination first (fast paths)
    unsigned long long* fallbacks;  // device counter of systems re-solved with partial pivoting
    unsigned long long* fallback_mask;  // optional [n_pairs] bit mask of the systems that fell back (diagnostic)
    unsigned long long* work_counter;  // device counter handing out (sample, energy) pairs, zero at launch
    long long n_pairs;
    // packed modes (P = 48, 56, 64): a pair with a system that fails the threshold test is not finished here but appended
    // to defer_list / defer_count; the partial-pivoting modes then run over that list (pair_list / pair_count)
    int* defer_list;
    int* defer_count;
    const int* pair_list;
    const int* pair_count;
    // small batches (fewer pairs than warps): one work item per (pair, l) instead of per pair, no early break inside the
    // kernel; the partial waves kept (and the status) are decided afterwards from lsplit_flags[pair][l] (sweep_finalize)
    int lsplit;
    int* lsplit_flags;
    int* lsplit_seen;  // packed modes: [n_pairs], set when a pair has been appended to the deferred list (once per pair)
    const double* That;
    const double* mesh_x;
    const double* bvec;
    const double* etab;
    const double2* asym;
    const double* params;
    const double2* central;
    const double2* spin_orbit;
    const double2* coulomb;
    double tol;
    double2* S_out;
    int* nl_out;
    int* status;
};

// one instantiation per solver MODE (elastic_sweep.cuh); defined in sweep_inst.cu compiled with -DJITR_SWEEP_MODE
template <bool PARAMS, int MODE>
int launch_sweep_mode(const SweepArgs& a, cudaStream_t st);

// solver modes compiled into the library: fast paths (P = MODE) and generic paths (-2, -3 row slots)
#define JITR_SWEEP_MODES(X) X(16) X(24) X(32) X(40) X(48) X(56) X(64) X(-2) X(-3)

}  // namespace jitr
