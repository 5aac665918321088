// Declarations shared by the dispatcher (elastic.cu) and the per-N instantiation units.
#pragma once
#include <cuda_runtime.h>

namespace jitr {

struct SweepArgs {
    int N, lmax, model, n_energies, n_params;
    long long n_pairs;
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

template <bool PARAMS>
int launch_sweep(const SweepArgs& a, cudaStream_t st);

}  // namespace jitr
