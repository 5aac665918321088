// One translation unit per nbasis (compile with -DJITR_SWEEP_N=<N>) so the fully unrolled
// per-size kernels build in parallel.
#include "elastic_sweep.cuh"

#ifndef JITR_SWEEP_N
#error "compile with -DJITR_SWEEP_N=<nbasis>"
#endif

namespace jitr {
template int launch_sweep<JITR_SWEEP_N, true>(const SweepArgs&, cudaStream_t);
template int launch_sweep<JITR_SWEEP_N, false>(const SweepArgs&, cudaStream_t);
}  // namespace jitr
