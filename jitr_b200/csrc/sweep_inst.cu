// One translation unit per solver MODE of the fused sweep (compile with -DJITR_SWEEP_MODE=<mode>) so the
// instantiations build in parallel.
#include "elastic_sweep.cuh"

#ifndef JITR_SWEEP_MODE
#error "compile with -DJITR_SWEEP_MODE=<mode>"
#endif

namespace jitr {
template int launch_sweep_mode<true, JITR_SWEEP_MODE>(const SweepArgs&, cudaStream_t);
template int launch_sweep_mode<false, JITR_SWEEP_MODE>(const SweepArgs&, cudaStream_t);
}  // namespace jitr
