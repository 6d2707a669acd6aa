// kernels_sweep_iid.cu -- sweep_kernel instantiations for the iid / Massart-noise families
// (sequence_generation.py:54-90); separate translation unit so the template grid compiles in parallel.
#include "sweep_impl.cuh"

namespace oco {
cudaError_t launch_sweep_iid(const SweepArgs& a, int d, cudaStream_t st) { return launch_sweep_s<IidStream>(a, d, st); }
}  // namespace oco
