// kernels_sweep_gpath.cu -- sweep_kernel instantiations for the g-path family
// (fast_algorithms.py:231-239); separate translation unit so the template grid compiles in parallel.
#include "sweep_impl.cuh"

namespace oco {
cudaError_t launch_sweep_gpath(const SweepArgs& a, int d, cudaStream_t st) { return launch_sweep_s<GPathStream>(a, d, st); }
}  // namespace oco
