// kernels_gsweep_lo.cu -- gsweep_kernel instantiations for d = 1, 2, 3 (separate TU: parallel compilation)
#include "gsweep_impl.cuh"

namespace oco {
cudaError_t launch_gsweep_lo(const SweepArgs& a, int d, cudaStream_t st) {
    switch (d) {
    case 1: return launch_gsweep_d<1>(a, st);
    case 2: return launch_gsweep_d<2>(a, st);
    case 3: return launch_gsweep_d<3>(a, st);
    default: return cudaErrorInvalidValue;
    }
}
}  // namespace oco
