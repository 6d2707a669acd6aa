// kernels_gsweep_wide.cu -- gsweep_kernel instantiations for d = 6, 7, 8 (single-pass comparator only)
#include "gsweep_impl.cuh"

namespace oco {
cudaError_t launch_gsweep_wide(const SweepArgs& a, int d, cudaStream_t st) {
    switch (d) {
    case 6: return launch_gsweep_d<6>(a, st);
    case 7: return launch_gsweep_d<7>(a, st);
    case 8: return launch_gsweep_d<8>(a, st);
    default: return cudaErrorInvalidValue;
    }
}
}  // namespace oco
