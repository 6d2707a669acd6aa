// kernels_gsweep_hi.cu -- gsweep_kernel instantiations for d = 4, 5 (d = 5 is BASELINE config 5's shape)
#include "gsweep_impl.cuh"

namespace oco {
cudaError_t launch_gsweep_hi(const SweepArgs& a, int d, cudaStream_t st) {
    switch (d) {
    case 4: return launch_gsweep_d<4>(a, st);
    case 5: return launch_gsweep_d<5>(a, st);
    default: return cudaErrorInvalidValue;
    }
}
}  // namespace oco
