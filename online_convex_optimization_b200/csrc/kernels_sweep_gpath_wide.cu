// kernels_sweep_gpath_wide.cu -- sweep_kernel (two-pass comparator, bit-exact) instantiations for the g-path family at
// d = 6, 7, 8 (separate TU: parallel compilation).  theta of every variant lives in registers: d = 8 with 8 FTRL variants
// per thread spills heavily -- these shapes exist for coverage (the reference has no limit on d), not for speed.
#include "sweep_impl.cuh"

namespace oco {
cudaError_t launch_sweep_gpath_wide(const SweepArgs& a, int d, cudaStream_t st) {
    switch (d) {
    case 6: return launch_sweep_d<GPathStream, 6>(a, st);
    case 7: return launch_sweep_d<GPathStream, 7>(a, st);
    case 8: return launch_sweep_d<GPathStream, 8>(a, st);
    default: return cudaErrorInvalidValue;
    }
}
}  // namespace oco
