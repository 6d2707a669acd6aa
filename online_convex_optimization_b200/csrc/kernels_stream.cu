// kernels_stream.cu -- FTL / FTRL trajectories on device-resident sequences, small d.
//
// sim_stream_kernel<D, AV, TIN>: one thread per sequence b (layout z[B][T][D], y[B][T],
// float32 or float64 exactly as the reference's drivers hand them to simulate_alg,
// fast_algorithms.py:171-177); AV (algorithm, eta0) variants of the same sequence live
// in registers and share every loaded row.  theta, x and the loss accumulators never
// leave registers; summation is left to right, no FMA => bit-exact with the numba path
// for every d handled here (d <= 8), not only d = 1.
//
// Round-1 data path: each thread reads its own rows straight from global memory (the
// per-sequence rows z[b][t][:] are contiguous, so every 128-byte line a lane touches is
// fully consumed over ~3 steps and stays in L1 meanwhile).  The comparator needs the
// FINAL theta (fast_algorithms.py:113-114), so the sequence is streamed twice; the
// second pass hits L2 when B*T*(D+1)*sizeof(TIN) fits in it.
#include "oco_device.cuh"
#include "oco_internal.h"

namespace oco {

constexpr int kStreamBlock = 128;

struct StreamArgs {
    const void* z; const void* y;
    int64_t B, T;
    const double* tab;
    double* regret; double* cum; double* comp; double* theta_final; double* x_trace;
    Variants v;
};

template <int D, int AV, typename TIN>
__global__ void __launch_bounds__(kStreamBlock) sim_stream_kernel(const StreamArgs p) {
    const int64_t b = (int64_t)blockIdx.x * kStreamBlock + threadIdx.x;
    if (b >= p.B) return;
    const int A = p.v.A;
    const int v0 = blockIdx.y * AV;
    const int nv = min(AV, A - v0);
    const TIN* __restrict__ zb = (const TIN*)p.z + b * p.T * D;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * p.T;

    TrajState<D> s[AV];
#pragma unroll
    for (int v = 0; v < AV; ++v) traj_init<D>(s[v]);

    for (int64_t t = 0; t < p.T; ++t) {
        double z[D], hz[D];
#pragma unroll
        for (int j = 0; j < D; ++j) { z[j] = (double)zb[t * D + j]; hz[j] = 0.5 * z[j]; }
        const double y = (double)yb[t];
        const double* trow = p.tab + t * A + v0;
#pragma unroll
        for (int v = 0; v < AV; ++v) {
            if (v < nv) {
                double x[D];
                if (p.v.alg_flag[v0 + v] == 0) action_ftrl<D>(s[v].th, __ldg(trow + v), x);
                else action_ftl<D>(s[v].th, x);
                if (p.x_trace) {
                    double* xo = p.x_trace + ((b * A + v0 + v) * p.T + t) * D;
#pragma unroll
                    for (int j = 0; j < D; ++j) xo[j] = x[j];
                }
                play_and_update<D>(s[v], x, z, hz, y);
            }
        }
    }
    double act[AV][D], comp[AV];
#pragma unroll
    for (int v = 0; v < AV; ++v) { action_ftl<D>(s[v].th, act[v]); comp[v] = 0.0; }
    for (int64_t t = 0; t < p.T; ++t) {
        double z[D];
#pragma unroll
        for (int j = 0; j < D; ++j) z[j] = (double)zb[t * D + j];
        const double y = (double)yb[t];
#pragma unroll
        for (int v = 0; v < AV; ++v)
            if (v < nv) comp[v] = comp[v] + nhinge(dot_seq<D>(z, act[v]), y);
    }
#pragma unroll
    for (int v = 0; v < AV; ++v) {
        if (v < nv) {
            const int64_t o = b * A + v0 + v;
            if (p.regret) p.regret[o] = s[v].cum - comp[v];
            if (p.cum) p.cum[o] = s[v].cum;
            if (p.comp) p.comp[o] = comp[v];
            if (p.theta_final) {
#pragma unroll
                for (int j = 0; j < D; ++j) p.theta_final[o * D + j] = s[v].th[j];
            }
        }
    }
}

template <int D, typename TIN>
static cudaError_t launch_stream_d(const StreamArgs& a, cudaStream_t st) {
    const int A = a.v.A;
    const unsigned gx = (unsigned)((a.B + kStreamBlock - 1) / kStreamBlock);
    if (A == 1) sim_stream_kernel<D, 1, TIN><<<dim3(gx, 1), kStreamBlock, 0, st>>>(a);
    else sim_stream_kernel<D, 3, TIN><<<dim3(gx, (A + 2) / 3), kStreamBlock, 0, st>>>(a);
    return cudaGetLastError();
}

template <typename TIN>
static cudaError_t launch_stream_t(const StreamArgs& a, int d, cudaStream_t st) {
    switch (d) {
    case 1: return launch_stream_d<1, TIN>(a, st);
    case 2: return launch_stream_d<2, TIN>(a, st);
    case 3: return launch_stream_d<3, TIN>(a, st);
    case 4: return launch_stream_d<4, TIN>(a, st);
    case 5: return launch_stream_d<5, TIN>(a, st);
    case 6: return launch_stream_d<6, TIN>(a, st);
    case 7: return launch_stream_d<7, TIN>(a, st);
    case 8: return launch_stream_d<8, TIN>(a, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_sim_stream(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                              const Variants& v, const double* scale_tab,
                              double* regret, double* cum, double* comp, double* theta_final, double* x_trace,
                              cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    StreamArgs a;
    a.z = z; a.y = y; a.B = B; a.T = T; a.tab = scale_tab;
    a.regret = regret; a.cum = cum; a.comp = comp; a.theta_final = theta_final; a.x_trace = x_trace; a.v = v;
    return in_f32 ? launch_stream_t<float>(a, d, st) : launch_stream_t<double>(a, d, st);
}

}  // namespace oco
