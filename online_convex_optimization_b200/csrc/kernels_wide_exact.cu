// kernels_wide_exact.cu -- the exact_ftl.py paths for 8 < d <= 256: one WARP per sequence, lane l holds the
// coordinates j = l, l + 32, ... (coalesced row loads), per-step dot products / norms by xor-butterflies.
// The reference has no dimension limit (exact_ftl.py:67-117; its own smoke run uses d = 10, :460-475); the
// thread-per-sequence kernels of kernels_stream.cu keep theta in registers and stop at d = 8.
//
//   wide_exact_prefix_kernel   compute_prefix_actions + replay_exact_ftl (exact_ftl.py:280-333) in the linear
//                              regime (closed form v_k/||v_k||, 0 on ties), same contract as exact_ftl_linear_kernel
//   wide_replay_kernel         replay_exact_ftl for caller-supplied actions (:306-333)
//   wide_fixed_comp_kernel     _simulate_ftrl / simulate_alg with a FIXED comparator action and x_last
//                              (exact_ftl.py:230-277, run_ftrl(comparator_action=...)): one warp per (sequence, variant)
// Reductions change the summation order: 1e-12 relative (north_star, d > 1), asserted in tests.
#include "oco_device.cuh"
#include "oco_internal.h"

namespace oco {
namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxK = 8;                       // 32 * 8 = 256 coordinates

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ void warp_sum2(double& a, double& b) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a = a + __shfl_xor_sync(0xffffffffu, a, off);
        b = b + __shfl_xor_sync(0xffffffffu, b, off);
    }
}

struct WideExactArgs {
    const void* z; const void* y; int64_t B, T; int d;
    double* actions; const double* actions_in; double* cum; double* comp; double* regret; int* status;
    // fixed-comparator simulation
    Variants v; const double* scale_tab; const double* comparator; double* theta_final; double* x_last;
};

template <typename TIN, int K>
__device__ __forceinline__ void load_row(const TIN* __restrict__ row, int d, int lane, double (&z)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; z[k] = j < d ? (double)row[j] : 0.0; }
}

template <typename TIN, int K>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) wide_exact_prefix_kernel(const WideExactArgs p) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (b >= p.B) return;
    const int d = p.d;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * p.T * d;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * p.T;
    double* act = p.actions ? p.actions + b * (p.T + 1) * d : nullptr;
    double v[K], x[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { v[k] = 0.0; x[k] = 0.0; }
    if (act) {
#pragma unroll
        for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; if (j < d) act[j] = 0.0; }
    }
    double cum = 0.0; int bad = 0;
    for (int64_t t = 0; t < p.T; ++t) {
        double z[K];
        load_row<TIN, K>(zb + t * d, d, lane, z);
        const double y = (double)yb[t];
        double q = 0.0, n2 = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) { q = q + z[k] * x[k]; n2 = n2 + z[k] * z[k]; }
        warp_sum2(q, n2);
        cum = cum + nhinge(q, y);                            // x = actions[t], the FTL play at round t
        if (!(n2 <= 1.0 + 1e-6) || !(y == 1.0 || y == -1.0)) bad = 1;
        double vn2 = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) { v[k] = v[k] + y * z[k]; vn2 = vn2 + v[k] * v[k]; }
        vn2 = warp_sum(vn2);
        const double inv = vn2 == 0.0 ? 0.0 : 1.0 / sqrt(vn2);
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = vn2 == 0.0 ? 0.0 : inv * v[k];
        if (act) {
#pragma unroll
            for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; if (j < d) act[(t + 1) * d + j] = x[k]; }
        }
    }
    double comp = 0.0;                                       // comparator actions[T] (exact_ftl.py:328-330)
    for (int64_t t = 0; t < p.T; ++t) {
        double z[K];
        load_row<TIN, K>(zb + t * d, d, lane, z);
        double q = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) q = q + z[k] * x[k];
        q = warp_sum(q);
        comp = comp + fabs(q - (double)yb[t]);
    }
    comp = 0.5 * comp;
    if (lane == 0) {
        if (p.cum) p.cum[b] = cum;
        if (p.comp) p.comp[b] = comp;
        if (p.regret) p.regret[b] = cum - comp;
        if (p.status) p.status[b] = bad;
    }
}

template <typename TIN, int K>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) wide_replay_kernel(const WideExactArgs p) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (b >= p.B) return;
    const int d = p.d;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * p.T * d;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * p.T;
    const double* __restrict__ act = p.actions_in + b * (p.T + 1) * d;
    double xT[K];
    load_row<double, K>(act + p.T * d, d, lane, xT);
    double cum = 0.0, comp = 0.0;
    for (int64_t t = 0; t < p.T; ++t) {
        double z[K], x[K];
        load_row<TIN, K>(zb + t * d, d, lane, z);
        load_row<double, K>(act + t * d, d, lane, x);
        double q = 0.0, qc = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) { q = q + z[k] * x[k]; qc = qc + z[k] * xT[k]; }
        warp_sum2(q, qc);
        const double y = (double)yb[t];
        cum = cum + nhinge(q, y);
        comp = comp + fabs(qc - y);
    }
    comp = 0.5 * comp;
    if (lane == 0) {
        if (p.cum) p.cum[b] = cum;
        if (p.comp) p.comp[b] = comp;
        if (p.regret) p.regret[b] = cum - comp;
    }
}

// one warp per (sequence, variant); comparator [B][d] fixed (or nullptr: the FTL action of the final theta, second pass)
template <typename TIN, int K>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) wide_fixed_comp_kernel(const WideExactArgs p) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int A = p.v.A;
    if (w >= p.B * A) return;
    const int64_t b = w / A; const int a = (int)(w - b * A);
    const int d = p.d;
    const bool ftl = p.v.alg_flag[a] != 0;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * p.T * d;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * p.T;
    double th[K], x[K], c[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { th[k] = 0.0; x[k] = 0.0; c[k] = 0.0; }
    if (p.comparator) load_row<double, K>(p.comparator + b * d, d, lane, c);
    double cum = 0.0, comp = 0.0;
    for (int64_t t = 0; t < p.T; ++t) {
        double z[K];
        load_row<TIN, K>(zb + t * d, d, lane, z);
        const double y = (double)yb[t];
        if (ftl) {                                           // fast_algorithms.py:37-49
            double n2 = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) n2 = n2 + th[k] * th[k];
            n2 = warp_sum(n2);
            const double sc = n2 == 0.0 ? 0.0 : -(1.0 / sqrt(n2));
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = n2 == 0.0 ? 0.0 : sc * th[k];
        } else {                                             // fast_algorithms.py:52-66 / exact_ftl.py:200-210
            const double sc = p.scale_tab[t * A + a];
            double n2 = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) { x[k] = sc * th[k]; n2 = n2 + x[k] * x[k]; }
            n2 = warp_sum(n2);
            if (!(n2 <= 1.0)) {
                const double fac = 1.0 / sqrt(n2);
#pragma unroll
                for (int k = 0; k < K; ++k) x[k] = x[k] * fac;
            }
        }
        double q = 0.0, qc = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) { q = q + z[k] * x[k]; qc = qc + z[k] * c[k]; }
        warp_sum2(q, qc);
        const double df = q - y;
        cum = cum + 0.5 * fabs(df);
        comp = comp + fabs(qc - y);
        const double g = df > 0.0 ? 0.5 : (df < 0.0 ? -0.5 : 0.0);
#pragma unroll
        for (int k = 0; k < K; ++k) th[k] = th[k] + g * z[k];
    }
    if (!p.comparator) {                                     // reference comparator: FTL action of the final theta
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) n2 = n2 + th[k] * th[k];
        n2 = warp_sum(n2);
        const double sc = n2 == 0.0 ? 0.0 : -(1.0 / sqrt(n2));
#pragma unroll
        for (int k = 0; k < K; ++k) c[k] = n2 == 0.0 ? 0.0 : sc * th[k];
        comp = 0.0;
        for (int64_t t = 0; t < p.T; ++t) {
            double z[K];
            load_row<TIN, K>(zb + t * d, d, lane, z);
            double qc = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) qc = qc + z[k] * c[k];
            qc = warp_sum(qc);
            comp = comp + fabs(qc - (double)yb[t]);
        }
    }
    comp = 0.5 * comp;
    if (lane == 0) {
        if (p.cum) p.cum[w] = cum;
        if (p.comp) p.comp[w] = comp;
        if (p.regret) p.regret[w] = cum - comp;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = lane + 32 * k;
        if (j < d) {
            if (p.theta_final) p.theta_final[w * d + j] = th[k];
            if (p.x_last) p.x_last[w * d + j] = x[k];
        }
    }
}

template <typename TIN, int K>
cudaError_t launch_k(int which, const WideExactArgs& a, cudaStream_t st) {
    const int64_t warps = which == 2 ? a.B * a.v.A : a.B;
    const unsigned gx = (unsigned)((warps + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (which == 0) wide_exact_prefix_kernel<TIN, K><<<gx, 32 * kWarpsPerBlock, 0, st>>>(a);
    else if (which == 1) wide_replay_kernel<TIN, K><<<gx, 32 * kWarpsPerBlock, 0, st>>>(a);
    else wide_fixed_comp_kernel<TIN, K><<<gx, 32 * kWarpsPerBlock, 0, st>>>(a);
    return cudaGetLastError();
}
template <typename TIN>
cudaError_t launch_t(int which, const WideExactArgs& a, cudaStream_t st) {
    const int K = (a.d + 31) / 32;
    if (K <= 1) return launch_k<TIN, 1>(which, a, st);
    if (K <= 2) return launch_k<TIN, 2>(which, a, st);
    if (K <= 4) return launch_k<TIN, 4>(which, a, st);
    if (K <= kMaxK) return launch_k<TIN, kMaxK>(which, a, st);
    return cudaErrorInvalidValue;
}

}  // namespace

__global__ void status_scale_kernel(int* status, int64_t B, int factor) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) status[i] = status[i] ? factor : 0;
}
cudaError_t launch_status_scale(int* status, int64_t B, int factor, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    status_scale_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(status, B, factor);
    return cudaGetLastError();
}

cudaError_t launch_wide_exact_prefix(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                                     double* actions, double* cum, double* comp, double* regret, int* status, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    WideExactArgs a{};
    a.z = z; a.y = y; a.B = B; a.T = T; a.d = d; a.actions = actions; a.cum = cum; a.comp = comp; a.regret = regret; a.status = status;
    return in_f32 ? launch_t<float>(0, a, st) : launch_t<double>(0, a, st);
}

cudaError_t launch_wide_replay(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                               const double* actions, double* cum, double* comp, double* regret, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    WideExactArgs a{};
    a.z = z; a.y = y; a.B = B; a.T = T; a.d = d; a.actions_in = actions; a.cum = cum; a.comp = comp; a.regret = regret;
    return in_f32 ? launch_t<float>(1, a, st) : launch_t<double>(1, a, st);
}

cudaError_t launch_wide_fixed_comp(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                                   const Variants& v, const double* scale_tab, const double* comparator,
                                   double* regret, double* cum, double* comp, double* theta_final, double* x_last, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    WideExactArgs a{};
    a.z = z; a.y = y; a.B = B; a.T = T; a.d = d; a.v = v; a.scale_tab = scale_tab; a.comparator = comparator;
    a.regret = regret; a.cum = cum; a.comp = comp; a.theta_final = theta_final; a.x_last = x_last;
    return in_f32 ? launch_t<float>(2, a, st) : launch_t<double>(2, a, st);
}

}  // namespace oco
