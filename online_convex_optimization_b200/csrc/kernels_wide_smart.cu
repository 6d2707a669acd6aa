// kernels_wide_smart.cu -- single-switch SMART (fast_algorithms.py:118-164) for 8 < d <= 256.
//
//  smart_wide_kernel<K, TIN>: one warp per sequence, all A <= 4 thresholds of that sequence at once (the FTL
//  trajectory and the prefix loss of its leader do not depend on the threshold).  The reference has no limit on d;
//  smart_kernel<D> (kernels_warp.cu) keeps every state vector in each lane's registers and stops at d = 8.  Here
//  lane l owns coordinates l, l + 32, ... (K = ceil(d / 32) per lane) of theta_FTL, of the A FTRL states and of
//  v = sum y_i z_i; the dot products and squared norms of a round are lane-local partial sums followed by an xor
//  butterfly (every lane ends with the same bits, so all control flow stays warp-uniform).  The summation order
//  therefore differs from the reference's left-to-right `_dot` (fast_algorithms.py:11-16): 1e-12 relative, the
//  tolerance class of the other d > 8 kernels.
//
//  The O(t) prefix loss `_comparator_loss_prefix` (fast_algorithms.py:79-85, called at :158) is evaluated only in the
//  rounds the exact pruning bound of kernels_warp.cu (smart_prune_margin) cannot rule out.  When it is, the leader s
//  goes to shared memory and lane l sums rows i = l, l + 32, ... <= t, each row's dot product in the reference's own
//  sequential order, then a butterfly over the lanes.
#include "oco_device.cuh"
#include "oco_internal.h"
#include <algorithm>

namespace oco {

namespace {
constexpr int kWsMaxA = 4;          // == OCO_MAX_SMART_THRESHOLDS
constexpr int kWsWarps = 4;         // warps (sequences) per CTA
constexpr int kWsMaxD = 256;

__device__ __forceinline__ double ws_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ double ws_margin(int64_t t) {          // kernels_warp.cu::smart_prune_margin
    const double n = (double)(t + 1);
    return 1e-9 * n + 2.3e-16 * n * n;
}

struct WideSmartArgs {
    const void* z; const void* y;
    int64_t B, T; int d;
    const double* tab;       // [T][A] FTRL scales -(eta0 / sqrt(t + 1))
    const double* thresh;    // [A]
    int A;
    double* regret; int64_t* switch_t;    // [B][A]
};

template <int K, typename TIN>
__global__ void __launch_bounds__(kWsWarps * 32) smart_wide_kernel(const WideSmartArgs p) {
    __shared__ double s_sm[kWsWarps][kWsMaxD];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * kWsWarps + wib;
    if (b >= p.B) return;
    const int A = p.A, d = p.d;
    const int64_t T = p.T;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * T * d;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * T;
    double* __restrict__ ss = s_sm[wib];

    double thl[K], vy[K], thr[kWsMaxA][K], total[kWsMaxA], thresh[kWsMaxA];
    int64_t sw[kWsMaxA];
#pragma unroll
    for (int k = 0; k < K; ++k) { thl[k] = 0.0; vy[k] = 0.0; }
#pragma unroll
    for (int a = 0; a < kWsMaxA; ++a) {
        total[a] = 0.0; sw[a] = -1; thresh[a] = a < A ? p.thresh[a] : 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) thr[a][k] = 0.0;
    }
    double ftl_loss = 0.0;
    int n_unswitched = A;
    bool pm1 = true;

    // FTL action of theta (fast_algorithms.py:37-49), lane-distributed
    auto ftl_act = [&](const double (&th)[K], double (&x)[K]) {
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) n2 = n2 + th[k] * th[k];
        n2 = ws_sum(n2);
        if (n2 == 0.0) {
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = 0.0;
        } else {
            const double sc = -(1.0 / sqrt(n2));
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = sc * th[k];
        }
    };
    // sum_{i < n} 0.5 |z_i . s - y_i| with s in shared memory: rows strided over the lanes, sequential dot per row
    auto prefix_loss = [&](int64_t n) -> double {
        double part = 0.0;
        for (int64_t i = lane; i < n; i += 32) {
            const TIN* __restrict__ zi = zb + i * d;
            double q = (double)zi[0] * ss[0];
            for (int j = 1; j < d; ++j) q = q + (double)zi[j] * ss[j];
            part = part + nhinge(q, (double)yb[i]);
        }
        return ws_sum(part);
    };

    for (int64_t t = 0; t < T; ++t) {
        double z[K], hz[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = lane + 32 * k;
            z[k] = j < d ? (double)zb[t * d + j] : 0.0;
            hz[k] = 0.5 * z[k];
        }
        const double y = (double)yb[t];
        pm1 = pm1 && (y == 1.0 || y == -1.0);
#pragma unroll
        for (int k = 0; k < K; ++k) vy[k] = vy[k] + y * z[k];

        // FTL is always advanced (fast_algorithms.py:140-146)
        double x[K];
        ftl_act(thl, x);
        double pf = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) pf = pf + z[k] * x[k];
        pf = ws_sum(pf);
        const double dff = pf - y;
        if (dff > 0.0) {
#pragma unroll
            for (int k = 0; k < K; ++k) thl[k] = thl[k] + hz[k];
        } else if (dff < 0.0) {
#pragma unroll
            for (int k = 0; k < K; ++k) thl[k] = thl[k] - hz[k];
        }
        const double lf = 0.5 * (dff < 0.0 ? -dff : dff);
        ftl_loss = ftl_loss + lf;

        // played action: FTL before the switch, FTRL on a fresh theta with the global t + 1 after it (:148-154)
#pragma unroll
        for (int a = 0; a < kWsMaxA; ++a) {
            if (a < A) {
                if (sw[a] >= 0) {
                    const double sc = __ldg(p.tab + t * A + a);
                    double xr[K], n2 = 0.0;
#pragma unroll
                    for (int k = 0; k < K; ++k) { xr[k] = sc * thr[a][k]; n2 = n2 + xr[k] * xr[k]; }
                    n2 = ws_sum(n2);
                    if (!(n2 <= 1.0)) {
                        const double fac = 1.0 / sqrt(n2);
#pragma unroll
                        for (int k = 0; k < K; ++k) xr[k] = xr[k] * fac;
                    }
                    double pr = 0.0;
#pragma unroll
                    for (int k = 0; k < K; ++k) pr = pr + z[k] * xr[k];
                    pr = ws_sum(pr);
                    const double dr = pr - y;
                    total[a] = total[a] + 0.5 * (dr < 0.0 ? -dr : dr);
                    if (dr > 0.0) {
#pragma unroll
                        for (int k = 0; k < K; ++k) thr[a][k] = thr[a][k] + hz[k];
                    } else if (dr < 0.0) {
#pragma unroll
                        for (int k = 0; k < K; ++k) thr[a][k] = thr[a][k] - hz[k];
                    }
                } else {
                    total[a] = total[a] + lf;
                }
            }
        }

        // switch test (fast_algorithms.py:156-160): leader of the UPDATED theta_FTL, prefix loss over rows 0 .. t
        if (n_unswitched > 0) {
            double s[K];
            ftl_act(thl, s);
            bool need = true;
            if (pm1) {                                            // exact pruning (kernels_warp.cu)
                double thmin = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
                for (int a = 0; a < kWsMaxA; ++a) if (a < A && sw[a] < 0) thmin = fmin(thmin, thresh[a]);
                double sv = 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) sv = sv + s[k] * vy[k];
                sv = ws_sum(sv);
                const double lin = 0.5 * ((double)(t + 1) - sv);
                if (ftl_loss - lin < thmin - ws_margin(t)) need = false;
            }
            if (need) {                                           // warp-uniform: every lane holds the same sums
                __syncwarp();
#pragma unroll
                for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; if (j < d) ss[j] = s[k]; }
                __syncwarp();
                const double s_loss = prefix_loss(t + 1);
#pragma unroll
                for (int a = 0; a < kWsMaxA; ++a) {
                    if (a < A && sw[a] < 0 && ftl_loss - s_loss >= thresh[a]) { sw[a] = t; --n_unswitched; }
                }
            }
        }
    }

    // regret against the FTL leader of the whole sequence (fast_algorithms.py:162-164)
    double s[K];
    ftl_act(thl, s);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; if (j < d) ss[j] = s[k]; }
    __syncwarp();
    const double comp = prefix_loss(T);
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < kWsMaxA; ++a) {
            if (a < A) {
                p.regret[b * A + a] = total[a] - comp;
                if (p.switch_t) p.switch_t[b * A + a] = sw[a];
            }
        }
    }
}

template <int K, typename TIN>
cudaError_t launch_ws_k(const WideSmartArgs& a, cudaStream_t st) {
    const int64_t blocks = (a.B + kWsWarps - 1) / kWsWarps;
    if (blocks > 0x7fffffff) return cudaErrorInvalidValue;
    smart_wide_kernel<K, TIN><<<(unsigned)blocks, kWsWarps * 32, 0, st>>>(a);
    return cudaGetLastError();
}
template <typename TIN>
cudaError_t launch_ws_t(const WideSmartArgs& a, cudaStream_t st) {
    const int k = (a.d + 31) / 32;
    if (k <= 1) return launch_ws_k<1, TIN>(a, st);
    if (k <= 2) return launch_ws_k<2, TIN>(a, st);
    if (k <= 4) return launch_ws_k<4, TIN>(a, st);
    return launch_ws_k<8, TIN>(a, st);
}
}  // namespace

cudaError_t launch_smart_wide(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                              const double* thresh_dev, const double* scale_tab, int A,
                              double* regret, int64_t* switch_t, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (A < 1 || A > kWsMaxA || d < 1 || d > kWsMaxD) return cudaErrorInvalidValue;
    WideSmartArgs a;
    a.z = z; a.y = y; a.B = B; a.T = T; a.d = d; a.tab = scale_tab; a.thresh = thresh_dev; a.A = A;
    a.regret = regret; a.switch_t = switch_t;
    return in_f32 ? launch_ws_t<float>(a, st) : launch_ws_t<double>(a, st);
}

}  // namespace oco
