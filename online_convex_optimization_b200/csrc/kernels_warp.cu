// kernels_warp.cu -- warp-per-sequence kernels.
//
//  sim_wide_kernel<K, TIN>   FTL / FTRL for larger d (d <= 32*K, BASELINE config 4: d = 64):
//      lane l holds elements l, l+32, ... of theta / x / z_t; dot products and squared
//      norms are lane-local partial sums (in k order) followed by an xor-butterfly, so
//      the reduction order differs from the reference's left-to-right sum: parity is
//      1e-12 relative regret (north_star), not bit-exact.  Rows are read with one
//      coalesced load per K (512 B per row at d = 64 fp64) and prefetched one step ahead.
//      This is the HBM-bound path: (d+1)*sizeof(TIN) bytes per trajectory-step.
//      comparator_mode 0: second pass over the sequence (exact restatement of
//      fast_algorithms.py:113-114).  comparator_mode 1: single pass -- while playing,
//      accumulate v = sum_t y_t z_t; when every row satisfies ||z_t||_2 <= 1 and
//      y_t in {-1,+1} (true for every generator in sequence_generation.py) then
//      |z_t.a| <= 1 for the unit-norm comparator a and 0.5*|z_t.a - y_t| = 0.5*(1 - y_t z_t.a),
//      hence comp = 0.5*(T - a.v) with no second read.  Rows violating the
//      precondition are detected on the fly and the warp falls back to the second pass.
//
//  smart_kernel<D, TIN>      single-switch SMART (fast_algorithms.py:118-164), one warp per
//      sequence, all thresholds of that sequence at once (the FTL trajectory and the
//      prefix loss of its leader do not depend on the threshold).  The O(t) prefix loss
//      (fast_algorithms.py:158 -> :79-85) is spread over the lanes (rows i = lane, lane+32,
//      ...; each row's dot product is the reference's sequential one) and combined by a
//      butterfly; the FTL / FTRL states are replicated in every lane, bit-identically.
#include "oco_device.cuh"
#include "oco_internal.h"
#include <algorithm>
#include <cstdlib>

namespace oco {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// ------------------------------------------------------------------ wide d
struct WideArgs {
    const void* z; const void* y;
    int64_t B, T; int d;
    const double* tab;
    double* regret; double* cum; double* comp; double* theta_final;
    int comparator_mode;
    Variants v;
};

constexpr int kWideWarps = 8;   // warps per CTA

template <int K, typename TIN>
__global__ void __launch_bounds__(kWideWarps * 32) sim_wide_kernel(const WideArgs p) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * kWideWarps + (threadIdx.x >> 5);
    const int A = p.v.A;
    if (w >= p.B * A) return;
    const int64_t b = w / A; const int a = (int)(w - b * A);
    const int d = p.d;
    const bool ftrl = p.v.alg_flag[a] == 0;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * p.T * d;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * p.T;

    double th[K], vy[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { th[k] = 0.0; vy[k] = 0.0; }
    double cum = 0.0;
    bool linear_ok = true;
    double excess = 0.0;

    double zn[K], yn = 0.0;          // prefetched row
#pragma unroll
    for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; zn[k] = (p.T > 0 && j < d) ? (double)zb[j] : 0.0; }
    if (p.T > 0) yn = (double)yb[0];

    for (int64_t t = 0; t < p.T; ++t) {
        double z[K]; const double y = yn;
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = zn[k];
        if (t + 1 < p.T) {
#pragma unroll
            for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; zn[k] = (j < d) ? (double)zb[(t + 1) * d + j] : 0.0; }
            yn = (double)yb[t + 1];
        }
        double x[K];
        if (ftrl) {
            const double sc = __ldg(p.tab + t * A + a);
            double n2 = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) { x[k] = sc * th[k]; n2 = n2 + x[k] * x[k]; }
            n2 = warp_sum(n2);
            if (!(n2 <= 1.0)) {
                const double fac = 1.0 / sqrt(n2);
#pragma unroll
                for (int k = 0; k < K; ++k) x[k] = x[k] * fac;
            }
        } else {
            double n2 = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) n2 = n2 + th[k] * th[k];
            n2 = warp_sum(n2);
            if (n2 == 0.0) {
#pragma unroll
                for (int k = 0; k < K; ++k) x[k] = 0.0;
            } else {
                const double sc = -(1.0 / sqrt(n2));
#pragma unroll
                for (int k = 0; k < K; ++k) x[k] = sc * th[k];
            }
        }
        double q = 0.0, zz = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) { q = q + z[k] * x[k]; zz = zz + z[k] * z[k]; }
        q = warp_sum(q);
        const double df = q - y;
        cum = cum + 0.5 * (df < 0.0 ? -df : df);
        const double g = df > 0.0 ? 0.5 : (df < 0.0 ? -0.5 : 0.0);
#pragma unroll
        for (int k = 0; k < K; ++k) th[k] = th[k] + g * z[k];
        if (p.comparator_mode == 1) {
            zz = warp_sum(zz);
            excess = excess + fmax(zz - 1.0, 0.0);            // budget: see kernels_stream.cu (OCO_COMPARATOR_LINEAR)
            if (!(y == 1.0 || y == -1.0)) linear_ok = false;
#pragma unroll
            for (int k = 0; k < K; ++k) vy[k] = vy[k] + y * z[k];
        }
    }
    // comparator = FTL action of the final theta
    double act[K];
    {
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) n2 = n2 + th[k] * th[k];
        n2 = warp_sum(n2);
        const double sc = (n2 == 0.0) ? 0.0 : -(1.0 / sqrt(n2));
#pragma unroll
        for (int k = 0; k < K; ++k) act[k] = (n2 == 0.0) ? 0.0 : sc * th[k];
    }
    double comp = 0.0;
    if (!(excess <= 1e-13 * (double)p.T)) linear_ok = false;
    if (p.comparator_mode == 1 && linear_ok) {
        double av = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) av = av + act[k] * vy[k];
        av = warp_sum(av);
        comp = 0.5 * ((double)p.T - av);
    } else {
        for (int64_t t = 0; t < p.T; ++t) {
            double q = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; if (j < d) q = q + (double)zb[t * d + j] * act[k]; }
            q = warp_sum(q);
            comp = comp + nhinge(q, (double)yb[t]);
        }
    }
    if (lane == 0) {
        if (p.regret) p.regret[w] = cum - comp;
        if (p.cum) p.cum[w] = cum;
        if (p.comp) p.comp[w] = comp;
    }
    if (p.theta_final) {
#pragma unroll
        for (int k = 0; k < K; ++k) { const int j = lane + 32 * k; if (j < d) p.theta_final[w * d + j] = th[k]; }
    }
}

template <typename TIN>
static cudaError_t launch_wide_t(const WideArgs& a, cudaStream_t st) {
    const int64_t nw = a.B * a.v.A;
    const unsigned gx = (unsigned)((nw + kWideWarps - 1) / kWideWarps);
    const int K = (a.d + 31) / 32;
    switch (K) {
    case 1: sim_wide_kernel<1, TIN><<<gx, kWideWarps * 32, 0, st>>>(a); break;
    case 2: sim_wide_kernel<2, TIN><<<gx, kWideWarps * 32, 0, st>>>(a); break;
    case 3: sim_wide_kernel<3, TIN><<<gx, kWideWarps * 32, 0, st>>>(a); break;
    case 4: sim_wide_kernel<4, TIN><<<gx, kWideWarps * 32, 0, st>>>(a); break;
    case 5: case 6: case 7: case 8: sim_wide_kernel<8, TIN><<<gx, kWideWarps * 32, 0, st>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_sim_wide(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                            const Variants& v, const double* scale_tab,
                            double* regret, double* cum, double* comp, double* theta_final, int comparator_mode,
                            cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    WideArgs a;
    a.z = z; a.y = y; a.B = B; a.T = T; a.d = d; a.tab = scale_tab;
    a.regret = regret; a.cum = cum; a.comp = comp; a.theta_final = theta_final;
    a.comparator_mode = comparator_mode; a.v = v;
    return in_f32 ? launch_wide_t<float>(a, st) : launch_wide_t<double>(a, st);
}

// ------------------------------------------------------------------ SMART
constexpr int kSmartMaxA = 4;
constexpr int kSmartWarps = 4;

struct SmartArgs {
    const void* z; const void* y;
    int64_t B, T;
    const double* tab;       // [T][A] FTRL scales (eta0 per variant)
    const double* thresh;    // [A]
    int A;
    double* regret; int64_t* switch_t;    // [B][A]
    const int* only_flagged;              // optional [B]: the exact kernel processes sequence b only if non-zero
    int* flags;                           // linear kernel: [B] set to 1 when a row violates the precondition
};

// EXACT pruning of the O(t) prefix loss (both exact kernels).  For labels y_i = +-1, |q - y| >= 1 - y q for ANY q, so
//   s_loss = sum_{i<=t} 0.5|z_i.s - y_i|  >=  0.5((t + 1) - s.vy),   vy = sum_{i<=t} y_i z_i        (O(d) per round)
// and the switch test `ftl_loss - s_loss >= thresh` (fast_algorithms.py:157-160) cannot fire while
// ftl_loss - 0.5((t + 1) - s.vy) < min(unswitched thresholds) - delta.  delta = 1e-9 (t+1) + 2.3e-16 (t+1)^2 covers the
// worst-case rounding of every sum involved (the reference's own sequential s_loss included), so a skipped round is a
// round the reference does not switch in; otherwise the literal prefix loss is evaluated and decides, as before.
// Sequences that never come near their threshold (every iid stream of the reference's drivers) cost O(T d) instead of
// O(T^2 d) with unchanged results; no assumption on the row norms (unlike smart_linear_kernel).
__device__ __forceinline__ double smart_prune_margin(int64_t t) {
    const double n = (double)(t + 1);
    return 1e-9 * n + 2.3e-16 * n * n;
}

// CACHED: the warp first copies its sequence, widened to double, into shared memory as D + 1 planes
// [j][i] (conflict-free for lane-strided rows); the O(t) prefix loss of every round then reads LDS.64
// instead of re-fetching and re-converting the rows from L1 / L2 (35 % of the stall samples and, for
// float32 inputs, a conversion-pipe bottleneck in the uncached version).  Used for short sequences only
// (launch_smart_d: <= 16 KB per warp), where it does not cost occupancy.
template <int D, typename TIN, bool CACHED>
__global__ void __launch_bounds__(kSmartWarps * 32) smart_kernel(const SmartArgs p) {
    extern __shared__ __align__(16) double smart_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (b >= p.B) return;
    if (p.only_flagged && p.only_flagged[b] == 0) return;
    const int A = p.A;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * p.T * D;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * p.T;
    const int64_t Tp = (p.T + 31) & ~(int64_t)31;                // plane stride
    double* __restrict__ pl = smart_smem + (int64_t)wib * (D + 1) * Tp;
    if constexpr (CACHED) {
        for (int64_t i = lane; i < p.T; i += 32) {
#pragma unroll
            for (int j = 0; j < D; ++j) pl[j * Tp + i] = (double)zb[i * D + j];
            pl[D * Tp + i] = (double)yb[i];
        }
        __syncwarp();
    }
    auto zval = [&](int64_t i, int j) -> double {
        if constexpr (CACHED) return pl[j * Tp + i]; else return (double)zb[i * D + j];
    };
    auto yval = [&](int64_t i) -> double {
        if constexpr (CACHED) return pl[D * Tp + i]; else return (double)yb[i];
    };

    double thl[D];                       // FTL state (always advanced, :140-146)
    double thr[kSmartMaxA][D];           // FTRL state per threshold, starts at the switch (:148-154)
    double total[kSmartMaxA], thresh[kSmartMaxA];
    int64_t sw[kSmartMaxA];
#pragma unroll
    for (int j = 0; j < D; ++j) thl[j] = 0.0;
#pragma unroll
    for (int a = 0; a < kSmartMaxA; ++a) {
        total[a] = 0.0; sw[a] = -1; thresh[a] = a < A ? p.thresh[a] : 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) thr[a][j] = 0.0;
    }
    double ftl_loss = 0.0;
    int n_unswitched = A;
    double vy[D];                        // sum y_i z_i (pruning bound)
    bool pm1 = true;                     // every label so far is +-1
#pragma unroll
    for (int j = 0; j < D; ++j) vy[j] = 0.0;

    for (int64_t t = 0; t < p.T; ++t) {
        double z[D], hz[D];
#pragma unroll
        for (int j = 0; j < D; ++j) { z[j] = zval(t, j); hz[j] = 0.5 * z[j]; }
        const double y = yval(t);
        pm1 = pm1 && (y == 1.0 || y == -1.0);
#pragma unroll
        for (int j = 0; j < D; ++j) vy[j] = vy[j] + y * z[j];
        double x[D];
        action_ftl<D>(thl, x);
        const double pf = dot_seq<D>(z, x);
        const double dff = pf - y;
        if (dff > 0.0) {
#pragma unroll
            for (int j = 0; j < D; ++j) thl[j] = thl[j] + hz[j];
        } else if (dff < 0.0) {
#pragma unroll
            for (int j = 0; j < D; ++j) thl[j] = thl[j] - hz[j];
        }
        const double lf = 0.5 * (dff < 0.0 ? -dff : dff);
        ftl_loss = ftl_loss + lf;
#pragma unroll
        for (int a = 0; a < kSmartMaxA; ++a) {
            if (a < A) {
                if (sw[a] >= 0) {
                    double xr[D];
                    action_ftrl<D>(thr[a], __ldg(p.tab + t * A + a), xr);
                    const double pr = dot_seq<D>(z, xr);
                    const double dr = pr - y;
                    total[a] = total[a] + 0.5 * (dr < 0.0 ? -dr : dr);
                    if (dr > 0.0) {
#pragma unroll
                        for (int j = 0; j < D; ++j) thr[a][j] = thr[a][j] + hz[j];
                    } else if (dr < 0.0) {
#pragma unroll
                        for (int j = 0; j < D; ++j) thr[a][j] = thr[a][j] - hz[j];
                    }
                } else {
                    total[a] = total[a] + lf;
                }
            }
        }
        bool need = n_unswitched > 0;
        double s[D];
        if (need) {
            action_ftl<D>(thl, s);
            if (pm1) {                                           // exact pruning (see smart_prune_margin)
                double thmin = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
                for (int a = 0; a < kSmartMaxA; ++a) if (a < A && sw[a] < 0) thmin = fmin(thmin, thresh[a]);
                const double lin = 0.5 * ((double)(t + 1) - dot_seq<D>(s, vy));
                if (ftl_loss - lin < thmin - smart_prune_margin(t)) need = false;
            }
        }
        if (need) {
            double part = 0.0;
#pragma unroll 4
            for (int64_t i = lane; i <= t; i += 32) {
                double zi[D];
#pragma unroll
                for (int j = 0; j < D; ++j) zi[j] = zval(i, j);
                part = part + nhinge(dot_seq<D>(zi, s), yval(i));
            }
            const double s_loss = warp_sum(part);
#pragma unroll
            for (int a = 0; a < kSmartMaxA; ++a) {
                if (a < A && sw[a] < 0 && ftl_loss - s_loss >= thresh[a]) { sw[a] = t; --n_unswitched; }
            }
        }
    }
    double s[D];
    action_ftl<D>(thl, s);
    double part = 0.0;
    for (int64_t i = lane; i < p.T; i += 32) {
        double zi[D];
#pragma unroll
        for (int j = 0; j < D; ++j) zi[j] = zval(i, j);
        part = part + nhinge(dot_seq<D>(zi, s), yval(i));
    }
    const double comp = warp_sum(part);
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < kSmartMaxA; ++a) {
            if (a < A) {
                p.regret[b * A + a] = total[a] - comp;
                if (p.switch_t) p.switch_t[b * A + a] = sw[a];
            }
        }
    }
}

// CTA-per-sequence variant for FEW long sequences (T >= kSmartCtaMinT and B <= 592, see launch_smart_d).  The O(t) prefix loss of every round is
// spread over the 4 warps of a CTA; warp 0 alone advances the (inherently sequential) FTL / FTRL states and
// publishes the leader s through shared memory, so the serial part is issued once per sequence, not once per
// warp.  Two barriers per round.  CACHED keeps the widened sequence in shared memory (D + 1 planes) when it
// fits, which also removes the per-round float32 -> float64 conversions.  The prefix loss is summed lane-strided
// (stride 128), butterfly per warp, then the four warp sums in order: a different association than the warp
// kernel, same 1e-12 tolerance class.
constexpr int kSmartCtaMinT = 384;
constexpr int kSmartCtaThreads = 128;

template <int D, typename TIN, bool CACHED>
__global__ void __launch_bounds__(kSmartCtaThreads) smart_cta_kernel(const SmartArgs p) {
    extern __shared__ __align__(16) double smart_smem[];
    __shared__ double sh_s[D];
    __shared__ double sh_part[kSmartCtaThreads / 32];
    __shared__ int sh_need;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int64_t b = blockIdx.x;
    if (p.only_flagged && p.only_flagged[b] == 0) return;             // uniform per CTA
    const int A = p.A;
    const int64_t T = p.T;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * T * D;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * T;
    const int64_t Tp = (T + 31) & ~(int64_t)31;
    double* __restrict__ pl = smart_smem;
    if constexpr (CACHED) {
        for (int64_t i = tid; i < T; i += kSmartCtaThreads) {
#pragma unroll
            for (int j = 0; j < D; ++j) pl[j * Tp + i] = (double)zb[i * D + j];
            pl[D * Tp + i] = (double)yb[i];
        }
        __syncthreads();
    }
    auto zval = [&](int64_t i, int j) -> double {
        if constexpr (CACHED) return pl[j * Tp + i]; else return (double)zb[i * D + j];
    };
    auto yval = [&](int64_t i) -> double {
        if constexpr (CACHED) return pl[D * Tp + i]; else return (double)yb[i];
    };
    auto prefix_part = [&](const double (&s)[D], int64_t n) -> double {   // this thread's share of sum_{i<n} 0.5|z_i.s - y_i|
        double part = 0.0;
#pragma unroll 4
        for (int64_t i = tid; i < n; i += kSmartCtaThreads) {
            double zi[D];
#pragma unroll
            for (int j = 0; j < D; ++j) zi[j] = zval(i, j);
            part = part + nhinge(dot_seq<D>(zi, s), yval(i));
        }
        return warp_sum(part);
    };

    // warp 0 state (the other warps carry dead copies)
    double thl[D], thr[kSmartMaxA][D], total[kSmartMaxA], thresh[kSmartMaxA];
    int64_t sw[kSmartMaxA];
#pragma unroll
    for (int j = 0; j < D; ++j) thl[j] = 0.0;
#pragma unroll
    for (int a = 0; a < kSmartMaxA; ++a) {
        total[a] = 0.0; sw[a] = -1; thresh[a] = a < A ? p.thresh[a] : 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) thr[a][j] = 0.0;
    }
    double ftl_loss = 0.0;
    int n_unswitched = A;
    double vy[D];                        // sum y_i z_i (pruning bound, warp 0)
    bool pm1 = true;
#pragma unroll
    for (int j = 0; j < D; ++j) vy[j] = 0.0;

    for (int64_t t = 0; t < T; ++t) {
        if (wib == 0) {                                               // fast_algorithms.py:133-156, as in smart_kernel
            double z[D], hz[D];
#pragma unroll
            for (int j = 0; j < D; ++j) { z[j] = zval(t, j); hz[j] = 0.5 * z[j]; }
            const double y = yval(t);
            pm1 = pm1 && (y == 1.0 || y == -1.0);
#pragma unroll
            for (int j = 0; j < D; ++j) vy[j] = vy[j] + y * z[j];
            double x[D];
            action_ftl<D>(thl, x);
            const double dff = dot_seq<D>(z, x) - y;
            if (dff > 0.0) {
#pragma unroll
                for (int j = 0; j < D; ++j) thl[j] = thl[j] + hz[j];
            } else if (dff < 0.0) {
#pragma unroll
                for (int j = 0; j < D; ++j) thl[j] = thl[j] - hz[j];
            }
            const double lf = 0.5 * (dff < 0.0 ? -dff : dff);
            ftl_loss = ftl_loss + lf;
#pragma unroll
            for (int a = 0; a < kSmartMaxA; ++a) {
                if (a < A) {
                    if (sw[a] >= 0) {
                        double xr[D];
                        action_ftrl<D>(thr[a], __ldg(p.tab + t * A + a), xr);
                        const double dr = dot_seq<D>(z, xr) - y;
                        total[a] = total[a] + 0.5 * (dr < 0.0 ? -dr : dr);
                        if (dr > 0.0) {
#pragma unroll
                            for (int j = 0; j < D; ++j) thr[a][j] = thr[a][j] + hz[j];
                        } else if (dr < 0.0) {
#pragma unroll
                            for (int j = 0; j < D; ++j) thr[a][j] = thr[a][j] - hz[j];
                        }
                    } else {
                        total[a] = total[a] + lf;
                    }
                }
            }
            bool need_w = n_unswitched > 0;
            if (need_w) {
                double s[D];
                action_ftl<D>(thl, s);
                if (pm1) {                                        // exact pruning (see smart_prune_margin)
                    double thmin = __longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
                    for (int a = 0; a < kSmartMaxA; ++a) if (a < A && sw[a] < 0) thmin = fmin(thmin, thresh[a]);
                    const double lin = 0.5 * ((double)(t + 1) - dot_seq<D>(s, vy));
                    if (ftl_loss - lin < thmin - smart_prune_margin(t)) need_w = false;
                }
                if (need_w && lane < D) {
                    double out = 0.0;
#pragma unroll
                    for (int j = 0; j < D; ++j) if (j == lane) out = s[j];
                    sh_s[lane] = out;
                }
            }
            if (lane == 0) sh_need = need_w ? 1 : 0;
        }
        __syncthreads();
        const bool need = sh_need != 0;                               // uniform across the CTA
        if (need) {
            double s[D];
#pragma unroll
            for (int j = 0; j < D; ++j) s[j] = sh_s[j];
            const double ws = prefix_part(s, t + 1);
            if (lane == 0) sh_part[wib] = ws;
        }
        __syncthreads();
        if (need && wib == 0) {
            double s_loss = sh_part[0];
#pragma unroll
            for (int w = 1; w < kSmartCtaThreads / 32; ++w) s_loss = s_loss + sh_part[w];
#pragma unroll
            for (int a = 0; a < kSmartMaxA; ++a) {
                if (a < A && sw[a] < 0 && ftl_loss - s_loss >= thresh[a]) { sw[a] = t; --n_unswitched; }
            }
        }
    }
    // regret against the final leader (fast_algorithms.py:162-164)
    if (wib == 0 && lane < D) {
        double s[D];
        action_ftl<D>(thl, s);
        double out = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) if (j == lane) out = s[j];
        sh_s[lane] = out;
    }
    __syncthreads();
    {
        double s[D];
#pragma unroll
        for (int j = 0; j < D; ++j) s[j] = sh_s[j];
        const double ws = prefix_part(s, T);
        if (lane == 0) sh_part[wib] = ws;
    }
    __syncthreads();
    if (tid == 0) {
        double comp = sh_part[0];
#pragma unroll
        for (int w = 1; w < kSmartCtaThreads / 32; ++w) comp = comp + sh_part[w];
#pragma unroll
        for (int a = 0; a < kSmartMaxA; ++a) {
            if (a < A) {
                p.regret[b * A + a] = total[a] - comp;
                if (p.switch_t) p.switch_t[b * A + a] = sw[a];
            }
        }
    }
}

// SMART in the linear regime, O(T*d) instead of O(T^2*d): when the rows are in the unit ball up to the budget
// sum_i (||z_i||^2 - 1)+ <= 1e-13 * T (float64-clipped rows; float32-normalised rows are re-run exactly) and
// y_i = +-1, the prefix loss of the leader s (fast_algorithms.py:157-158) is 0.5*sum_{i<=t}(1 - y_i z_i.s)
// = 0.5*((t+1) - s.v_t) with v_t = sum_{i<=t} y_i z_i, so the switch test needs one dot product per round.
// One thread per sequence, all thresholds at once.  Sequences with a violating row are flagged and
// re-run by the exact warp kernel.  Tolerance class 1e-12 (the switch test compares against a sum
// formed in a different order).
template <int D, typename TIN>
__global__ void __launch_bounds__(128) smart_linear_kernel(const SmartArgs p) {
    const int64_t b = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (b >= p.B) return;
    const int A = p.A;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * p.T * D;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * p.T;
    double thl[D], vy[D], thr[kSmartMaxA][D], total[kSmartMaxA], thresh[kSmartMaxA];
    int64_t sw[kSmartMaxA];
#pragma unroll
    for (int j = 0; j < D; ++j) { thl[j] = 0.0; vy[j] = 0.0; }
#pragma unroll
    for (int a = 0; a < kSmartMaxA; ++a) {
        total[a] = 0.0; sw[a] = -1; thresh[a] = a < A ? p.thresh[a] : 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) thr[a][j] = 0.0;
    }
    double ftl_loss = 0.0, excess = 0.0;
    int bad = 0, n_unswitched = A;
    for (int64_t t = 0; t < p.T; ++t) {
        double z[D], hz[D];
#pragma unroll
        for (int j = 0; j < D; ++j) { z[j] = (double)zb[t * D + j]; hz[j] = 0.5 * z[j]; }
        const double y = (double)yb[t];
        double zz = z[0] * z[0];
#pragma unroll
        for (int j = 1; j < D; ++j) zz = zz + z[j] * z[j];
        excess = excess + fmax(zz - 1.0, 0.0);
        if (!(y == 1.0 || y == -1.0)) bad = 1;
#pragma unroll
        for (int j = 0; j < D; ++j) vy[j] = vy[j] + y * z[j];
        double x[D];
        ftl_action<D>(thl, x);                                   // fast_algorithms.py:140-146
        const double dff = dot_seq<D>(z, x) - y;
        grad_step<D>(thl, hz, dff);
        const double lf = 0.5 * fabs(dff);
        ftl_loss = ftl_loss + lf;
#pragma unroll
        for (int a = 0; a < kSmartMaxA; ++a) {
            if (a < A) {
                if (sw[a] >= 0) {                                // :148-154, FTRL on its own theta, global t+1
                    double xr[D];
                    action_ftrl<D>(thr[a], __ldg(p.tab + t * A + a), xr);
                    const double dr = dot_seq<D>(z, xr) - y;
                    total[a] = total[a] + 0.5 * fabs(dr);
                    grad_step<D>(thr[a], hz, dr);
                } else {
                    total[a] = total[a] + lf;
                }
            }
        }
        if (n_unswitched > 0) {                                  // :156-160 with the closed-form prefix loss
            double s[D];
            ftl_action<D>(thl, s);
            const double s_loss = 0.5 * ((double)(t + 1) - dot_seq<D>(s, vy));
#pragma unroll
            for (int a = 0; a < kSmartMaxA; ++a)
                if (a < A && sw[a] < 0 && ftl_loss - s_loss >= thresh[a]) { sw[a] = t; --n_unswitched; }
        }
    }
    double s[D];
    ftl_action<D>(thl, s);                                       // :162-164
    const double comp = 0.5 * ((double)p.T - dot_seq<D>(s, vy));
    // every prefix loss above used 0.5|z.s - y| = 0.5(1 - y z.s): off by at most 0.5 * sum (||z_i||^2 - 1)+ in total
    if (!(excess <= 1e-13 * (double)p.T)) bad = 1;
    p.flags[b] = bad;
#pragma unroll
    for (int a = 0; a < kSmartMaxA; ++a) {
        if (a < A) {
            p.regret[b * A + a] = total[a] - comp;
            if (p.switch_t) p.switch_t[b * A + a] = sw[a];
        }
    }
}

template <typename TIN>
static cudaError_t launch_smart_linear_t(const SmartArgs& a, int d, cudaStream_t st) {
    const unsigned gx = (unsigned)((a.B + 127) / 128);
    switch (d) {
    case 1: smart_linear_kernel<1, TIN><<<gx, 128, 0, st>>>(a); break;
    case 2: smart_linear_kernel<2, TIN><<<gx, 128, 0, st>>>(a); break;
    case 3: smart_linear_kernel<3, TIN><<<gx, 128, 0, st>>>(a); break;
    case 4: smart_linear_kernel<4, TIN><<<gx, 128, 0, st>>>(a); break;
    case 5: smart_linear_kernel<5, TIN><<<gx, 128, 0, st>>>(a); break;
    case 6: smart_linear_kernel<6, TIN><<<gx, 128, 0, st>>>(a); break;
    case 7: smart_linear_kernel<7, TIN><<<gx, 128, 0, st>>>(a); break;
    case 8: smart_linear_kernel<8, TIN><<<gx, 128, 0, st>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int D, typename TIN>
static cudaError_t launch_smart_d(const SmartArgs& a, cudaStream_t st) {
    const int64_t Tp = (a.T + 31) & ~(int64_t)31;
    const size_t per_warp = (size_t)(D + 1) * Tp * sizeof(double);
    const size_t budget = 227 * 1024;
    // Measured (B200, d = 5): with the whole GPU's worth of sequences the warp-per-sequence kernel is 1.4-2.9x faster
    // (warp 0's serial part leaves the other three warps waiting at the barriers); with few sequences (< 4 warps
    // per SM) the CTA kernel is 1.7x (T = 1e4) to 2.5x (T = 4000) faster because it brings 4x the warps.
    bool use_cta = a.T >= kSmartCtaMinT && a.B <= 4 * 148;
    if (const char* e = getenv("OCO_SMART_CTA")) use_cta = e[0] == '1';              // tuning knob
    if (use_cta) {                                                  // long sequences: one CTA per sequence
        if (a.B > 0x7fffffff) return cudaErrorInvalidValue;
        if (per_warp <= 110 * 1024) {                               // the widened sequence fits twice per SM
            auto kern = smart_cta_kernel<D, TIN, true>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per_warp);
            if (e != cudaSuccess) return e;
            kern<<<(unsigned)a.B, kSmartCtaThreads, per_warp, st>>>(a);
        } else {
            smart_cta_kernel<D, TIN, false><<<(unsigned)a.B, kSmartCtaThreads, 0, st>>>(a);
        }
        return cudaGetLastError();
    }
    int W = (int)std::min<size_t>(kSmartWarps, per_warp ? budget / per_warp : kSmartWarps);
    // Measured (B200, d = 5, float32 input): cached wins 8-22 % for T <= 300 and loses 1.6-3.7x for T >= 500, where
    // the shared-memory footprint cuts the resident warps below what the redundant per-lane sequential part needs.
    if (per_warp > 16 * 1024) W = 0;
    if (const char* e = getenv("OCO_SMART_CACHED")) { if (e[0] == '0') W = 0; }      // tuning knob
    if (W >= 2) {                                                  // small sequence: cached kernel, W warps per CTA
        auto kern = smart_kernel<D, TIN, true>;
        const size_t smem = (size_t)W * per_warp;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)((a.B + W - 1) / W), W * 32, smem, st>>>(a);
        return cudaGetLastError();
    }
    smart_kernel<D, TIN, false><<<(unsigned)((a.B + kSmartWarps - 1) / kSmartWarps), kSmartWarps * 32, 0, st>>>(a);
    return cudaGetLastError();
}

template <typename TIN>
static cudaError_t launch_smart_t(const SmartArgs& a, int d, cudaStream_t st) {
    switch (d) {
    case 1: return launch_smart_d<1, TIN>(a, st);
    case 2: return launch_smart_d<2, TIN>(a, st);
    case 3: return launch_smart_d<3, TIN>(a, st);
    case 4: return launch_smart_d<4, TIN>(a, st);
    case 5: return launch_smart_d<5, TIN>(a, st);
    case 6: return launch_smart_d<6, TIN>(a, st);
    case 7: return launch_smart_d<7, TIN>(a, st);
    case 8: return launch_smart_d<8, TIN>(a, st);
    default: return cudaErrorInvalidValue;
    }
}

// mode 0: exact O(T^2) warp kernel; mode 1: linear-regime O(T) kernel, then the exact kernel on the
// sequences it flagged (flags: device scratch [B] ints, required for mode 1)
cudaError_t launch_smart(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                         const double* thresh_dev, const double* scale_tab, int A,
                         double* regret, int64_t* switch_t, int mode, int* flags, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (A < 1 || A > kSmartMaxA) return cudaErrorInvalidValue;
    SmartArgs a;
    a.z = z; a.y = y; a.B = B; a.T = T; a.tab = scale_tab; a.thresh = thresh_dev; a.A = A;
    a.regret = regret; a.switch_t = switch_t; a.only_flagged = nullptr; a.flags = flags;
    if (mode == 1) {
        if (!flags) return cudaErrorInvalidValue;
        cudaError_t e = in_f32 ? launch_smart_linear_t<float>(a, d, st) : launch_smart_linear_t<double>(a, d, st);
        if (e != cudaSuccess) return e;
        a.only_flagged = flags;
    }
    return in_f32 ? launch_smart_t<float>(a, d, st) : launch_smart_t<double>(a, d, st);
}

// ------------------------------------------------------------------ fp64 issue-rate probe
// Roofline denominator for the generated-on-device kernels ("must be
// microbenchmarked by the build"): 8 independent chains of alternating DMUL / DADD per thread,
// no FMA (this file is built with -fmad=false), enough warps to saturate the fp64 pipe.
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters, double m, double c) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-6 * (threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { a[k] = a[k] * m; a[k] = a[k] + c; }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456) out[0] = s;   // keep the chains alive
}
cudaError_t launch_fp64_probe(double* out, int blocks, int iters, cudaStream_t st) {
    fp64_probe_kernel<<<blocks, 256, 0, st>>>(out, iters, 0.9999999, 1e-7);
    return cudaGetLastError();
}

}  // namespace oco
