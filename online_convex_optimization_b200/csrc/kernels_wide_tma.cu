// kernels_wide_tma.cu -- the HBM-bound streamed path (BASELINE config 4: d = 64 constrained FTRL on
// streamed loss vectors): one warp per (sequence, variant) trajectory, rows delivered by the TMA
// engine.
//
// Data path.  In the reference layout z[B][T][d] the rows t0 .. t0+TT-1 of one sequence are one
// contiguous run of TT*d*sizeof(TIN) bytes, so each tile is ONE 1-D bulk async copy
// (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, SASS UBLKCP) plus one 16/32-byte
// bulk copy for the TT labels, both completing on the stage's mbarrier.  Every warp owns a private
// ring of kStages such tiles in shared memory; lane 0 is the producer (arm barrier, issue the two
// copies), all 32 lanes are consumers (mbarrier.try_wait.parity, then conflict-free 8-byte LDS:
// lane l reads elements l, l+32, ...).  A stage is recycled after __syncwarp(), so kStages-1 tiles
// (~6 KB at d = 64) are in flight per warp while one is being consumed: with 32 resident warps per
// SM that is ~200 KB of loads in flight per SM, far more than latency x bandwidth requires.
// No register staging, no LDG in the steady state.
//
// Arithmetic (tolerance 1e-12, the reduction order differs from the reference's left-to-right
// sum): theta / x are distributed K = d/32 elements per lane; the three warp reductions a round
// needs (||x||^2 for the projection, z.x for the prediction, ||z||^2 for the linear-comparator
// precondition) are issued together as independent xor-butterflies so that their shuffle
// latencies overlap, and the projection factor is applied to the reduced scalar
// (q = fac * (z.x_unprojected)) instead of re-reducing.
//
// Comparator.  OCO_COMPARATOR_LINEAR: one pass, comp = 0.5*(T - a.v), v = sum y_t z_t, valid when
// every row has ||z_t|| <= 1 and y_t = +-1 (checked on the fly; otherwise the warp re-streams the
// sequence).  OCO_COMPARATOR_TWO_PASS: the literal second pass of fast_algorithms.py:113-114.
//
// Bytes per trajectory-step: (d+1)*sizeof(TIN) (520 B at d = 64 fp64) -- the roofline figure of
// SURVEY.md 8(d) -- once in LINEAR mode, twice in TWO_PASS mode.
#include <cstdlib>
#include "oco_device.cuh"
#include "oco_internal.h"

namespace oco {

namespace {

constexpr int kWarps = 8;      // warps (trajectories) per CTA
constexpr int kTT = 4;         // rows per tile: TT*sizeof(TIN) is a multiple of 16 for fp32 and fp64
constexpr int kStages = 3;


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct WideTmaArgs {
    const void* z; const void* y;
    int64_t B; int T; int d;
    const double* tab;        // [T][A]
    double* regret; double* cum; double* comp; double* theta_final;
    int comparator_mode;
    Variants v;
};

// two fp64 warp sums for the price of ~1.2: after the first exchange the low half-warp carries `a`
// and the high half-warp carries `b`; four more exchanges finish both, one more redistributes.
__device__ __forceinline__ void warp_sum2(double& a, double& b, int lane) {
    const bool hi = (lane & 16) != 0;
    double keep = hi ? b : a;
    const double send = hi ? a : b;
    keep = keep + __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) keep = keep + __shfl_xor_sync(0xffffffffu, keep, off);
    const double other = __shfl_xor_sync(0xffffffffu, keep, 16);
    a = hi ? other : keep;
    b = hi ? keep : other;
}
__device__ __forceinline__ double warp_sum1(double a) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a = a + __shfl_xor_sync(0xffffffffu, a, off);
    return a;
}

// One warp per trajectory; D (multiple of 32) is a template parameter so that every shared-memory
// address is an immediate offset from one per-lane base pointer.
template <int D, typename TIN>
__global__ void __launch_bounds__(kWarps * 32, (D <= 64 ? 4 : (D <= 128 ? 2 : 1))) sim_wide_tma_kernel(const WideTmaArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[kWarps][kStages];
    constexpr int K = D / 32;
    constexpr uint32_t z_tile_bytes = kTT * D * sizeof(TIN);
    constexpr uint32_t y_tile_bytes = kTT * sizeof(TIN);
    constexpr uint32_t stage_bytes = z_tile_bytes + y_tile_bytes;

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int A = p.v.A, T = p.T;
    const int64_t w = (int64_t)blockIdx.x * kWarps + wib;
    unsigned char* warp_smem = smem_raw + (size_t)wib * kStages * stage_bytes;
    uint64_t* my_bars = bars[wib];
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&my_bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    if (w >= p.B * A) return;
    const int64_t b = w / A;
    const int a = (int)(w - b * A);
    const bool ftrl = p.v.alg_flag[a] == 0;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * (int64_t)T * D;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * (int64_t)T;

    const int ntiles = T / kTT;
    int cs = 0, is = 0;                        // consumer / producer stage cursors
    uint32_t cparity = 0;

    auto issue = [&](int tile) {               // lane 0 only
        unsigned char* dst = warp_smem + is * stage_bytes;
        mbar_arrive_expect_tx(&my_bars[is], stage_bytes);
        bulk_g2s(dst, zb + (int64_t)tile * (kTT * D), z_tile_bytes, &my_bars[is]);
        bulk_g2s(dst + z_tile_bytes, yb + (int64_t)tile * kTT, y_tile_bytes, &my_bars[is]);
        is = (is + 1 == kStages) ? 0 : is + 1;
    };

    double th[K], vy[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { th[k] = 0.0; vy[k] = 0.0; }
    double S = 0.0;                            // sum |q - y| ; cum_loss = 0.5 * S (exact)
    bool linear_ok = true;
    const bool want_linear = p.comparator_mode == 1;
    const double* __restrict__ tabp = p.tab + a;

    auto play_row = [&](const double (&z)[K], double y) {
        double x[K], n2 = 0.0, zx = 0.0;
        if (ftrl) {
            const double sc = __ldg(tabp);
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = sc * th[k];
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) x[k] = th[k];        // direction only; normalised below
        }
        tabp += A;
#pragma unroll
        for (int k = 0; k < K; ++k) { n2 = n2 + x[k] * x[k]; zx = zx + z[k] * x[k]; }
        warp_sum2(n2, zx, lane);
        double q = zx;
        if (ftrl) { if (!(n2 <= 1.0)) q = zx * (1.0 / sqrt(n2)); }              // fast_algorithms.py:61-66 (warp-uniform)
        else q = (n2 == 0.0) ? 0.0 : zx * -(1.0 / sqrt(n2));                    // fast_algorithms.py:44-49
        const double df = q - y;
        S = S + fabs(df);
        const double g = df > 0.0 ? 0.5 : (df < 0.0 ? -0.5 : 0.0);
#pragma unroll
        for (int k = 0; k < K; ++k) th[k] = th[k] + g * z[k];
        if (want_linear) {
            // precondition of the linear comparator: ||z_t||^2 <= 1 (+1e-6) and y = +-1.  The squared
            // norm is reduced in 2^-24 fixed point with ONE integer REDUX instead of ten shuffles.
            double zz = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) zz = zz + z[k] * z[k];
            const unsigned tot = __reduce_add_sync(0xffffffffu, __double2uint_rz(fmin(zz, 4.0) * 16777216.0));
            if (tot > 16777216u + 17u || !(y == 1.0 || y == -1.0)) linear_ok = false;
#pragma unroll
            for (int k = 0; k < K; ++k) vy[k] = vy[k] + y * z[k];
        }
    };

    double act[K], C = 0.0;
    auto compare_row = [&](const double (&z)[K], double y) {
        double q = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) q = q + z[k] * act[k];
        C = C + fabs(warp_sum1(q) - y);
    };

    auto stream_pass = [&](auto&& row_fn) {
        if (lane == 0) {
            const int pre = ntiles < kStages ? ntiles : kStages;
            for (int i = 0; i < pre; ++i) issue(i);
        }
        for (int tile = 0; tile < ntiles; ++tile) {
            mbar_wait(&my_bars[cs], cparity);
            const TIN* zt = reinterpret_cast<const TIN*>(warp_smem + cs * stage_bytes) + lane;
            const TIN* yt = reinterpret_cast<const TIN*>(warp_smem + cs * stage_bytes + z_tile_bytes);
#pragma unroll
            for (int r = 0; r < kTT; ++r) {
                double z[K];
#pragma unroll
                for (int k = 0; k < K; ++k) z[k] = (double)zt[r * D + 32 * k];
                row_fn(z, (double)yt[r]);
            }
            if (++cs == kStages) { cs = 0; cparity ^= 1u; }
            __syncwarp();
            if (lane == 0 && tile + kStages < ntiles) issue(tile + kStages);
        }
        for (int t = ntiles * kTT; t < T; ++t) {              // ragged tail: direct loads
            double z[K];
#pragma unroll
            for (int k = 0; k < K; ++k) z[k] = (double)zb[(int64_t)t * D + lane + 32 * k];
            row_fn(z, (double)yb[t]);
        }
    };

    stream_pass(play_row);

    {   // comparator = FTL action of the final theta (fast_algorithms.py:113)
        double n2 = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) n2 = n2 + th[k] * th[k];
        n2 = warp_sum1(n2);
        const double sc = (n2 == 0.0) ? 0.0 : -(1.0 / sqrt(n2));
#pragma unroll
        for (int k = 0; k < K; ++k) act[k] = (n2 == 0.0) ? 0.0 : sc * th[k];
    }
    double comp;
    if (want_linear && linear_ok) {
        double av = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) av = av + act[k] * vy[k];
        comp = 0.5 * ((double)T - warp_sum1(av));
    } else {
        stream_pass(compare_row);
        comp = 0.5 * C;
    }
    const double cum = 0.5 * S;
    if (lane == 0) {
        if (p.regret) p.regret[w] = cum - comp;
        if (p.cum) p.cum[w] = cum;
        if (p.comp) p.comp[w] = comp;
    }
    if (p.theta_final) {
#pragma unroll
        for (int k = 0; k < K; ++k) p.theta_final[w * D + lane + 32 * k] = th[k];
    }
}

template <int D, typename TIN>
cudaError_t launch_k(const WideTmaArgs& a, cudaStream_t st) {
    const int64_t nw = a.B * a.v.A;
    const unsigned gx = (unsigned)((nw + kWarps - 1) / kWarps);
    const size_t smem = (size_t)kWarps * kStages * (kTT * (D + 1) * sizeof(TIN));
    auto kern = sim_wide_tma_kernel<D, TIN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<gx, kWarps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <typename TIN>
cudaError_t launch_t(const WideTmaArgs& a, cudaStream_t st) {
    switch (a.d) {
    case 32: return launch_k<32, TIN>(a, st);
    case 64: return launch_k<64, TIN>(a, st);
    case 128: return launch_k<128, TIN>(a, st);
    case 256: return launch_k<256, TIN>(a, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace

// The bulk-copy path needs every tile source 16-byte aligned: base pointers, the per-sequence
// strides T*d*es and T*es, and the row size d*es must all be multiples of 16.
bool wide_tma_eligible(const void* z, const void* y, int in_f32, int64_t T, int d) {
    const size_t es = in_f32 ? 4 : 8;
    return ((uintptr_t)z % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((size_t)d * es) % 16 == 0 &&
           ((size_t)T * es) % 16 == 0 && T < 0x7fffffff &&
           (d == 32 || d == 64 || d == 128 || d == 256);
}

cudaError_t launch_sim_wide_tma(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                                const Variants& v, const double* scale_tab,
                                double* regret, double* cum, double* comp, double* theta_final, int comparator_mode,
                                cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    WideTmaArgs a;
    a.z = z; a.y = y; a.B = B; a.T = (int)T; a.d = d; a.tab = scale_tab;
    a.regret = regret; a.cum = cum; a.comp = comp; a.theta_final = theta_final;
    a.comparator_mode = comparator_mode; a.v = v;
    return in_f32 ? launch_t<float>(a, st) : launch_t<double>(a, st);
}

}  // namespace oco
