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
#include <type_traits>
#include "oco_device.cuh"
#include "oco_internal.h"

namespace oco {

namespace {

constexpr int kWarps = 8;      // warps (trajectories) per CTA
constexpr int kTT = 4;         // rows per tile: TT*sizeof(TIN) is a multiple of 16 for fp32 and fp64
constexpr int kStages = 2;


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

// Reductions confined to groups of G lanes (G lanes cooperate on one trajectory).
// group_sum2: two fp64 sums for the price of ~1.2 -- after the first exchange the low half of the
// group carries `a` and the high half carries `b`; log2(G)-1 more exchanges finish both, one more
// redistributes.
template <int G>
__device__ __forceinline__ void group_sum2(double& a, double& b, int sl) {
    const bool hi = (sl & (G / 2)) != 0;
    double keep = hi ? b : a;
    const double send = hi ? a : b;
    keep = keep + __shfl_xor_sync(0xffffffffu, send, G / 2);
#pragma unroll
    for (int off = G / 4; off > 0; off >>= 1) keep = keep + __shfl_xor_sync(0xffffffffu, keep, off);
    const double other = __shfl_xor_sync(0xffffffffu, keep, G / 2);
    a = hi ? other : keep;
    b = hi ? keep : other;
}
// Four fp64 sums at once (the squared norms of the four rows of a tile), result q left in quarter q of the group
// only (lanes sl / (G/4) == q): log2(G) + 1 exchanges instead of 4 * log2(G), and no redistribution.
template <int G>
__device__ __forceinline__ double group_sum4_quartered(double p0, double p1, double p2, double p3, int sl) {
    static_assert(G >= 4, "four quarters");
    const bool hi = (sl & (G / 2)) != 0, up = (sl & (G / 4)) != 0;
    // exchange 1: the low half takes rows 0, 1, the high half rows 2, 3
    double u = hi ? p2 : p0, v = hi ? p3 : p1;
    u = u + __shfl_xor_sync(0xffffffffu, hi ? p0 : p2, G / 2);
    v = v + __shfl_xor_sync(0xffffffffu, hi ? p1 : p3, G / 2);
    // exchange 2: within each half the lower quarter takes u, the upper quarter v
    double keep = (up ? v : u) + __shfl_xor_sync(0xffffffffu, up ? u : v, G / 4);
#pragma unroll
    for (int off = G / 8; off > 0; off >>= 1) keep = keep + __shfl_xor_sync(0xffffffffu, keep, off);
    return keep;
}
template <int G>
__device__ __forceinline__ double group_sum1(double a) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) a = a + __shfl_xor_sync(0xffffffffu, a, off);
    return a;
}

// ---- per-trajectory state (G lanes hold K = D/G elements each, element j = sl + G*k)
// State is kept doubled: th2 = 2*theta accumulates +-z_t (a sign flip and one DADD per element; the
// factor 0.5 of the gradient is folded into the per-step scale, an exact power-of-two scaling),
// vy = sum y_t z_t likewise (labels are +-1 whenever the linear comparator is usable).
template <int K>
struct WideState {
    double th2[K], vy[K], act[K];
    double S, C;                 // sum |q - y| over the played / comparator actions (losses are 0.5 * these)
    double excess;               // linear comparator: sum_t max(0, ||z_t||^2 - 1) of this lane's quarter's rows
    bool linear_ok;              // labels are +-1 so far
    const double* tabp;          // walks the [T][A] FTRL scale table
};

__device__ __forceinline__ double flip_sign(double v, int sgn) {
    return __hiloint2double(__double2hiint(v) ^ sgn, __double2loint(v));
}

// one round of play; MODE 0: every trajectory of the warp is FTRL, 1: all FTL, 2: mixed (per-lane select)
template <int MODE, int G, int K>
__device__ __forceinline__ void play_row(WideState<K>& st, const double (&z)[K], double y, bool ftrl, bool want_linear,
                                         int A, int sl, int grp, double& zz_part) {
    const bool is_ftrl = MODE == 0 || (MODE == 2 && ftrl);
    double x[K];
    if (MODE != 1) {
        const double sc = 0.5 * __ldg(st.tabp);              // x = (0.5*scale) * (2*theta)
        st.tabp += A;
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = (MODE == 0 || ftrl) ? sc * st.th2[k] : st.th2[k];
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = st.th2[k];        // FTL: direction only, normalised below
    }
    double n2 = x[0] * x[0], zx = z[0] * x[0];
#pragma unroll
    for (int k = 1; k < K; ++k) { n2 = n2 + x[k] * x[k]; zx = zx + z[k] * x[k]; }
    group_sum2<G>(n2, zx, sl);
    double q = zx;
    if (is_ftrl) { if (!(n2 <= 1.0)) q = zx * (1.0 / sqrt(n2)); }               // fast_algorithms.py:61-66
    else q = (n2 == 0.0) ? 0.0 : zx * -(1.0 / sqrt(n2));                        // fast_algorithms.py:44-49
    const double df = q - y;
    st.S = st.S + fabs(df);
    if (df > 0.0 || df < 0.0) {                              // theta += 0.5*sign(df)*z  <=>  th2 += sign(df)*z
        const int sgn = __double2hiint(df) & 0x80000000;
#pragma unroll
        for (int k = 0; k < K; ++k) st.th2[k] = st.th2[k] + flip_sign(z[k], sgn);
    }
    if (want_linear) {
        // Precondition of the linear comparator: y = +-1 and ||z_t|| <= 1, the latter up to a BUDGET: with |z_t.a| <=
        // ||z_t|| the identity 0.5|z_t.a - y| = 0.5(1 - y z_t.a) is off by at most (||z_t|| - 1)+ <= 0.5 (||z_t||^2 - 1)+
        // per row, so the comparator loss is within 0.5 * excess of the literal pass; excess <= 1e-13 * T keeps it
        // inside 1e-12 * (T/2).  float64 rows clipped to the ball (fast_algorithms.py:234-236) stay 100x below the
        // budget; float32-normalised rows (||z||^2 up to 1 + 1e-7) exceed it and take the literal second pass.
        // This lane's share of ||z_t||^2 is returned; the caller reduces four rows at a time (group_sum4_quartered).
        double zz = z[0] * z[0];
#pragma unroll
        for (int k = 1; k < K; ++k) zz = zz + z[k] * z[k];
        zz_part = zz;
        if (!(y == 1.0 || y == -1.0)) st.linear_ok = false;
        const int ysgn = __double2hiint(y) & 0x80000000;
#pragma unroll
        for (int k = 0; k < K; ++k) st.vy[k] = st.vy[k] + flip_sign(z[k], ysgn);
    }
}

template <int G, int K>
__device__ __forceinline__ void compare_row(WideState<K>& st, const double (&z)[K], double y) {
    double q = z[0] * st.act[0];
#pragma unroll
    for (int k = 1; k < K; ++k) q = q + z[k] * st.act[k];
    st.C = st.C + fabs(group_sum1<G>(q) - y);
}

// ---- the per-warp TMA ring
template <int D, int G, typename TIN>
struct WidePipe {
    static constexpr int NT = 32 / G;
    static constexpr uint32_t z_tile_bytes = kTT * D * sizeof(TIN);
    static constexpr uint32_t y_tile_bytes = kTT * sizeof(TIN);
    static constexpr uint32_t traj_bytes = z_tile_bytes + y_tile_bytes;
    static constexpr uint32_t stage_bytes = traj_bytes * NT;
    unsigned char* smem; uint64_t* bars;
    const TIN* z; const TIN* y;       // batch base pointers
    int64_t w0; int A, T, nvalid;
    int cs, is; uint32_t cparity;

    __device__ __forceinline__ void issue(int tile) {        // lane 0 only: the NT tiles of a stage complete on one barrier
        unsigned char* dst = smem + is * stage_bytes;
        mbar_arrive_expect_tx(&bars[is], traj_bytes * (uint32_t)nvalid);
#pragma unroll
        for (int g = 0; g < NT; ++g) {
            if (g < nvalid) {
                const int64_t bg = (w0 + g) / A;
                bulk_g2s(dst + g * traj_bytes, z + (bg * T + (int64_t)tile * kTT) * D, z_tile_bytes, &bars[is]);
                bulk_g2s(dst + g * traj_bytes + z_tile_bytes, y + bg * T + (int64_t)tile * kTT, y_tile_bytes, &bars[is]);
            }
        }
        is = (is + 1 == kStages) ? 0 : is + 1;
    }
};

// PASS 0/1/2: play with that MODE; PASS 3: comparator pass
template <int PASS, int D, int G, typename TIN>
__device__ __forceinline__ void stream_pass(WidePipe<D, G, TIN>& pp, WideState<D / G>& st, const TIN* zb, const TIN* yb,
                                            bool ftrl, bool want_linear, bool valid, int lane, int sl, int grp) {
    constexpr int K = D / G;
    typedef WidePipe<D, G, TIN> P;
    const int ntiles = pp.T / kTT;
    if (lane == 0) {
        const int pre = ntiles < kStages ? ntiles : kStages;
        for (int i = 0; i < pre; ++i) pp.issue(i);
    }
    for (int tile = 0; tile < ntiles; ++tile) {
        mbar_wait(&pp.bars[pp.cs], pp.cparity);
        const unsigned char* base = pp.smem + pp.cs * P::stage_bytes + (valid ? grp : 0) * P::traj_bytes;
        const TIN* zt = reinterpret_cast<const TIN*>(base) + sl;
        const TIN* yt = reinterpret_cast<const TIN*>(base + P::z_tile_bytes);
        double zzp[kTT];
#pragma unroll
        for (int r = 0; r < kTT; ++r) {
            double z[K];
            zzp[r] = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) z[k] = (double)zt[r * D + G * k];
            if (PASS == 3) compare_row<G, K>(st, z, (double)yt[r]);
            else play_row<PASS, G, K>(st, z, (double)yt[r], ftrl, want_linear, pp.A, sl, grp, zzp[r]);
        }
        if (PASS != 3 && want_linear) {                      // squared norms of the tile's four rows, one per quarter of the group
            static_assert(kTT == 4, "group_sum4_quartered reduces four rows");
            const double zz = group_sum4_quartered<G>(zzp[0], zzp[1], zzp[2], zzp[3], sl);
            st.excess = st.excess + fmax(zz - 1.0, 0.0);
        }
        if (++pp.cs == kStages) { pp.cs = 0; pp.cparity ^= 1u; }
        __syncwarp();
        if (lane == 0 && tile + kStages < ntiles) pp.issue(tile + kStages);
    }
    for (int t = ntiles * kTT; t < pp.T; ++t) {              // ragged tail: direct loads
        double z[K];
#pragma unroll
        for (int k = 0; k < K; ++k) z[k] = (double)zb[(int64_t)t * D + sl + G * k];
        if (PASS == 3) compare_row<G, K>(st, z, (double)yb[t]);
        else {
            double zzp = 0.0;
            play_row<PASS, G, K>(st, z, (double)yb[t], ftrl, want_linear, pp.A, sl, grp, zzp);
            if (want_linear) {                               // ragged tail rows: plain reduction, counted in quarter 0
                const double zz = group_sum1<G>(zzp);
                if (sl < G / 4) st.excess = st.excess + fmax(zz - 1.0, 0.0);
            }
        }
    }
}

// G lanes per trajectory, NT = 32/G trajectories per warp.  The per-step warp-level overhead (shuffles,
// selects, barrier polling, loop control) is shared by the NT trajectories of a warp, which is what
// brings the instruction count per trajectory-step under the HBM-roofline budget (~90 issue slots at
// d = 64).  D is a template parameter so that every shared-memory address is an immediate offset.
template <int D, int G, typename TIN>
__global__ void __launch_bounds__(kWarps * 32, (D / G <= 2 ? 4 : (D / G <= 4 ? 3 : 1))) sim_wide_tma_kernel(const WideTmaArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[kWarps][kStages];
    constexpr int K = D / G, NT = 32 / G;
    typedef WidePipe<D, G, TIN> P;

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int grp = lane / G, sl = lane % G;
    const int A = p.v.A, T = p.T;
    const int64_t total = p.B * A;
    const int64_t w0 = ((int64_t)blockIdx.x * kWarps + wib) * NT;
    P pp;
    pp.smem = smem_raw + (size_t)wib * kStages * P::stage_bytes;
    pp.bars = bars[wib];
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&pp.bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    if (w0 >= total) return;
    pp.z = (const TIN*)p.z; pp.y = (const TIN*)p.y; pp.w0 = w0; pp.A = A; pp.T = T;
    pp.nvalid = (int)((total - w0) < NT ? (total - w0) : NT);
    pp.cs = 0; pp.is = 0; pp.cparity = 0;
    const int64_t w = w0 + grp;
    const bool valid = grp < pp.nvalid;
    const int64_t b = valid ? w / A : 0;
    const int a = valid ? (int)(w - b * A) : 0;
    const bool ftrl = p.v.alg_flag[a] == 0;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * (int64_t)T * D;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * (int64_t)T;
    const bool want_linear = p.comparator_mode == 1;

    WideState<K> st;
#pragma unroll
    for (int k = 0; k < K; ++k) { st.th2[k] = 0.0; st.vy[k] = 0.0; st.act[k] = 0.0; }
    st.S = 0.0; st.C = 0.0; st.excess = 0.0; st.linear_ok = true; st.tabp = p.tab + a;

    // all trajectories of the warp run the same algorithm? then the step has no per-lane algorithm select
    if (__all_sync(0xffffffffu, ftrl || !valid)) stream_pass<0, D, G, TIN>(pp, st, zb, yb, ftrl, want_linear, valid, lane, sl, grp);
    else if (__all_sync(0xffffffffu, !ftrl || !valid)) stream_pass<1, D, G, TIN>(pp, st, zb, yb, ftrl, want_linear, valid, lane, sl, grp);
    else stream_pass<2, D, G, TIN>(pp, st, zb, yb, ftrl, want_linear, valid, lane, sl, grp);

    {   // comparator = FTL action of the final theta (fast_algorithms.py:113); scale-invariant in theta
        double n2 = st.th2[0] * st.th2[0];
#pragma unroll
        for (int k = 1; k < K; ++k) n2 = n2 + st.th2[k] * st.th2[k];
        n2 = group_sum1<G>(n2);
        const double sc = (n2 == 0.0) ? 0.0 : -(1.0 / sqrt(n2));
#pragma unroll
        for (int k = 0; k < K; ++k) st.act[k] = (n2 == 0.0) ? 0.0 : sc * st.th2[k];
    }
    // the second pass shares the warp's barriers: taken if ANY trajectory of the warp needs it
    bool use_linear = false;
    if (want_linear) {       // total excess of the sequence: one value per quarter -> sum of the four quarters
        double e = st.excess;
        e = e + __shfl_xor_sync(0xffffffffu, e, G / 4);
        e = e + __shfl_xor_sync(0xffffffffu, e, G / 2);
        use_linear = st.linear_ok && e <= 1e-13 * (double)T;
    }
    if (__any_sync(0xffffffffu, valid && !use_linear)) stream_pass<3, D, G, TIN>(pp, st, zb, yb, ftrl, want_linear, valid, lane, sl, grp);
    double comp;
    if (use_linear) {
        double av = st.act[0] * st.vy[0];
#pragma unroll
        for (int k = 1; k < K; ++k) av = av + st.act[k] * st.vy[k];
        comp = 0.5 * ((double)T - group_sum1<G>(av));
    } else {
        comp = 0.5 * st.C;
    }
    const double cum = 0.5 * st.S;
    if (valid && sl == 0) {
        if (p.regret) p.regret[w] = cum - comp;
        if (p.cum) p.cum[w] = cum;
        if (p.comp) p.comp[w] = comp;
    }
    if (valid && p.theta_final) {
#pragma unroll
        for (int k = 0; k < K; ++k) p.theta_final[w * D + sl + G * k] = 0.5 * st.th2[k];
    }
}

template <int D, int G, typename TIN>
cudaError_t launch_k(const WideTmaArgs& a, cudaStream_t st) {
    constexpr int NT = 32 / G;
    const int64_t nw = (a.B * a.v.A + NT - 1) / NT;
    const unsigned gx = (unsigned)((nw + kWarps - 1) / kWarps);
    const size_t smem = (size_t)kWarps * kStages * NT * (kTT * (D + 1) * sizeof(TIN));
    auto kern = sim_wide_tma_kernel<D, G, TIN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<gx, kWarps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <typename TIN>
cudaError_t launch_t(const WideTmaArgs& a, cudaStream_t st) {
    static const int g64 = [] { const char* e = getenv("OCO_WIDE_G"); return e ? atoi(e) : 16; }();   // tuning knob
    switch (a.d) {
    case 32: return launch_k<32, 8, TIN>(a, st);
    case 64: return g64 == 32 ? launch_k<64, 32, TIN>(a, st) : (g64 == 8 ? launch_k<64, 8, TIN>(a, st) : launch_k<64, 16, TIN>(a, st));
    case 128: return launch_k<128, 32, TIN>(a, st);
    case 256: return launch_k<256, 32, TIN>(a, st);
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
