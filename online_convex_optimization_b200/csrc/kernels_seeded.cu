// kernels_seeded.cu -- host-side glue and small kernels around the seeded families.
//
//  make_sweep_plan / launch_sweep   canonical variant layout + dispatch of the fused sweep kernel
//      (sweep_impl.cuh; instantiated in kernels_sweep_gpath.cu / kernels_sweep_iid.cu)
//  fill_kernel<Stream, D, TOUT>     write the sequences to HBM in the reference layout
//      (z[B][T][d], y[B][T]) for the streamed kernels (BASELINE config 4) and for callers that
//      want the arrays; rows first, then labels, in numpy's draw order
//  sweep_table_kernel / ftrl_scale_table_kernel   -(eta0/sqrt(t)) tables (IEEE sqrt and divide)
//  stats_finalize_kernel            fixed-order reduction of the per-chunk (max, sum, sumsq)
//  rng_probe_kernel                 known-answer probes of the device generator
#include "sweep_impl.cuh"

#include <mutex>

namespace oco {

// ------------------------------------------------------------------ private scratch pool (one per device)
namespace {
constexpr int kMaxDev = 64;
cudaMemPool_t g_pool[kMaxDev];
std::once_flag g_pool_once[kMaxDev];
cudaError_t g_pool_err[kMaxDev];

cudaMemPool_t scratch_pool(int dev, cudaError_t* err) {
    std::call_once(g_pool_once[dev], [dev]() {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        g_pool_err[dev] = cudaMemPoolCreate(&g_pool[dev], &props);
        if (g_pool_err[dev] == cudaSuccess) {
            unsigned long long thr = ~0ull;
            g_pool_err[dev] = cudaMemPoolSetAttribute(g_pool[dev], cudaMemPoolAttrReleaseThreshold, &thr);
        }
    });
    *err = g_pool_err[dev];
    return g_pool[dev];
}
}  // namespace

cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDev) return cudaMallocAsync(p, bytes, st);
    cudaMemPool_t pool = scratch_pool(dev, &e);
    if (e != cudaSuccess) return e;
    return cudaMallocFromPoolAsync(p, bytes, pool, st);
}

cudaError_t scratch_trim() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= kMaxDev) return e;
    cudaMemPool_t pool = scratch_pool(dev, &e);
    if (e != cudaSuccess) return e;
    return cudaMemPoolTrimTo(pool, 0);
}

size_t scratch_cached_free_bytes() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= kMaxDev) return 0;
    cudaMemPool_t pool = scratch_pool(dev, &e);
    if (e != cudaSuccess) return 0;
    unsigned long long reserved = 0, used = 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) != cudaSuccess) return 0;
    if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) != cudaSuccess) return 0;
    return reserved > used ? (size_t)(reserved - used) : 0;
}

// One 256-thread block per variant; thread l folds chunks l, l + 256, ... in order, then a fixed shared-memory
// tree: the association order depends only on (nblocks), so the statistics are reproducible run to run.
__global__ void __launch_bounds__(256) stats_finalize_kernel(const double* __restrict__ partials, int nblocks, int A,
                                                             int64_t n_runs, double* __restrict__ stats) {
    __shared__ double sm[3][256];
    const int a = blockIdx.x, l = threadIdx.x;
    double m = -__longlong_as_double(0x7ff0000000000000LL), s1 = 0.0, s2 = 0.0;
    for (int b = l; b < nblocks; b += 256) {
        const double* o = partials + ((int64_t)b * A + a) * 3;
        m = fmax(m, o[0]); s1 += o[1]; s2 += o[2];
    }
    sm[0][l] = m; sm[1][l] = s1; sm[2][l] = s2;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (l < w) {
            sm[0][l] = fmax(sm[0][l], sm[0][l + w]);
            sm[1][l] = sm[1][l] + sm[1][l + w];
            sm[2][l] = sm[2][l] + sm[2][l + w];
        }
        __syncthreads();
    }
    if (l == 0) { stats[a * 4 + 0] = sm[0][0]; stats[a * 4 + 1] = sm[1][0]; stats[a * 4 + 2] = sm[2][0]; stats[a * 4 + 3] = (double)n_runs; }
}

__global__ void ftrl_scale_table_kernel(double* __restrict__ tab, const Etas eta0, int64_t T, int A) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * A) return;
    const int64_t t = i / A; const int a = (int)(i - t * A);
    int64_t tt = t + 1; if (tt < 1) tt = 1;
    tab[i] = -(eta0.e[a] / sqrt((double)tt));     // fast_algorithms.py:54, IEEE sqrt and divide
}

cudaError_t launch_ftrl_scale_table(double* tab, const Etas& eta0, int64_t T, int A, cudaStream_t st) {
    const int64_t n = T * A;
    if (n <= 0) return cudaSuccess;
    ftrl_scale_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tab, eta0, T, A);
    return cudaGetLastError();
}

__global__ void select_candidates_kernel(const double* __restrict__ regret, int64_t n, int stride, const double* __restrict__ stats_max,
                                         double margin, int64_t* __restrict__ idx, int* __restrict__ count, int cap) {
    const double lim = stats_max[0] - margin;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (regret[i * stride] >= lim) {
            const int k = atomicAdd(count, 1);
            if (k < cap) idx[k] = i;
        }
    }
}
cudaError_t launch_select_candidates(const double* regret, int64_t n, int stride, const double* stats_max, double margin,
                                     int64_t* idx, int* count, int cap, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int), st);
    if (e != cudaSuccess || n <= 0) return e;
    const unsigned gx = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
    select_candidates_kernel<<<gx, 256, 0, st>>>(regret, n, stride, stats_max, margin, idx, count, cap);
    return cudaGetLastError();
}

int sweep_gpath_blocks(int64_t n_runs) { return (int)((n_runs + kSweepBlock - 1) / kSweepBlock); }


void make_sweep_plan(const Variants& v, const Etas& e, SweepPlan* pl) {
    pl->A = v.A; pl->n_ftl = 0; pl->n_ftrl = 0;
    for (int a = 0; a < kMaxVariants + 8; ++a) { pl->ftrl_col[a] = 0; pl->ftrl_eta[a] = 0.0; }
    for (int a = 0; a < kMaxVariants; ++a) pl->ftl_col[a] = 0;
    for (int a = 0; a < v.A; ++a) {
        if (v.alg_flag[a] == 0) { pl->ftrl_col[pl->n_ftrl] = a; pl->ftrl_eta[pl->n_ftrl] = e.e[a]; ++pl->n_ftrl; }
        else pl->ftl_col[pl->n_ftl++] = a;
    }
    pl->nf_tpl = pl->n_ftrl <= 1 ? 1 : (pl->n_ftrl <= 4 ? 4 : 8);
    if (const char* e = getenv("OCO_SWEEP_NF")) { const int v = atoi(e); if (v == 1 || v == 4 || v == 8) pl->nf_tpl = v; }   // tuning knob
    pl->groups = pl->n_ftrl == 0 ? 1 : (pl->n_ftrl + pl->nf_tpl - 1) / pl->nf_tpl;
}

__global__ void sweep_table_kernel(double* __restrict__ tab, const SweepPlan pl, int64_t T) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)pl.groups * pl.nf_tpl * T;
    if (i >= n) return;
    const int k = (int)(i % pl.nf_tpl);
    const int64_t gt = i / pl.nf_tpl;
    const int64_t t = gt % T; const int g = (int)(gt / T);
    tab[i] = -(pl.ftrl_eta[g * pl.nf_tpl + k] / sqrt((double)(t + 1)));   // fast_algorithms.py:54
}
cudaError_t launch_sweep_table(double* tab, const SweepPlan& plan, int64_t T, cudaStream_t st) {
    const int64_t n = sweep_table_elems(plan, T);
    if (n <= 0) return cudaSuccess;
    sweep_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tab, plan, T);
    return cudaGetLastError();
}

cudaError_t launch_sweep(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys, int64_t T, int64_t B, int d,
                         double p_flip, const SweepPlan& plan, const double* sweep_tab,
                         double* regret, double* cum, double* comp, double* partials, int comparator_mode, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (T > 0x7fffffff) return cudaErrorInvalidValue;
    SweepArgs a;
    a.key_a = key_a; a.key_c0 = key_c0; a.keys = keys; a.T = (int)T; a.B = B; a.family = family; a.p_flip = p_flip;
    a.tab = sweep_tab; a.regret = regret; a.cum = cum; a.comp = comp; a.partials = partials; a.plan = plan;
    {   // test hook: take the literal second pass even where the d = 1 candidate sums apply (tests/test_gpu_parity.py)
        const char* e = getenv("OCO_SWEEP_FORCE_SECOND_PASS");
        a.force_second_pass = (e && e[0] == '1') ? 1 : 0;
    }
    a.debug_phase = 0;
    if (const char* e = getenv("OCO_GSWEEP_DEBUG")) a.debug_phase = atoi(e);     // timing experiments: one phase only
    a.literal = comparator_mode == 1 ? 0 : 1;
    const char* legacy = getenv("OCO_SWEEP_LEGACY");                             // A/B knob: the round-1 three-phase kernel (d <= 5)
    if (!legacy) legacy = getenv("OCO_SWEEP_D1_LEGACY");
    // (d = 1 with the test hook: the literal second pass lives in the round-1 kernel only, see gsweep_impl.cuh)
    const bool use_legacy = (legacy && legacy[0] == '1' && d <= 5) || (d == 1 && a.force_second_pass);
    if (family == FAM_GPATH && !use_legacy) {
        // software-pipelined generate / play kernel (gsweep_impl.cuh), any d <= 8: d = 1 keeps the exact candidate-sum
        // comparator (either mode, bit-exact); d >= 2: OCO_COMPARATOR_LINEAR uses T/2 - ||theta_T|| (no second pass over
        // the rows), the two-pass mode re-reads the item's rows from the slab for the literal comparator loss (bit-exact)
        return d <= 3 ? launch_gsweep_lo(a, d, st) : (d <= 5 ? launch_gsweep_hi(a, d, st) : launch_gsweep_wide(a, d, st));
    }
    if (family == FAM_GPATH) return launch_sweep_gpath(a, d, st);
    if (family == FAM_IID || family == FAM_NOISY) return launch_sweep_iid(a, d, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_stats_finalize(const double* partials, int nblocks, int A, int64_t n_runs, double* stats, cudaStream_t st) {
    stats_finalize_kernel<<<A, 256, 0, st>>>(partials, nblocks, A, n_runs, stats);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ fill kernels
struct FillArgs {
    uint64_t key_a, key_c0; const uint64_t* keys;
    int64_t T, B; int family; double p_flip; int64_t block_len;
    void* z; void* y;
};

template <class Stream, int D, typename TOUT>
__global__ void __launch_bounds__(128) fill_kernel(const FillArgs p) {
    __shared__ ZigTables tb;
    zig_tables_load(tb);
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    typedef typename Stream::elem_t E;
    Stream st;
    const uint64_t ka = p.keys ? p.keys[2 * b] : p.key_a;
    const uint64_t kc = p.keys ? p.keys[2 * b + 1] : p.key_c0 + (uint64_t)b;
    st.init(tb, ka, p.T, kc, p.p_flip, p.family);
    TOUT* z = (TOUT*)p.z + b * p.T * D;
    TOUT* y = (TOUT*)p.y + b * p.T;
    const int T = (int)p.T;
    for (int t = 0; t < T; ++t) {                 // all normals first (numpy order), provisional label parked in y
        E ze[D]; int ybit;
        st.row(tb, ze, ybit);
#pragma unroll
        for (int j = 0; j < D; ++j) z[(int64_t)t * D + j] = (TOUT)ze[j];
        y[t] = (TOUT)ybit;
    }
    st.begin_labels();
    for (int t = 0; t < T; ++t) y[t] = (TOUT)st.label(t, (int)y[t]);
}

// sequence_generation.py:24-28 (flip) and :36-47 (switching two leaders); any d.
template <typename TOUT>
__global__ void fill_pattern_kernel(int family, int64_t B, int64_t T, int d, int64_t block_len, TOUT* z, TOUT* y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * T) return;
    const int64_t t = i % T;
    for (int j = 0; j < d; ++j) z[i * d + j] = (TOUT)(j == 0 ? 1.0 : 0.0);
    if (family == FAM_FLIP) y[i] = (TOUT)(((t + 1) % 2) ? 1.0 : -1.0);
    else y[i] = (TOUT)(((t / block_len) % 2) ? -1.0 : 1.0);
}

template <template <int> class StreamT, int D>
static cudaError_t launch_fill_d(const FillArgs& a, int out_f32, cudaStream_t st) {
    const unsigned gx = (unsigned)((a.B + 127) / 128);
    if (out_f32) fill_kernel<StreamT<D>, D, float><<<gx, 128, 0, st>>>(a);
    else fill_kernel<StreamT<D>, D, double><<<gx, 128, 0, st>>>(a);
    return cudaGetLastError();
}
template <template <int> class StreamT>
static cudaError_t launch_fill_s(const FillArgs& a, int d, int out_f32, cudaStream_t st) {
    switch (d) {
    case 1: return launch_fill_d<StreamT, 1>(a, out_f32, st);
    case 2: return launch_fill_d<StreamT, 2>(a, out_f32, st);
    case 3: return launch_fill_d<StreamT, 3>(a, out_f32, st);
    case 4: return launch_fill_d<StreamT, 4>(a, out_f32, st);
    case 5: return launch_fill_d<StreamT, 5>(a, out_f32, st);
    case 8: return launch_fill_d<StreamT, 8>(a, out_f32, st);
    case 16: return launch_fill_d<StreamT, 16>(a, out_f32, st);
    case 32: return launch_fill_d<StreamT, 32>(a, out_f32, st);
    case 64: return launch_fill_d<StreamT, 64>(a, out_f32, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fill(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys, int64_t B, int64_t T, int d,
                        double p_flip, int64_t block_len, int out_f32, void* z, void* y, cudaStream_t st) {
    if (B <= 0 || T <= 0) return cudaSuccess;
    if (family == FAM_FLIP || family == FAM_SWITCH) {
        const int64_t n = B * T;
        if (out_f32) fill_pattern_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(family, B, T, d, block_len, (float*)z, (float*)y);
        else fill_pattern_kernel<double><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(family, B, T, d, block_len, (double*)z, (double*)y);
        return cudaGetLastError();
    }
    {   // g-path: one WARP per sequence (jump-ahead PCG64, coalesced stores), any d <= 128; OCO_FILL_LEGACY=1 keeps the
        // thread-per-sequence kernel (A/B knob)
        const char* legacy = getenv("OCO_FILL_LEGACY");
        if (fill_warp_eligible(family, d) && T <= 0x7fffffff && !(legacy && legacy[0] == '1'))
            return launch_fill_warp(key_a, key_c0, keys, B, T, d, out_f32, z, y, st);
    }
    FillArgs a;
    a.key_a = key_a; a.key_c0 = key_c0; a.keys = keys; a.T = T; a.B = B; a.family = family; a.p_flip = p_flip;
    a.block_len = block_len; a.z = z; a.y = y;
    if (family == FAM_GPATH) return launch_fill_s<GPathStream>(a, d, out_f32, st);
    if (family == FAM_IID || family == FAM_NOISY) return launch_fill_s<IidStream>(a, d, out_f32, st);
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ RNG probes (known-answer tests)
__global__ void rng_probe_kernel(int what, uint64_t a, uint64_t b, uint64_t c, int64_t n_skip, int64_t n, void* out) {
    __shared__ ZigTables tb;
    zig_tables_load(tb);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Pcg64 s; PcgInc inc;
    pcg_seed3(s, inc, a, b, c);
    if (what == 4) {                       // SeedSequence state words
        uint64_t w[4]; seedseq3(a, b, c, w);
        for (int i = 0; i < 4 && i < n; ++i) ((uint64_t*)out)[i] = w[i];
        return;
    }
    for (int64_t i = 0; i < n_skip; ++i) zig_skip(tb, s, inc);
    if (what == 0) for (int64_t i = 0; i < n; ++i) ((uint64_t*)out)[i] = pcg_next64(s, inc);
    else if (what == 1) for (int64_t i = 0; i < n; ++i) ((double*)out)[i] = zig_normal(tb, s, inc);
    else if (what == 2) for (int64_t i = 0; i < n; ++i) ((double*)out)[i] = pcg_double(s, inc);
    else if (what == 3) { LabelCursor lc; lc.s = s; lc.cur = 0; for (int64_t i = 0; i < n; ++i) ((double*)out)[i] = lc.next(i, inc); }
    else if (what == 5) for (int64_t i = 0; i < n; ++i) ((double*)out)[i] = log1p_glibc_fma(-pcg_double(s, inc));
}
// what == 6: rcp_inrange(r) vs the compiler's 1.0 / r over n_skip values per thread of n threads' worth
// of pseudo-random doubles (both signs, exponents spread over [2^-1000, 2^1000] and, every other
// draw, over the [1, 2^12) range the kernels actually divide by); out[0] = number of mismatching bit patterns.
__global__ void rcp_check_kernel(uint64_t seed, int64_t per_thread, unsigned long long* mismatches) {
    uint64_t x = seed + 0x9e3779b97f4a7c15ull * (uint64_t)(blockIdx.x * (int64_t)blockDim.x + threadIdx.x + 1);
    unsigned long long bad = 0;
    for (int64_t i = 0; i < per_thread; ++i) {
        x += 0x9e3779b97f4a7c15ull;                       // splitmix64
        uint64_t v = x;
        v = (v ^ (v >> 30)) * 0xbf58476d1ce4e5b9ull;
        v = (v ^ (v >> 27)) * 0x94d049bb133111ebull;
        v ^= v >> 31;
        const uint64_t mant = v & 0xfffffffffffffull, sign = v >> 63;
        const uint64_t e11 = (v >> 52) & 0x7ff;
        const uint64_t ex = (i & 1) ? 1023 + (e11 % 12) : 23 + (e11 % 2001);   // biased exponent
        const double r = __longlong_as_double((long long)((sign << 63) | (ex << 52) | mant));
        const double a = rcp_inrange(r), b = 1.0 / r;
        bad += __double_as_longlong(a) != __double_as_longlong(b);
    }
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t launch_rng_probe(int what, uint64_t a, uint64_t b, uint64_t c, int64_t n_skip, int64_t n, void* out, cudaStream_t st) {
    if (what == 6) {
        cudaError_t e = cudaMemsetAsync(out, 0, (size_t)n * 8, st);
        if (e != cudaSuccess) return e;
        rcp_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, n_skip, (unsigned long long*)out);
        return cudaGetLastError();
    }
    rng_probe_kernel<<<1, 32, 0, st>>>(what, a, b, c, n_skip, n, out);
    return cudaGetLastError();
}

}  // namespace oco
