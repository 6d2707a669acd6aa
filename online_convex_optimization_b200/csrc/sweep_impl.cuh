// sweep_impl.cuh -- the fused generate + simulate + compare kernel (the engine's hot kernel).
//
//   sweep_kernel<Stream, D, NFTL, NF>: one thread per seeded sequence.  The thread owns the
//   numpy-compatible RNG cursors of its stream (registers), the state of one optional FTL
//   trajectory and NF FTRL trajectories (theta and the running sum of |q - y|, registers),
//   and regenerates the sequence for the comparator pass instead of storing it (the
//   comparator fast_algorithms.py:113-114 needs the FINAL theta).  HBM traffic: the 8*A-byte
//   result per sequence plus a broadcast read of the [T][NF] FTRL scale table (L2 resident).
//   Bound: fp64 pipe / issue slots.
//
//   Bit-exact restatement of fast_algorithms.py:88-115 (see oco_device.cuh); the shortcuts
//   below are exact identities of IEEE-754 binary64 round-to-nearest arithmetic:
//    * loss accumulators hold S = sum |q_t - y_t| and are halved once at the end: scaling by
//      a power of two commutes with rounding, so 0.5*S == sum(0.5*|q_t - y_t|) bit for bit;
//    * the gradient step adds (+-)0.5*z by flipping the sign bit of hz = 0.5*z;
//    * for D == 1, sqrt(fl(x*x)) == |x| whenever x*x neither under- nor overflows (guarded),
//      which removes the square roots of the d = 1 row clip and FTL action;
//    * FTL variants ignore eta0 (fast_algorithms.py:102-103): one FTL trajectory serves every
//      FTL column; the FTRL variants are processed NF at a time (grid.y groups) so that the
//      per-step code has no data-dependent branch except the rare projection fix-up.
#pragma once
#include "oco_device.cuh"
#include "oco_internal.h"
#include "streams.cuh"

namespace oco {

constexpr int kSweepBlock = 128;

struct SweepArgs {
    uint64_t key_a;          // first entropy word (base_seed / run_seed) when keys == nullptr
    uint64_t key_c0;         // third entropy word of sequence 0; sequence b uses key_c0 + b
    const uint64_t* keys;    // optional device [B][2] explicit (a, c) pairs
    int64_t B; int T;
    int family;
    double p_flip;
    const double* tab;       // [groups][T][NF] FTRL scales -(eta0/sqrt(t+1)), padded with -0.0
    double* regret; double* cum; double* comp;   // [B][A] or nullptr
    double* partials;        // [gridDim.x][A][3] or nullptr
    SweepPlan plan;
};

template <int D>
__device__ __forceinline__ void grad_step(double (&th)[D], const double (&hz)[D], double df) {
    // theta += g*z, g = 0.5*sign(df) (fast_algorithms.py:27-34,110-111); NaN / zero diff: no step
    if (df > 0.0 || df < 0.0) {
        const int sgn = __double2hiint(df) & 0x80000000;
#pragma unroll
        for (int j = 0; j < D; ++j)
            th[j] = th[j] + __hiloint2double(__double2hiint(hz[j]) ^ sgn, __double2loint(hz[j]));
    }
}

template <int D>
__device__ __forceinline__ void ftl_action(const double (&th)[D], double (&x)[D]) {
    // fast_algorithms.py:37-49, branch-free: 1/sqrt(0) = inf is masked by the select
    double n2 = th[0] * th[0];
#pragma unroll
    for (int j = 1; j < D; ++j) n2 = n2 + th[j] * th[j];
    const double sc = -(1.0 / sqrt_of_sumsq<D>(n2, th[0]));
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = (n2 == 0.0) ? 0.0 : sc * th[j];
}

template <class Stream, int D, int NFTL, int NF>
__global__ void __launch_bounds__(kSweepBlock) sweep_kernel(const SweepArgs p) {
    __shared__ ZigTables tb;
    __shared__ double red[kSweepBlock / 32][3];
    zig_tables_load(tb);

    const int64_t b = (int64_t)blockIdx.x * kSweepBlock + threadIdx.x;
    const bool active = b < p.B;
    const int A = p.plan.A;
    const int g = blockIdx.y;
    const int nf = min(NF, p.plan.n_ftrl - g * NF);          // live FTRL variants of this group (may be <= 0)
    const bool do_ftl = NFTL > 0 && g == 0 && p.plan.n_ftl > 0;
    const int T = p.T;
    typedef typename Stream::elem_t E;

    double reg_r[NF], reg_l = 0.0;
#pragma unroll
    for (int k = 0; k < NF; ++k) reg_r[k] = 0.0;

    if (active) {
        Stream st;
        const uint64_t ka = p.keys ? p.keys[2 * b] : p.key_a;
        const uint64_t kc = p.keys ? p.keys[2 * b + 1] : p.key_c0 + (uint64_t)b;
        st.init(tb, ka, T, kc, p.p_flip, p.family);

        double thr[NF][D], Sr[NF];
        double thl[D], Sl = 0.0;
#pragma unroll
        for (int k = 0; k < NF; ++k) {
            Sr[k] = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) thr[k][j] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < D; ++j) thl[j] = 0.0;
        const double* __restrict__ trow = p.tab + (int64_t)g * T * NF;

        // ---- pass 1: play (fast_algorithms.py:99-111)
        for (int t = 0; t < T; ++t) {
            E ze[D]; E ye;
            st.next(tb, t, ze, ye);
            double z[D], hz[D];
#pragma unroll
            for (int j = 0; j < D; ++j) { z[j] = (double)ze[j]; hz[j] = 0.5 * z[j]; }
            const double y = (double)ye;

            if (NFTL > 0 && do_ftl) {
                double x[D];
                ftl_action<D>(thl, x);
                const double df = dot_seq<D>(z, x) - y;
                Sl = Sl + fabs(df);
                grad_step<D>(thl, hz, df);
            }

            double sc[NF];
            if constexpr (NF % 2 == 0) {
#pragma unroll
                for (int k = 0; k < NF; k += 2) {
                    const double2 v = __ldg(reinterpret_cast<const double2*>(trow + (int64_t)t * NF + k));
                    sc[k] = v.x; sc[k + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < NF; ++k) sc[k] = __ldg(trow + (int64_t)t * NF + k);
            }
            double x[NF][D], n2[NF];
            bool any = false;
#pragma unroll
            for (int k = 0; k < NF; ++k) {                   // fast_algorithms.py:52-60
#pragma unroll
                for (int j = 0; j < D; ++j) x[k][j] = sc[k] * thr[k][j];
                n2[k] = x[k][0] * x[k][0];
#pragma unroll
                for (int j = 1; j < D; ++j) n2[k] = n2[k] + x[k][j] * x[k][j];
                any = any || !(n2[k] <= 1.0);
            }
            if (any) {                                        // projection onto the unit ball (:61-66), rare on random labels
#pragma unroll
                for (int k = 0; k < NF; ++k) {
                    if (!(n2[k] <= 1.0)) {
                        const double fac = 1.0 / sqrt_of_sumsq<D>(n2[k], x[k][0]);
#pragma unroll
                        for (int j = 0; j < D; ++j) x[k][j] = x[k][j] * fac;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                const double df = dot_seq<D>(z, x[k]) - y;
                Sr[k] = Sr[k] + fabs(df);
                grad_step<D>(thr[k], hz, df);
            }
        }

        // ---- pass 2: comparator = FTL action of the final theta (fast_algorithms.py:113-114, 69-76)
        double ar[NF][D], al[D], Cr[NF], Cl = 0.0;
#pragma unroll
        for (int k = 0; k < NF; ++k) { ftl_action<D>(thr[k], ar[k]); Cr[k] = 0.0; }
        ftl_action<D>(thl, al);
        st.rewind();
        for (int t = 0; t < T; ++t) {
            E ze[D]; E ye;
            st.next(tb, t, ze, ye);
            double z[D];
#pragma unroll
            for (int j = 0; j < D; ++j) z[j] = (double)ze[j];
            const double y = (double)ye;
            if (NFTL > 0 && do_ftl) Cl = Cl + fabs(dot_seq<D>(z, al) - y);
#pragma unroll
            for (int k = 0; k < NF; ++k) Cr[k] = Cr[k] + fabs(dot_seq<D>(z, ar[k]) - y);
        }

        // ---- results: cum = 0.5*S, comp = 0.5*C (exact), regret = cum - comp (fast_algorithms.py:115)
        if (NFTL > 0 && do_ftl) {
            const double cum = 0.5 * Sl, comp = 0.5 * Cl;
            reg_l = cum - comp;
            for (int c = 0; c < p.plan.n_ftl; ++c) {
                const int64_t o = b * A + p.plan.ftl_col[c];
                if (p.regret) p.regret[o] = reg_l;
                if (p.cum) p.cum[o] = cum;
                if (p.comp) p.comp[o] = comp;
            }
        }
#pragma unroll
        for (int k = 0; k < NF; ++k) {
            if (k < nf) {
                const double cum = 0.5 * Sr[k], comp = 0.5 * Cr[k];
                reg_r[k] = cum - comp;
                const int64_t o = b * A + p.plan.ftrl_col[g * NF + k];
                if (p.regret) p.regret[o] = reg_r[k];
                if (p.cum) p.cum[o] = cum;
                if (p.comp) p.comp[o] = comp;
            }
        }
    }

    // ---- per-block (max, sum, sumsq) of the regrets, fixed order -> deterministic statistics
    if (p.partials) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int nslots = (nf > 0 ? nf : 0) + (do_ftl ? 1 : 0);
        for (int s = 0; s < nslots; ++s) {
            const bool is_ftl = do_ftl && s == nslots - 1;
            double r = reg_l;
            if (!is_ftl) {
#pragma unroll
                for (int k = 0; k < NF; ++k) if (k == s) r = reg_r[k];
            }
            double mx = active ? r : -__longlong_as_double(0x7ff0000000000000LL);
            double sm = active ? r : 0.0, sq = active ? r * r : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, off));
                sm = sm + __shfl_down_sync(0xffffffffu, sm, off);
                sq = sq + __shfl_down_sync(0xffffffffu, sq, off);
            }
            if (lane == 0) { red[warp][0] = mx; red[warp][1] = sm; red[warp][2] = sq; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double m = red[0][0], s1 = red[0][1], s2 = red[0][2];
                for (int w = 1; w < kSweepBlock / 32; ++w) { m = fmax(m, red[w][0]); s1 += red[w][1]; s2 += red[w][2]; }
                const int ncols = is_ftl ? p.plan.n_ftl : 1;
                for (int c = 0; c < ncols; ++c) {
                    const int col = is_ftl ? p.plan.ftl_col[c] : p.plan.ftrl_col[g * NF + s];
                    double* o = p.partials + ((int64_t)blockIdx.x * A + col) * 3;
                    o[0] = m; o[1] = s1; o[2] = s2;
                }
            }
            __syncthreads();
        }
    }
}

template <template <int> class StreamT, int D, int NFTL, int NF>
static cudaError_t launch_sweep_cfg(const SweepArgs& a, cudaStream_t st) {
    const unsigned gx = (unsigned)((a.B + kSweepBlock - 1) / kSweepBlock);
    sweep_kernel<StreamT<D>, D, NFTL, NF><<<dim3(gx, (unsigned)a.plan.groups), kSweepBlock, 0, st>>>(a);
    return cudaGetLastError();
}

template <template <int> class StreamT, int D>
static cudaError_t launch_sweep_d(const SweepArgs& a, cudaStream_t st) {
    const bool ftl = a.plan.n_ftl > 0;
    switch (a.plan.nf_tpl) {
    case 1: return ftl ? launch_sweep_cfg<StreamT, D, 1, 1>(a, st) : launch_sweep_cfg<StreamT, D, 0, 1>(a, st);
    case 4: return ftl ? launch_sweep_cfg<StreamT, D, 1, 4>(a, st) : launch_sweep_cfg<StreamT, D, 0, 4>(a, st);
    case 8: return ftl ? launch_sweep_cfg<StreamT, D, 1, 8>(a, st) : launch_sweep_cfg<StreamT, D, 0, 8>(a, st);
    default: return cudaErrorInvalidValue;
    }
}

template <template <int> class StreamT>
static cudaError_t launch_sweep_s(const SweepArgs& a, int d, cudaStream_t st) {
    switch (d) {
    case 1: return launch_sweep_d<StreamT, 1>(a, st);
    case 2: return launch_sweep_d<StreamT, 2>(a, st);
    case 3: return launch_sweep_d<StreamT, 3>(a, st);
    case 4: return launch_sweep_d<StreamT, 4>(a, st);
    case 5: return launch_sweep_d<StreamT, 5>(a, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace oco
