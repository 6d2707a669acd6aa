// sweep_impl.cuh -- the fused generate + simulate + compare kernel (the engine's hot kernel).
//
//   sweep_kernel<Stream, D, NFTL, NF>: one thread per seeded sequence, persistent CTAs.
//   Phase 0  generate: the thread runs its numpy-compatible stream ONCE (PCG64 + ziggurat + row
//            clip) and parks the rows in a per-CTA scratch slab in HBM, laid out [t][j][thread]
//            so that every warp access is one fully coalesced 256-byte (fp64) request.
//   Phase 1  play: rows stream back (prefetched one step ahead), labels are drawn from the same
//            stream (numpy draws them after all normals) and packed 32 per word into the slab;
//            one optional FTL trajectory and NF FTRL trajectories live in registers.
//   Phase 2  compare: the comparator of fast_algorithms.py:113-114 needs the FINAL theta, so the
//            rows and label words stream back once more (D > 1).  For D == 1 the comparator can
//            only be +-1, +-(1 - 2^-53) or 0, so phase 1 keeps the four candidate loss sums and
//            phase 2 is a table lookup.
//   Inputs never exist in the reference's [B][T][d] layout and never cross PCIe; the slab is
//   private to a thread (it re-reads only what it wrote), sized by the number of resident
//   threads, not by B.  Per trajectory-step the kernel moves (8 + 16)*D/A bytes of HBM/L2 traffic
//   and issues ~9D+6 algorithmic fp64 ops: it is bound by the fp64 pipe / issue slots
//   (SURVEY.md 8d), not by HBM.
//
//   Bit-exact restatement of fast_algorithms.py:88-115 (see oco_device.cuh); the shortcuts
//   below are exact identities of IEEE-754 binary64 round-to-nearest arithmetic:
//    * loss accumulators hold S = sum |q_t - y_t| and are halved once at the end: scaling by
//      a power of two commutes with rounding, so 0.5*S == sum(0.5*|q_t - y_t|) bit for bit;
//    * the gradient step adds (+-)0.5*z by flipping the sign bit of hz = 0.5*z;
//    * for D == 1, sqrt(fl(x*x)) == |x| whenever x*x neither under- nor overflows (guarded),
//      which removes the square roots of the d = 1 row clip, projection and FTL action;
//    * FTL variants ignore eta0 (fast_algorithms.py:102-103): one FTL trajectory serves every
//      FTL column; the FTRL variants are processed NF at a time (work items = chunk x group);
//    * the projection x *= 1/||x|| (fast_algorithms.py:61-66) is applied under a warp vote with
//      the non-projecting lanes multiplying by 1.0/1.0 = 1.0 (exact), so the step has no
//      divergent branch at all.
#pragma once
#include "oco_device.cuh"
#include "oco_internal.h"
#include <algorithm>
#include "streams.cuh"

namespace oco {

#ifndef OCO_SWEEP_BLOCK
#define OCO_SWEEP_BLOCK 128
#endif
constexpr int kSweepBlock = OCO_SWEEP_BLOCK;          // sequences (threads) per CTA and per work item
constexpr int kSweepBlockScale = 128 / kSweepBlock;   // resident-CTA targets below are stated for 128-thread CTAs
#ifndef OCO_SWEEP_ROW_AHEAD
#define OCO_SWEEP_ROW_AHEAD 3
#endif
constexpr int kRowAhead = OCO_SWEEP_ROW_AHEAD;      // phase 1 prefetch distance (rows) into L1

struct SweepArgs {
    uint64_t key_a;          // first entropy word (base_seed / run_seed) when keys == nullptr
    uint64_t key_c0;         // third entropy word of sequence 0; sequence b uses key_c0 + b
    const uint64_t* keys;    // optional device [B][2] explicit (a, c) pairs
    int64_t B; int T;
    int family;
    int force_second_pass;   // test hook (see launch_sweep)
    int literal;             // gsweep_kernel, d > 1: literal comparator pass for every trajectory (two-pass mode, bit-exact)
    int debug_phase;         // gsweep_kernel timing experiments only (OCO_GSWEEP_DEBUG): 1 = generation only, 2 = play only
    double p_flip;
    const double* tab;       // [groups][T][NF] FTRL scales -(eta0/sqrt(t+1)), padded with -0.0
    void* slab_z;            // [gridDim.x][T][D][128] elem_t rows
    uint32_t* slab_y;        // [gridDim.x][ceil(T/32)][128] packed labels (bit = 1 for y = +1)
    double* regret; double* cum; double* comp;   // [B][A] or nullptr
    double* partials;        // [chunks][A][3] or nullptr
    int64_t n_items;         // chunks * groups
    SweepPlan plan;
};

// resident CTAs per SM the register allocator must allow (more warps hide the serial RNG / fp64 chains)
#ifndef OCO_SWEEP_MIN_BLOCKS
#define OCO_SWEEP_MIN_BLOCKS 5
#endif
template <class Stream, int D, int NFTL, int NF>
__global__ void __launch_bounds__(kSweepBlock, (D <= 2 ? OCO_SWEEP_MIN_BLOCKS : 4) * kSweepBlockScale) sweep_kernel(const SweepArgs p) {
    __shared__ ZigTables tb;
    __shared__ double red[kSweepBlock / 32][3];
    zig_tables_load(tb);
    typedef typename Stream::elem_t E;
    const int T = p.T, A = p.plan.A, tid = threadIdx.x;
    const int groups = p.plan.groups;
    E* __restrict__ zs = reinterpret_cast<E*>(p.slab_z) + (int64_t)blockIdx.x * T * D * kSweepBlock + tid;
    uint32_t* __restrict__ ys = p.slab_y + (int64_t)blockIdx.x * ((T + 31) / 32) * kSweepBlock + tid;

    for (int64_t item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int64_t chunk = item / groups;
        const int g = (int)(item - chunk * groups);
        // lanes past the end of the batch replay the last sequence (results discarded), so that every warp
        // stays converged and the votes below take the full mask
        const bool active = chunk * kSweepBlock + tid < p.B;
        const int64_t b = active ? chunk * kSweepBlock + tid : p.B - 1;
        constexpr unsigned mask = 0xffffffffu;
        const int nf = min(NF, p.plan.n_ftrl - g * NF);          // live FTRL variants of this group (may be <= 0)
        const bool do_ftl = NFTL > 0 && g == 0 && p.plan.n_ftl > 0;
        double reg_r[NF], reg_l = 0.0;
#pragma unroll
        for (int k = 0; k < NF; ++k) reg_r[k] = 0.0;

        {
            // ---- phase 0: generate the rows once (fast_algorithms.py:231-236 / sequence_generation.py:61-66)
            Stream st;
            const uint64_t ka = p.keys ? p.keys[2 * b] : p.key_a;
            const uint64_t kc = p.keys ? p.keys[2 * b + 1] : p.key_c0 + (uint64_t)b;
            st.init(tb, ka, T, kc, p.p_flip, p.family);
            uint32_t word = 0;
            {
                E* __restrict__ zw = zs;                                  // running pointers: no index arithmetic per step
                for (int t = 0; t < T; ++t) {
                    E ze[D]; int ybit;
                    st.row(tb, ze, ybit);
#pragma unroll
                    for (int j = 0; j < D; ++j) zw[j * kSweepBlock] = ze[j];
                    zw += D * kSweepBlock;
                    if constexpr (Stream::kProvLabels) {
                        word |= (uint32_t)ybit << (t & 31);
                        if ((t & 31) == 31 || t == T - 1) { ys[(int64_t)(t >> 5) * kSweepBlock] = word; word = 0; }
                    }
                }
            }
            st.begin_labels();

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

            // d = 1: the comparator -(1/sqrt(th*th))*th is -th*fl(1/|th|), i.e. one of +-1, +-w (w = 1 - 2^-53) or 0,
            // so the comparator losses of ALL variants are among four running sums that depend on the sequence
            // only; they are accumulated here in the reference's order (fast_algorithms.py:69-76) and phase 2
            // -- a second pass over the rows, three fp64 ops per variant and step -- is not needed.
            // (and the final labels need not be kept for it: the never-taken fallback regenerates them)
            constexpr bool kKeepLabels = !(D == 1 && !Stream::kProvLabels);
            double cand[4] = {0.0, 0.0, 0.0, 0.0};                         // sum |c*z - y| for c = +1, -1, +w, -w
            const double kW = __longlong_as_double(0x3fefffffffffffffLL);

            // ---- phase 1: play (fast_algorithms.py:99-111); rows prefetched into L1 kRowAhead steps ahead
            E zn[D];
#pragma unroll
            for (int j = 0; j < D; ++j) zn[j] = T > 0 ? zs[(int64_t)j * kSweepBlock] : (E)0;
            uint32_t prov = 0;
            const E* __restrict__ zr = zs;                                // row t; the prefetch reads row t + 1
            const double* __restrict__ tp = trow;                         // scales of step t
            // the FTRL scales of step t + 1 are fetched at the END of step t, when the step's temporaries are dead:
            // their (L1-hit) latency hides behind the label / row work at the top of the next step
            double scn[NF];
            auto load_scales = [&](const double* __restrict__ q) {
                if constexpr (NF % 2 == 0) {
#pragma unroll
                    for (int k = 0; k < NF; k += 2) {
                        const double2 v = __ldg(reinterpret_cast<const double2*>(q + k));
                        scn[k] = v.x; scn[k + 1] = v.y;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < NF; ++k) scn[k] = __ldg(q + k);
                }
            };
#pragma unroll
            for (int k = 0; k < NF; ++k) scn[k] = 0.0;
            if (T > 0) load_scales(tp);
            for (int t = 0; t < T; ++t, zr += D * kSweepBlock, tp += NF) {
                double z[D], hz[D];
#pragma unroll
                for (int j = 0; j < D; ++j) { z[j] = (double)zn[j]; hz[j] = 0.5 * z[j]; }
                // Row t + 1 is loaded at the END of this step (below), not here: held across the whole step the
                // compiler spills it and the spill store waits out the full HBM latency (27 % of all stall
                // samples before this change).  Row t + kRowAhead is pulled towards L2 / L1 now instead.
                // (D > 1 has the registers to hold the next row across the step: it is loaded here, a full step ahead.)
                if constexpr (D == 1) {
                    if (t + kRowAhead < T) {
#pragma unroll
                        for (int j = 0; j < D; ++j)
                            asm volatile("prefetch.global.L1 [%0];" :: "l"(zr + (kRowAhead * D + j) * kSweepBlock));
                    }
                } else {
                    if (t + 1 < T) {
#pragma unroll
                        for (int j = 0; j < D; ++j) zn[j] = zr[(D + j) * kSweepBlock];
                    }
                }
                if constexpr (Stream::kProvLabels) {
                    if ((t & 31) == 0) prov = ys[(int64_t)(t >> 5) * kSweepBlock];
                }
                const double y = (double)st.label(t, (int)((prov >> (t & 31)) & 1u));
                if constexpr (kKeepLabels) {                              // final labels for the phase-2 pass
                    if ((t & 31) == 0) word = 0;
                    word |= (y > 0.0 ? 1u : 0u) << (t & 31);
                    if ((t & 31) == 31 || t == T - 1) ys[(int64_t)(t >> 5) * kSweepBlock] = word;
                }

                if constexpr (D == 1) {
                    const double pw = z[0] * kW;                            // fl(z*(-w)) = -fl(z*w), fl(-a - y) = -fl(a + y)
                    cand[0] = cand[0] + fabs(z[0] - y);
                    cand[1] = cand[1] + fabs(z[0] + y);
                    cand[2] = cand[2] + fabs(pw - y);
                    cand[3] = cand[3] + fabs(pw + y);
                }

                if (NFTL > 0 && do_ftl) {
                    double x[D];
                    if constexpr (D == 1) {
                        // -(1/sqrt(th*th))*th with sqrt(fl(th*th)) == |th| and the branch-free reciprocal
                        // (theta == 0 plays -0.0 instead of 0.0: z*x is a zero either way and df = -y exactly)
                        const double n2 = thl[0] * thl[0];
                        double rc = rcp_inrange(thl[0]);                       // garbage unless 1e-290 <= n2 <= 1e290
                        if (!(n2 >= 1e-290 && n2 <= 1e290)) rc = (n2 == 0.0) ? 1.0 : 1.0 / sqrt(n2);
                        x[0] = -fabs(rc) * thl[0];
                    } else {
                        ftl_action<D>(thl, x);
                    }
                    const double df = dot_seq<D>(z, x) - y;
                    Sl = Sl + fabs(df);
                    grad_step<D>(thl, hz, df);
                }

                double sc[NF];
#pragma unroll
                for (int k = 0; k < NF; ++k) sc[k] = scn[k];
                if constexpr (D == 1 && NF > 1) {
                    // The NF variants are independent chains: form all actions first, project them
                    // together (one vote, branch-free reciprocals that the scheduler can interleave),
                    // then take all the steps.  fast_algorithms.py:52-66, 105-111
                    double x1[NF], n2v[NF];
                    int top = 0;                                             // max over k of the high word of n2 (n2 >= 0 or NaN)
#pragma unroll
                    for (int k = 0; k < NF; ++k) {
                        x1[k] = sc[k] * thr[k][0];
                        n2v[k] = x1[k] * x1[k];
                        top = max(top, __double2hiint(n2v[k]));
                    }
                    // some n2 >= 1.0 in the warp (n2 == 1.0 exactly is sorted out per lane below)
                    if (__any_sync(mask, top >= 0x3ff00000)) {
                        if (__any_sync(mask, top >= 0x7c200000)) {          // n2 > ~1e290, inf or NaN: exact slow path
#pragma unroll
                            for (int k = 0; k < NF; ++k) {
                                double xx[1] = {x1[k]};
                                project_unit_ball<1>(xx, n2v[k], mask);
                                x1[k] = xx[0];
                            }
                        } else {
                            // sqrt(fl(x*x)) == |x|, and x * (1/|x|) == x * |1/x|; lanes with n2 <= 1 keep x
#pragma unroll
                            for (int k = 0; k < NF; ++k) {
                                const double fac = fabs(rcp_inrange(x1[k]));   // garbage where n2 <= 1: not used there
                                if (!(n2v[k] <= 1.0)) x1[k] = x1[k] * fac;
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < NF; ++k) {
                        const double df = z[0] * x1[k] - y;
                        Sr[k] = Sr[k] + fabs(df);
                        grad_step<D>(thr[k], hz, df);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < NF; ++k) {                   // fast_algorithms.py:52-66, 105-111
                        double x[D];
#pragma unroll
                        for (int j = 0; j < D; ++j) x[j] = sc[k] * thr[k][j];
                        double n2 = x[0] * x[0];
#pragma unroll
                        for (int j = 1; j < D; ++j) n2 = n2 + x[j] * x[j];
                        project_unit_ball<D>(x, n2, mask);
                        const double df = dot_seq<D>(z, x) - y;
                        Sr[k] = Sr[k] + fabs(df);
                        grad_step<D>(thr[k], hz, df);
                    }
                }
                if (t + 1 < T) {
                    load_scales(tp + NF);
                    if constexpr (D == 1) {
#pragma unroll
                        for (int j = 0; j < D; ++j) zn[j] = zr[(D + j) * kSweepBlock];
                    }
                }
            }

            // ---- phase 2: comparator = FTL action of the final theta (fast_algorithms.py:113-114, 69-76)
            double ar[NF][D], al[D], Cr[NF], Cl = 0.0;
#pragma unroll
            for (int k = 0; k < NF; ++k) { ftl_action<D>(thr[k], ar[k]); Cr[k] = 0.0; }
            ftl_action<D>(thl, al);
            bool second_pass = true;
            if constexpr (D == 1) {
                second_pass = false;
                // a == 0 (theta == 0): every term is |+-0 - y| = 1, the sum is T exactly
                auto pick = [&](double a, double& C) {
                    if (a == 1.0) C = cand[0]; else if (a == -1.0) C = cand[1];
                    else if (a == kW) C = cand[2]; else if (a == -kW) C = cand[3];
                    else if (a == 0.0) C = (double)T; else second_pass = true;   // not reachable for finite theta; exact fallback
                };
#pragma unroll
                for (int k = 0; k < NF; ++k) pick(ar[k][0], Cr[k]);
                if (NFTL > 0 && do_ftl) pick(al[0], Cl);
                if (p.force_second_pass) second_pass = true;
                if (second_pass) {
                    Cl = 0.0;
#pragma unroll
                    for (int k = 0; k < NF; ++k) Cr[k] = 0.0;
                    if constexpr (!kKeepLabels) {                           // rebuild the label words from the stream
                        Stream s2;
                        s2.init(tb, ka, T, kc, p.p_flip, p.family);
                        for (int t = 0; t < T; ++t) { E tmp[D]; int yb; s2.row(tb, tmp, yb); }
                        s2.begin_labels();
                        uint32_t w2 = 0;
                        for (int t = 0; t < T; ++t) {
                            w2 |= ((double)s2.label(t, 0) > 0.0 ? 1u : 0u) << (t & 31);
                            if ((t & 31) == 31 || t == T - 1) { ys[(int64_t)(t >> 5) * kSweepBlock] = w2; w2 = 0; }
                        }
                    }
                }
            }
            // rows are fetched kPF steps at a time: the compare step is only ~5 fp64 ops per variant, so many
            // loads must be in flight per warp to cover the HBM latency
            constexpr int kPF = (D == 1) ? 8 : 4;
            const E* __restrict__ zc = zs;
            for (int t0 = 0; second_pass && t0 < T; t0 += kPF, zc += kPF * D * kSweepBlock) {
                E zb[kPF][D];
#pragma unroll
                for (int i = 0; i < kPF; ++i) {
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        zb[i][j] = (t0 + i < T) ? zc[(i * D + j) * kSweepBlock] : (E)0;
                }
                if ((t0 & 31) == 0) word = ys[(int64_t)(t0 >> 5) * kSweepBlock];
#pragma unroll
                for (int i = 0; i < kPF; ++i) {
                    if (t0 + i < T) {
                        const int t = t0 + i;
                        double z[D];
#pragma unroll
                        for (int j = 0; j < D; ++j) z[j] = (double)zb[i][j];
                        const double y = ((word >> (t & 31)) & 1u) ? 1.0 : -1.0;
                        if (NFTL > 0 && do_ftl) Cl = Cl + fabs(dot_seq<D>(z, al) - y);
#pragma unroll
                        for (int k = 0; k < NF; ++k) Cr[k] = Cr[k] + fabs(dot_seq<D>(z, ar[k]) - y);
                    }
                }
            }

            // ---- results: cum = 0.5*S, comp = 0.5*C (exact), regret = cum - comp (fast_algorithms.py:115)
            if (NFTL > 0 && do_ftl) {
                const double cum = 0.5 * Sl, comp = 0.5 * Cl;
                reg_l = cum - comp;
                for (int c = 0; active && c < p.plan.n_ftl; ++c) {
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
                    if (active && p.regret) p.regret[o] = reg_r[k];
                    if (active && p.cum) p.cum[o] = cum;
                    if (active && p.comp) p.comp[o] = comp;
                }
            }
        }

        // ---- per-chunk (max, sum, sumsq) of the regrets, fixed order -> deterministic statistics
        if (p.partials) {
            const int lane = tid & 31, warp = tid >> 5;
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
                if (tid == 0) {
                    double m = red[0][0], s1 = red[0][1], s2 = red[0][2];
                    for (int w = 1; w < kSweepBlock / 32; ++w) { m = fmax(m, red[w][0]); s1 += red[w][1]; s2 += red[w][2]; }
                    const int ncols = is_ftl ? p.plan.n_ftl : 1;
                    for (int c = 0; c < ncols; ++c) {
                        const int col = is_ftl ? p.plan.ftl_col[c] : p.plan.ftrl_col[g * NF + s];
                        double* o = p.partials + (chunk * A + col) * 3;
                        o[0] = m; o[1] = s1; o[2] = s2;
                    }
                }
                __syncthreads();
            }
        }
    }
}

// persistent launch: one slab per resident CTA; the slab is allocated stream-ordered and freed after the kernel
template <template <int> class StreamT, int D, int NFTL, int NF>
static cudaError_t launch_sweep_cfg(SweepArgs a, cudaStream_t st) {
    typedef typename StreamT<D>::elem_t E;
    auto kern = sweep_kernel<StreamT<D>, D, NFTL, NF>;
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSweepBlock, 0);
    if (e != cudaSuccess) return e;
    const int64_t chunks = (a.B + kSweepBlock - 1) / kSweepBlock;
    a.n_items = chunks * a.plan.groups;
    int64_t grid = std::min<int64_t>(a.n_items, (int64_t)sms * std::max(per_sm, 1));
    {   // never ask for more than ~70 % of the free device memory: fewer resident CTAs, same result
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            free_b += scratch_cached_free_bytes();     // what the library's own pool holds idle is reusable by this call
            const size_t per_cta = (size_t)a.T * D * kSweepBlock * sizeof(E) + (size_t)((a.T + 31) / 32) * kSweepBlock * 4 + 1;
            const int64_t fit = (int64_t)((free_b / 10) * 7 / per_cta);
            if (fit < 1) return cudaErrorMemoryAllocation;
            grid = std::min<int64_t>(grid, fit);
        }
    }
    const size_t zbytes = (size_t)grid * a.T * D * kSweepBlock * sizeof(E);
    const size_t ybytes = (size_t)grid * ((a.T + 31) / 32) * kSweepBlock * sizeof(uint32_t);
    void* slab = nullptr;
    e = scratch_alloc(&slab, zbytes + ybytes + 256, st);
    if (e != cudaSuccess) return e;
    a.slab_z = slab;
    a.slab_y = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(slab) + ((zbytes + 255) / 256) * 256);
    kern<<<(unsigned)grid, kSweepBlock, 0, st>>>(a);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(slab, st);
    return e != cudaSuccess ? e : e2;
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
