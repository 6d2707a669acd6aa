// gsweep_impl.cuh -- the fused g-path sweep (d = 1: exact candidate-sum comparator, bit-exact; 2 <= d <= 8: the
// single-pass "linear" comparator): generate + play + compare with NO second pass over the rows and the two phases
// software-pipelined.
//
//   gsweep_kernel<D, NFTL, NF>: one thread per seeded sequence of fast_algorithms.py:231-239
//   (z = standard_normal((T, d)) clipped to the unit ball, y = choice([-1, 1], T)), persistent CTAs.
//
//   numpy draws the T labels AFTER all T*d normals from the same PCG64 stream, and the ziggurat
//   consumes a data-dependent number of raw draws, so the label cursor of a sequence is only known
//   once every one of its normals has been generated: its rows must either be kept (8*d bytes per
//   step written once, read once) or generated twice.  Generating twice was built and measured in
//   round 2 (count pass + regenerate, zero DRAM bytes): 1.96e10 steps/s against 3.49e10 for the slab
//   at d = 5 -- the generator IS the cost (profiles/history/r02_gsweep_regen_experiment.md) -- so rows
//   go through a per-CTA slab in HBM, laid out [t][j][thread] (every warp access one coalesced 256-B
//   request).  What changed against sweep_kernel (sweep_impl.cuh):
//     * no comparator pass (below): rows are read once, not twice; the labels are drawn by the generating side after
//       the last row and parked 32 per word behind the rows (0.25 B per step), so the play side carries no generator;
//     * the two phases of consecutive sequences are software-pipelined inside each thread: while
//       row t of sequence k+1 is being generated (integer-ALU bound), row t of sequence k is loaded
//       and played (HBM / fp64 bound) -- two independent dependency chains per thread and the slab
//       traffic hidden under the generator, instead of the two phases running back to back;
//     * streamlined ziggurat fast path (one LDS.128 per draw) and an inlined wedge test.
//
//   Comparator without a second pass (fast_algorithms.py:113-114, 69-76).  Rows have ||z_t|| <= 1
//   (clipped, up to 2 ulp) and y_t = +-1.  Whenever the played q_t = z_t.x_t satisfies
//   sign(q_t - y_t) == -y_t and q_t != y_t -- i.e. |q_t| < 1, checked on EVERY step below -- the
//   gradient is g_t = -0.5*y_t, so theta_T = -0.5 * sum_t y_t z_t =: -0.5 v holds bit for bit (the
//   same additions in the same order, scaled by a power of two).  The comparator a = -theta_T/||theta_T||
//   = v/||v|| has |z_t.a| <= 1, hence 0.5|z_t.a - y_t| = 0.5(1 - y_t z_t.a) and
//        comp = 0.5 (T - a.v) = T/2 - ||theta_T||,        regret = (cum - T/2) + ||theta_T||.
//   cum (the played loss) is accumulated exactly as the reference does, so only comp differs from
//   the reference's sequential sum, by that sum's own rounding noise: |d regret| <= 1e-12 *
//   max(1, cum) (north_star's tolerance for d > 1; tests/test_gpu_parity.py states it).  A sequence
//   for which the per-step check ever fails (never observed; needs |q_t| >= 1) is re-evaluated by
//   the literal second pass in gsweep_exact_comp below, so the kernel never returns an unchecked
//   shortcut.
//
//   All arithmetic of the played trajectory is the bit-exact restatement of oco_device.cuh
//   (-fmad=false, reference order); theta is kept doubled (theta2 = 2*theta: the gradient step is a
//   sign flip + DADD, the scales are halved once per step -- power-of-two scalings commute with
//   rounding).
#pragma once
#include "sweep_impl.cuh"
#include <type_traits>

namespace oco {

// ---------------------------------------------------------------- generator, streamlined fast path
struct ZigTab2 {                 // shared memory, 6 KB
    ulonglong2 wk[256];          // .x = bits of wi[idx], .y = ki[idx] | 0x4330000000000000 (one LDS.128 per draw)
    double fi[256];
};

__device__ __forceinline__ void zig_tab2_load(ZigTab2& s) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s.wk[i] = make_ulonglong2(g_zig_wi[i], g_zig_ki[i] | 0x4330000000000000ULL);
        s.fi[i] = __longlong_as_double((long long)g_zig_fi[i]);
    }
    __syncthreads();
}

// Top of numpy's random_standard_normal loop for one raw r: candidate x, its layer and magnitude;
// returns the fast-accept test rabs < ki[idx].  mag = rabs | 2^52-exponent: the same word serves the
// exact int -> double conversion (one DADD) and, against the pre-or'ed ki, the comparison.
__device__ __forceinline__ bool zig_try(const ZigTab2& tb, uint64_t r, double& x, int& idx, uint64_t& mag) {
    idx = (int)(r & 0xff);
    mag = ((r >> 9) & 0x000fffffffffffffULL) | 0x4330000000000000ULL;
    const ulonglong2 wk = tb.wk[idx];
    const int sgn = (int)(((uint32_t)r << 23) & 0x80000000u);                    // bit 8 of r -> sign bit
    const double w = __hiloint2double((int)(wk.x >> 32) ^ sgn, (int)(uint32_t)wk.x);   // (+-wi): (-a)*b == -(a*b)
    x = (__longlong_as_double((long long)mag) - 4503599627370496.0) * w;
    return mag < wk.y;
}
// the test alone (count pass): no float work
__device__ __forceinline__ bool zig_test(const ZigTab2& tb, uint64_t r, int& idx, uint64_t& mag) {
    idx = (int)(r & 0xff);
    mag = ((r >> 9) & 0x000fffffffffffffULL) | 0x4330000000000000ULL;
    return mag < tb.wk[idx].y;
}

// tail sampler (idx == 0, 0.028 % of draws): out of line, state by value
struct ZigTailRet { double x; uint64_t hi, lo; };
static __device__ __noinline__ ZigTailRet zig_tail(uint64_t s_hi, uint64_t s_lo, uint64_t inc_hi, uint64_t inc_lo, uint64_t rabs) {
    Pcg64 s; s.hi = s_hi; s.lo = s_lo;
    PcgInc inc; inc.hi = inc_hi; inc.lo = inc_lo;
    ZigTailRet out;
    for (;;) {
        const double xx = -OCO_ZIG_INV_R * log1p_glibc_fma(-pcg_double(s, inc));
        const double yy = -log1p_glibc_fma(-pcg_double(s, inc));
        if (yy + yy > xx * xx) {
            out.x = ((rabs >> 8) & 1) ? -(OCO_ZIG_R + xx) : OCO_ZIG_R + xx;
            out.hi = s.hi; out.lo = s.lo;
            return out;
        }
    }
}

// The wedge comparison yv < exp(arg) in double precision: only reached when the float bracket below cannot
// decide (probability ~2e-5 per wedge test); out of line so that the caller stays small.
static __device__ __noinline__ bool zig_wedge_exact(double yv, double arg) { return yv < exp(arg); }

// Everything after a failed fast test, until a normal is accepted (numpy's for(;;) body from the
// `if (idx == 0)` on, including the redraws).  The wedge comparison is bracketed with a float
// exponential exactly as in rng_numpy.cuh::zig_slow (same decisions, see the comment there).
// A warp enters this whenever ANY of its 32 lanes fails the fast test (24 % of the draws at warp level for
// 0.85 % per lane), so it is written to be short: ~45 instructions for an accepted wedge candidate.
#ifndef OCO_ZIG_SLOW_INLINE
#define OCO_ZIG_SLOW_INLINE 1
#endif
#if OCO_ZIG_SLOW_INLINE
#define OCO_ZIG_SLOW_ATTR __forceinline__
#else
#define OCO_ZIG_SLOW_ATTR __noinline__
#endif
static __device__ OCO_ZIG_SLOW_ATTR double zig_slow2(const ZigTab2& tb, Pcg64& s, const PcgInc& inc, int idx, uint64_t mag, double x) {
    for (;;) {
        if (idx == 0) {
            const ZigTailRet o = zig_tail(s.hi, s.lo, inc.hi, inc.lo, mag & 0x000fffffffffffffULL);
            s.hi = o.hi; s.lo = o.lo;
            return o.x;
        }
        const double f1 = tb.fi[idx], f0 = tb.fi[idx - 1];
        const double yv = (f0 - f1) * pcg_double(s, inc) + f1;
        const double arg = -0.5 * x * x;
        const double ef = (double)__expf((float)arg);
        const bool sure_yes = yv < ef * (1.0 - 1e-5), sure_no = yv > ef * (1.0 + 1e-5);
        bool accept = sure_yes;
        if (!sure_yes && !sure_no) accept = zig_wedge_exact(yv, arg);
        if (accept) return x;
        const uint64_t r = pcg_next64(s, inc);              // rejected: a fresh attempt
        if (zig_try(tb, r, x, idx, mag)) return x;
    }
}

__device__ __forceinline__ double zig_normal2(const ZigTab2& tb, Pcg64& s, const PcgInc& inc) {
    const uint64_t r = pcg_next64(s, inc);
    double x; int idx; uint64_t mag;
    if (zig_try(tb, r, x, idx, mag)) return x;
    return zig_slow2(tb, s, inc, idx, mag, x);
}
// advance the stream past one normal (count pass)
__device__ __forceinline__ void zig_skip2(const ZigTab2& tb, Pcg64& s, const PcgInc& inc) {
    const uint64_t r = pcg_next64(s, inc);
    int idx; uint64_t mag;
    if (zig_test(tb, r, idx, mag)) return;
    double x; zig_try(tb, r, x, idx, mag);
    zig_slow2(tb, s, inc, idx, mag, x);
}

// one row of fast_algorithms.py:234-236: d normals, z *= 1.0 / max(||z||, 1.0)
template <int D>
__device__ __forceinline__ void gpath_row2(const ZigTab2& tb, Pcg64& s, const PcgInc& inc, double (&z)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) z[j] = zig_normal2(tb, s, inc);
    double nrm;
    if constexpr (D == 1) nrm = fabs(z[0]);
    else nrm = sqrt_of_sumsq<D>(np_pairwise_sumsq<D, double>(z), z[0]);
    const double inv = rcp_inrange(nrm > 1.0 ? nrm : 1.0);      // 1 <= norm < 1e3
#pragma unroll
    for (int j = 0; j < D; ++j) z[j] = z[j] * inv;
}

// ---------------------------------------------------------------- generator with a one-raw lookahead
// The raw stream R_0, R_1, ... of a PCG64 is fixed; only the ROLE of a raw (start of a normal, wedge uniform,
// tail uniform, label word) depends on the accept tests before it.  So the next raw can always be drawn
// before the current one has been interpreted: the 128-bit LCG step + XSL-RR output of raw n + 1 (27
// instructions, ~60 cycles of dependent latency) then overlaps the table lookup, compare and branch of raw n
// instead of waiting behind them.  `rn` is the raw already drawn but not yet consumed.
struct GenLA { Pcg64 s; uint64_t rn; };

__device__ __forceinline__ uint64_t la_take(GenLA& g, const PcgInc& inc) {
    const uint64_t r = g.rn;
    g.rn = pcg_next64(g.s, inc);
    return r;
}
__device__ __forceinline__ double la_double(GenLA& g, const PcgInc& inc) {      // Generator.random()
    return __ull2double_rn(la_take(g, inc) >> 11) * (1.0 / 9007199254740992.0);
}
struct ZigTailLARet { double x; uint64_t hi, lo, rn; };
static __device__ __noinline__ ZigTailLARet zig_tail_la(uint64_t s_hi, uint64_t s_lo, uint64_t rn, uint64_t inc_hi, uint64_t inc_lo,
                                                        uint64_t rabs) {
    GenLA g; g.s.hi = s_hi; g.s.lo = s_lo; g.rn = rn;
    PcgInc inc; inc.hi = inc_hi; inc.lo = inc_lo;
    ZigTailLARet out;
    for (;;) {
        const double xx = -OCO_ZIG_INV_R * log1p_glibc_fma(-la_double(g, inc));
        const double yy = -log1p_glibc_fma(-la_double(g, inc));
        if (yy + yy > xx * xx) {
            out.x = ((rabs >> 8) & 1) ? -(OCO_ZIG_R + xx) : OCO_ZIG_R + xx;
            out.hi = g.s.hi; out.lo = g.s.lo; out.rn = g.rn;
            return out;
        }
    }
}
// ---- ziggurat with the lookahead stream, reading the tables through a PINNED 32-bit shared-memory address: ptxas
// otherwise rematerialises the shared-window base of a static __shared__ object (S2UR SR_CgaCtaId + UMOV + ULEA) at every use
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) { asm volatile("mov.u32 %0, %0;" : "+r"(v)); return v; }
__device__ __forceinline__ ulonglong2 lds_v2u64(uint32_t a) {
    ulonglong2 v; asm("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a)); return v;
}
__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_f64_v(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ double2 lds_v2f64_v(uint32_t a) {
    double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v;
}
// `ki` receives the (pre-or'ed) threshold of the layer.  The fast test compares the HIGH words only: hi(mag) < hi(ki)
// implies mag < ki, and the 2^-20 of the draws whose high words tie go to the slow path, which repeats the test on all
// 64 bits first (zig_slow_la_a) -- same decisions, one instruction less on every draw.  `k52` is the exponent word
// 0x43300000 held in a register so that mask-and-or is ONE LOP3 (two immediates would need two).
__device__ __forceinline__ bool zig_try_a(uint32_t tba, uint32_t k52, uint64_t r, double& x, int& idx, uint64_t& mag, uint64_t& ki) {
    idx = (int)((uint32_t)r & 0xffu);
    const uint32_t mhi = (((uint32_t)(r >> 41)) & 0x000fffffu) | k52;
    mag = ((uint64_t)mhi << 32) | (uint32_t)(r >> 9);
    const ulonglong2 wk = lds_v2u64(tba + (uint32_t)idx * 16u);
    const int sgn = (int)(((uint32_t)r << 23) & 0x80000000u);                    // bit 8 of r -> sign bit
    const double w = __hiloint2double((int)(wk.x >> 32) ^ sgn, (int)(uint32_t)wk.x);   // (+-wi): (-a)*b == -(a*b)
    x = (__longlong_as_double((long long)mag) - 4503599627370496.0) * w;
    ki = wk.y;
    return mhi < (uint32_t)(wk.y >> 32);
}
static __device__ __forceinline__ double zig_slow_la_a(uint32_t tba, uint32_t k52, GenLA& g, const PcgInc& inc, int idx, uint64_t mag,
                                                        uint64_t ki, double x) {
    for (;;) {
        if (mag < ki) return x;                                // high words tied: the exact 64-bit fast test
        if (idx == 0) {
            const ZigTailLARet o = zig_tail_la(g.s.hi, g.s.lo, g.rn, inc.hi, inc.lo, mag & 0x000fffffffffffffULL);
            g.s.hi = o.hi; g.s.lo = o.lo; g.rn = o.rn;
            return o.x;
        }
        const uint32_t fa = tba + (uint32_t)offsetof(ZigTab2, fi) + (uint32_t)idx * 8u;
        const double f1 = lds_f64(fa), f0 = lds_f64(fa - 8u);
        const double yv = (f0 - f1) * la_double(g, inc) + f1;
        const double arg = -0.5 * x * x;
        const double ef = (double)__expf((float)arg);
        const bool sure_yes = yv < ef * (1.0 - 1e-5), sure_no = yv > ef * (1.0 + 1e-5);
        bool accept = sure_yes;
        if (!sure_yes && !sure_no) accept = zig_wedge_exact(yv, arg);
        if (accept) return x;
        if (zig_try_a(tba, k52, la_take(g, inc), x, idx, mag, ki)) return x;     // rejected: a fresh attempt
    }
}
__device__ __forceinline__ double zig_normal_la_a(uint32_t tba, uint32_t k52, GenLA& g, const PcgInc& inc) {
    const uint64_t r = la_take(g, inc);
    double x; int idx; uint64_t mag, ki;
    if (zig_try_a(tba, k52, r, x, idx, mag, ki)) return x;
    return zig_slow_la_a(tba, k52, g, inc, idx, mag, ki, x);
}
template <int D>
__device__ __forceinline__ void gpath_row_la_a(uint32_t tba, uint32_t k52, GenLA& g, const PcgInc& inc, double (&z)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) z[j] = zig_normal_la_a(tba, k52, g, inc);
    if constexpr (D == 1) {
        // z * fl(1/max(|z|, 1)): untouched unless |z| > 1 (z * 1.0 == z); sqrt(fl(z*z)) == |z| there, and
        // z * fl(1/|z|) == z * |fl(1/z)| (the reciprocal is odd); rcp_inrange's result is only used where 1 < |z| < 1e3
        const double rc = fabs(rcp_inrange(z[0]));
        if (fabs(z[0]) > 1.0) z[0] = z[0] * rc;
    } else {
        const double nrm = sqrt_of_sumsq<D>(np_pairwise_sumsq<D, double>(z), z[0]);
        const double inv = rcp_inrange(nrm > 1.0 ? nrm : 1.0);      // 1 <= norm < 1e3
#pragma unroll
        for (int j = 0; j < D; ++j) z[j] = z[j] * inv;
    }
}
// ---------------------------------------------------------------- literal comparator pass
// sum_t |z_t.a - y_t| for one sequence and one comparator a, in the reference's order (fast_algorithms.py:69-76), over the
// rows and packed labels of the item just played: they are still in the slab (the generating side overwrites it only
// in the next round).  Taken by every trajectory in the two-pass comparator mode (bit-exact regrets for any d <= 8), by
// the test hook OCO_SWEEP_FORCE_SECOND_PASS, and by a sequence whose single-pass check failed (never observed).
// Four rows are requested before the first is consumed (thread-private columns: every warp access is coalesced).
template <int D>
static __device__ __noinline__ double gsweep_slab_comp(const double* __restrict__ zr, const uint32_t* __restrict__ lr, int T,
                                                        const double* __restrict__ a_in) {
    double a[D];
#pragma unroll
    for (int j = 0; j < D; ++j) a[j] = a_in[j];
    double C = 0.0;
    uint32_t lw = 0;
    int t = 0;
    for (; t + 4 <= T; t += 4) {
        if ((t & 31) == 0) lw = __ldcg(lr + (int64_t)(t >> 5) * kSweepBlock);
        double z[4][D];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) z[i][j] = __ldcg(zr + ((int64_t)(t + i) * D + j) * kSweepBlock);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double y = __hiloint2double((int)((lw << 31) | 0x3ff00000u), 0);
            lw >>= 1;
            C = C + fabs(dot_seq<D>(z[i], a) - y);
        }
    }
    for (; t < T; ++t) {
        if ((t & 31) == 0) lw = __ldcg(lr + (int64_t)(t >> 5) * kSweepBlock);
        double z[D];
#pragma unroll
        for (int j = 0; j < D; ++j) z[j] = __ldcg(zr + ((int64_t)t * D + j) * kSweepBlock);
        const double y = __hiloint2double((int)((lw << 31) | 0x3ff00000u), 0);
        lw >>= 1;
        C = C + fabs(dot_seq<D>(z, a) - y);
    }
    return C;
}

// ---------------------------------------------------------------- the kernel
// Rows come back from the slab through a per-thread ring of asynchronous copies (cp.async, SASS LDGSTS): row t + 3 is
// requested at the end of step t and lands in shared memory without touching a register, so no scoreboard wait can be
// scheduled next to the request (a register prefetch was rotated by the compiler into "load, then copy at the end of
// the same step": 26 % of all stall samples on that copy).  Every thread reads only the slots it filled itself.
//
// Round-2 diet (profiles/r02_ncu_summary.md, C3: 395 -> ~300 warp-instructions per warp-step):
//   * the labels of a sequence (numpy draws them after all of its normals) are drawn by the GENERATING side right
//     after its last row and parked, 32 per word, behind the rows in the slab: the play side carries no PCG64 stream
//     (12 registers, one 50 % taken branch per step) and decodes a label with a shift;
//   * with several FTRL variants per thread the -(eta0/sqrt(t)) table rows are staged per warp in shared memory,
//     32 steps at a time (each lane copies one row with cp.async; the step reads them as broadcast LDS.128): no
//     per-step global loads, no prefetch registers (16 at NF = 8), no long-scoreboard wait on the table;
//   * d = 1: |z| <= 1, |x| <= 1, y = +-1 make sign(q - y) = -y whenever q != y, so the step +-0.5 z is formed once
//     per row from the label's sign and shared by all variants.
// The freed registers stop the compiler from rematerialising addresses every step (~50 instructions).
constexpr int kRing = 4;
// th += g iff df is neither zero nor NaN, as ONE compare and ONE predicated add (the compiler's own choice is an
// unconditional add and two selects)
__device__ __forceinline__ void add_if_nonzero(double& th, double df, double g) {
    asm("{\n\t.reg .pred p;\n\tsetp.ne.f64 p, %1, 0d0000000000000000;\n\t@p add.rn.f64 %0, %0, %2;\n\t}" : "+d"(th) : "d"(df), "d"(g));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8a(uint32_t smem_dst, const double* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef OCO_GSWEEP_MIN_BLOCKS
#define OCO_GSWEEP_MIN_BLOCKS 5
#endif
#ifndef OCO_GSWEEP_MIN_BLOCKS_NF
#define OCO_GSWEEP_MIN_BLOCKS_NF 4          // several FTRL variants per thread: more registers, fewer resident CTAs
#endif
#ifndef OCO_GSWEEP_UNROLL
#define OCO_GSWEEP_UNROLL 1                 // tuning knob: unroll factor of the step loop
#endif
constexpr int kStepUnroll = OCO_GSWEEP_UNROLL;
constexpr int kLabelsPerWord = 32;          // packed labels: bit i of word w = 1 iff y[32 w + i] = -1
constexpr int kTabBlock = 32;               // steps of the FTRL scale table staged at a time (one row per lane)
static_assert(kTabBlock == kLabelsPerWord && kTabBlock == 32, "the step loop walks blocks of 32: one label word, one scale row per lane");

// bytes of one slab (rows of one item, then its packed labels); two per resident CTA
__host__ __device__ inline size_t gsweep_slab_bytes(int64_t T, int D) {
    return (size_t)T * D * kSweepBlock * sizeof(double) + (size_t)((T + kLabelsPerWord - 1) / kLabelsPerWord) * kSweepBlock * sizeof(uint32_t);
}

template <int D, int NFTL, int NF>
__global__ void __launch_bounds__(kSweepBlock, (NF <= 1 ? OCO_GSWEEP_MIN_BLOCKS : OCO_GSWEEP_MIN_BLOCKS_NF) * kSweepBlockScale) gsweep_kernel(const SweepArgs p) {
    static_assert(D >= 1 && D <= 8, "thread-per-sequence kernel: theta lives in registers");
    // scale rows through shared memory (16-byte copies) where the static 48 KB allow it (not d = 7, 8 with NF = 8)
    constexpr bool kStageTab = (NF == 1 || NF % 2 == 0) &&
        kRing * D * kSweepBlock * 8 + (kSweepBlock / 32) * 2 * kTabBlock * NF * 8 + (int)sizeof(ZigTab2) + 256 <= 48 * 1024;
    __shared__ ZigTab2 tb;
    __shared__ double red[kSweepBlock / 32][3];
    __shared__ double ring[kRing][D][kSweepBlock];              // [slot][coordinate][thread]: conflict-free LDS.64
    __shared__ __align__(16) double sct[kStageTab ? kSweepBlock / 32 : 1][2][kStageTab ? kTabBlock : 1][kStageTab ? NF : 1];
    zig_tab2_load(tb);
    const int T = p.T, A = p.plan.A, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // 32-bit shared-memory addresses, pinned in registers (see pin_u32)
    const uint32_t tba = pin_u32((uint32_t)__cvta_generic_to_shared(&tb));
    const uint32_t k52 = pin_u32(0x43300000u);                  // exponent word of 2^52 (zig_try_a)
    const uint32_t ringa = pin_u32((uint32_t)__cvta_generic_to_shared(&ring[0][0][tid]));
    const uint32_t scta = pin_u32((uint32_t)__cvta_generic_to_shared(&sct[kStageTab ? warp : 0][0][0][0]));
    constexpr uint32_t kSlotBytes = D * kSweepBlock * 8, kRingBytes = kRing * kSlotBytes, kTabBufBytes = kTabBlock * NF * 8;
    const int groups = p.plan.groups;
    constexpr unsigned mask = 0xffffffffu;
    // two slabs per CTA: round r writes slab r & 1 (generation of item r) and reads slab (r - 1) & 1 (play of item r - 1)
    const size_t slab_bytes = gsweep_slab_bytes(T, D);
    const int64_t row_bytes = (int64_t)T * D * kSweepBlock * sizeof(double);
    char* const slab_base = reinterpret_cast<char*>(p.slab_z) + (size_t)blockIdx.x * 2 * slab_bytes;
    const int nw = (T + kLabelsPerWord - 1) / kLabelsPerWord;

    // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t n_my = p.n_items > blockIdx.x ? (p.n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // ---- the sequence being PLAYED (item r - 1)
    int64_t chunkB = 0, bB = 0; int gB = 0; bool activeB = false;
    // ---- the sequence being GENERATED (item r)
    GenLA genA; genA.s.hi = 0; genA.s.lo = 0; genA.rn = 0;
    PcgInc incA; incA.hi = 0; incA.lo = 0;

    for (int64_t r = 0; r <= n_my; ++r) {
        const bool doA = r < n_my && p.debug_phase != 2, doB = r > 0 && p.debug_phase != 1;
        int64_t chunkA = 0, bA = 0; int gA = 0; bool activeA = false;
        if (r < n_my) {
            const int64_t item = blockIdx.x + r * gridDim.x;
            chunkA = item / groups; gA = (int)(item - chunkA * groups);
            activeA = chunkA * kSweepBlock + tid < p.B;
            bA = activeA ? chunkA * kSweepBlock + tid : p.B - 1;     // idle lanes replay the last sequence
            const uint64_t ka = p.keys ? p.keys[2 * bA] : p.key_a;
            const uint64_t kc = p.keys ? p.keys[2 * bA + 1] : p.key_c0 + (uint64_t)bA;
            pcg_seed3(genA.s, incA, ka, (uint64_t)T, kc);
            genA.rn = pcg_next64(genA.s, incA);
        }
        char* const slabW = slab_base + (size_t)(r & 1) * slab_bytes;               // generation writes here
        const char* const slabR = slab_base + (size_t)((r + 1) & 1) * slab_bytes;   // play reads the other one
        double* __restrict__ zw = reinterpret_cast<double*>(slabW) + tid;
        const double* __restrict__ zr = reinterpret_cast<const double*>(slabR) + tid;
        const uint32_t* __restrict__ lr = reinterpret_cast<const uint32_t*>(slabR + row_bytes) + tid;

        // play state (fast_algorithms.py:95-97); theta2 = 2*theta, S = sum |q - y|
        double th2r[NF][D], Sr[NF];
        double th2l[D], Sl = 0.0;
        int okr[NF], okl = (int)0x80000000;                 // sign bit stays set while sign(df) == -y on every step
        bool zero_df = false;                               // some df == 0 (no step taken): theta != -0.5 v
#pragma unroll
        for (int k = 0; k < NF; ++k) {
            Sr[k] = 0.0; okr[k] = (int)0x80000000;
#pragma unroll
            for (int j = 0; j < D; ++j) th2r[k][j] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < D; ++j) th2l[j] = 0.0;
        const double* __restrict__ tp = p.tab + (int64_t)gB * T * NF;
        const bool do_ftl = NFTL > 0 && gB == 0 && p.plan.n_ftl > 0;
        // d = 1: the comparator -(1/sqrt(th*th))*th = -th*fl(1/|th|) is one of +-1, +-w (w = 1 - 2^-53) or 0, so the
        // comparator losses of ALL variants are among four running sums that depend on the sequence only, accumulated
        // in the reference's order (fast_algorithms.py:69-76): exact, and no second pass (sweep_impl.cuh)
        double cand[4] = {0.0, 0.0, 0.0, 0.0};                  // sum |c*z - y| for c = +1, -1, +w, -w
        const double kW = __longlong_as_double(0x3fefffffffffffffLL);

        // FTRL scales of the step: NF >= 2 -> staged per warp in shared memory, kTabBlock steps at a time (lane l copies
        // the row of step t0 + l); NF == 1 -> one register fetched a step ahead
        double scn1 = 0.0;
        auto stage_tab = [&](int t0, int buf) {
            if constexpr (kStageTab) {
                const int t = t0 + lane;
                if (t < T) {
                    if constexpr (NF == 1) cp_async8(&sct[warp][buf][lane][0], tp + t);
                    else {
#pragma unroll
                        for (int k = 0; k < NF; k += 2) cp_async16(&sct[warp][buf][lane][k], tp + (int64_t)t * NF + k);
                    }
                }
            }
        };
        if (doB) {
            if constexpr (kStageTab) { stage_tab(0, 0); cp_async_commit(); }
            else if (NF == 1 && T > 0) scn1 = __ldg(tp);
#pragma unroll
            for (int i = 0; i < kRing - 1; ++i) {               // rows 0 .. kRing-2 of the sequence to play: one group per row
                if (i < T) {
#pragma unroll
                    for (int j = 0; j < D; ++j) cp_async8a(ringa + i * kSlotBytes + j * (kSweepBlock * 8), zr + ((int64_t)i * D + j) * kSweepBlock);
                }
                cp_async_commit();
            }
        }
        const double* __restrict__ zrn = zr + (int64_t)(kRing - 1) * D * kSweepBlock;   // next row to request
        uint32_t wnext = (doB && T > 0) ? __ldcg(lr) : 0u;
        uint32_t slot = 0, slotn = (kRing - 1) * kSlotBytes;    // byte offsets of the ring slots of row t and of row t + kRing - 1

        // The step loop exists twice: a generic copy with run-time doA / doB (the first and the last round of a CTA, which
        // only generate / only play) and a copy for the steady state in which both are compile-time true.  In that copy no
        // uniform branch separates the generating side from the play side, so the tail of the generator (row clip: a
        // dependent reciprocal chain that feeds nothing but the slab store; for d > 1 the row norm: sqrt + reciprocal) is in
        // the same basic block as the head of the play section and is scheduled under its loads and chains.
        auto run_steps = [&](auto both_tag) {
            constexpr bool kBoth = decltype(both_tag)::value;
            const bool runA = kBoth || doA, runB = kBoth || doB;
            for (int t0 = 0; t0 < T; t0 += kTabBlock) {
                const int tn = min(kTabBlock, T - t0);
                uint32_t lw = wnext;                                // labels of steps t0 .. t0 + 31 (bit = 1: y = -1)
                if (runB) {
                    if (t0 + kLabelsPerWord < T) wnext = __ldcg(lr + (int64_t)((t0 >> 5) + 1) * kSweepBlock);
                    if constexpr (kStageTab) {
                        cp_async_wait<kRing - 2>();                 // this lane's copy of the block's scale rows has landed ...
                        __syncwarp();                               // ... and so have the other lanes' (all past the previous block)
                        stage_tab(t0 + kTabBlock, ((t0 / kTabBlock) + 1) & 1);      // joins the commit group of step t0
                    }
                }
                uint32_t sca = scta + (kStageTab ? ((t0 / kTabBlock) & 1) * kTabBufBytes : 0u);      // scale row of step t

#pragma unroll kStepUnroll
                for (int i = 0; i < tn; ++i) {
                    const int t = t0 + i;
                    // ---- generate row t of item r (fast_algorithms.py:234-236), parked in the slab
                    if (runA) {
                        double z[D];
                        gpath_row_la_a<D>(tba, k52, genA, incA, z);
    #pragma unroll
                        for (int j = 0; j < D; ++j) __stcg(zw + j * kSweepBlock, z[j]);
                        zw += D * kSweepBlock;
                    }
                    // ---- play row t of item r - 1 (fast_algorithms.py:99-111)
                    if (runB) {
                        cp_async_wait<kRing - 2>();                     // all but the two youngest groups: row t has landed
                        double z[D];
    #pragma unroll
                        for (int j = 0; j < D; ++j) z[j] = lds_f64_v(ringa + slot + j * (kSweepBlock * 8));
                        const int yhi = (int)((lw << 31) | 0x3ff00000u);    // y = -1.0 iff the label bit is set
                        lw >>= 1;
                        const double y = __hiloint2double(yhi, 0);
                        double scn[NF];
                        if constexpr (kStageTab) {
                            if constexpr (NF == 1) scn[0] = lds_f64_v(sca);
                            else {
    #pragma unroll
                                for (int k = 0; k < NF; k += 2) {
                                    const double2 v = lds_v2f64_v(sca + k * 8);
                                    scn[k] = v.x; scn[k + 1] = v.y;
                                }
                            }
                            sca += NF * 8;
                        } else {
    #pragma unroll
                            for (int k = 0; k < NF; ++k) scn[k] = (NF == 1) ? scn1 : __ldg(tp + (int64_t)t * NF + k);
                        }
                        if constexpr (D == 1) {
                            // (d = 1 keeps theta itself in th2*: one shared hz = 0.5*z per step beats halving NF scales)
                            // |z| <= 1 (clipped row), |x| <= 1 (projected action), y = +-1: q - y has the sign of -y unless it
                            // is exactly 0 (or NaN), so the gradient step of EVERY variant is gz = -y * 0.5 z or nothing
                            const double z0 = z[0], hz0 = 0.5 * z0;
                            const double gz = __hiloint2double(__double2hiint(hz0) ^ (~yhi & 0x80000000), __double2loint(hz0));
                            const double pw = z0 * kW;                      // fl(z*(-w)) = -fl(z*w), fl(-a - y) = -fl(a + y)
                            cand[0] = cand[0] + fabs(z0 - y);
                            cand[1] = cand[1] + fabs(z0 + y);
                            cand[2] = cand[2] + fabs(pw - y);
                            cand[3] = cand[3] + fabs(pw + y);
                            if constexpr (NFTL > 0) {
                                // FTL action -(1/sqrt(th*th))*th = -th*|fl(1/th)| (sqrt(fl(th*th)) == |th|), branch-free and in the
                                // same basic block as the FTRL variants below, so that its 12-deep dependent chain is scheduled
                                // under them (computed even for items whose FTL columns are not written: do_ftl is only tested
                                // at the end).  th is 0 or 2^-107 <= |th| <= T/2: every row of the g-path is 0 or a double of
                                // magnitude >= 2^-54 (ziggurat: integer * wi[idx], clipped to [-1, 1]), hence a multiple of
                                // 2^-106, and th = sum +-0.5 z.  So th*th neither under- nor overflows and rcp_inrange's range
                                // holds unless th == 0, where the reference plays 0 (any zero: z*(+-0) - y == -y exactly).
                                const double rc = rcp_inrange(th2l[0]);
                                const double x = (th2l[0] == 0.0) ? 0.0 : -fabs(rc) * th2l[0];
                                const double df = z0 * x - y;
                                Sl = Sl + fabs(df);
                                add_if_nonzero(th2l[0], df, gz);                    // zero / NaN diff: no step (fast_algorithms.py:27-34)
                            }
                            // the NF FTRL variants are independent chains (fast_algorithms.py:52-66, 105-111),
                            // handled HALF by half (variants are usually passed in eta0 order: the small-eta half almost never
                            // projects, the large-eta half almost always): per half all actions first, one vote, branch-free
                            // reciprocals the scheduler can interleave, then the steps -- half the temporaries live at a time
    #ifndef OCO_GSWEEP_D1_GROUP
    #define OCO_GSWEEP_D1_GROUP 4               // variants handled together (tuning knob: 2 needs fewer temporaries, 4 gives more ILP)
    #endif
                            constexpr int H = NF >= OCO_GSWEEP_D1_GROUP ? OCO_GSWEEP_D1_GROUP : NF;
    #pragma unroll
                            for (int h0 = 0; h0 < NF; h0 += H) {
                                double x1[H], n2v[H];
                                int top = 0;                                // max over the half of the high word of n2 (n2 >= 0 or NaN)
    #pragma unroll
                                for (int k = 0; k < H; ++k) {
                                    x1[k] = scn[h0 + k] * th2r[h0 + k][0];
                                    n2v[k] = x1[k] * x1[k];
                                    top = max(top, __double2hiint(n2v[k]));
                                }
                                if (__any_sync(mask, top >= 0x3ff00000)) {              // some n2 >= 1.0 in the warp
                                    if (__any_sync(mask, top >= 0x7c200000)) {          // n2 > ~1e290, inf or NaN: exact slow path
    #pragma unroll
                                        for (int k = 0; k < H; ++k) {
                                            double xx[1] = {x1[k]};
                                            project_unit_ball<1>(xx, n2v[k], mask);
                                            x1[k] = xx[0];
                                        }
                                    } else {                                            // sqrt(fl(x*x)) == |x|, x * (1/|x|) == x * |1/x|
    #pragma unroll
                                        for (int k = 0; k < H; ++k) {
                                            // (select on the FACTOR: the reciprocal chains of all variants stay in this basic block
                                            //  and interleave; `if (need) x *= fac` let the compiler sink one chain under a branch)
                                            const double fac = fabs(rcp_inrange(x1[k]));   // garbage where n2 <= 1: replaced by 1.0
                                            x1[k] = x1[k] * ((n2v[k] <= 1.0) ? 1.0 : fac);
                                        }
                                    }
                                }
    #pragma unroll
                                for (int k = 0; k < H; ++k) {
                                    const double df = z0 * x1[k] - y;
                                    Sr[h0 + k] = Sr[h0 + k] + fabs(df);
                                    add_if_nonzero(th2r[h0 + k][0], df, gz);
                                }
                            }
                        } else {
                        if (NFTL > 0 && do_ftl) {
                            // FTL action of theta = 0.5*theta2: n2' = 4 n2, -(1/sqrt(n2')) = 0.5 * -(1/sqrt(n2)): same x bit for bit
                            double x[D];
                            ftl_action<D>(th2l, x);
                            const double df = dot_seq<D>(z, x) - y;
                            Sl = Sl + fabs(df);
                            const int dhi = __double2hiint(df);
                            okl &= dhi ^ yhi;
                            if (df > 0.0 || df < 0.0) {
                                const int sgn = dhi & 0x80000000;
    #pragma unroll
                                for (int j = 0; j < D; ++j)
                                    th2l[j] = th2l[j] + __hiloint2double(__double2hiint(z[j]) ^ sgn, __double2loint(z[j]));
                            } else zero_df = true;
                        }
    #pragma unroll
                        for (int k = 0; k < NF; ++k) {                  // fast_algorithms.py:52-66, 105-111
                            const double sch = 0.5 * scn[k];             // (0.5*scale) * theta2 == scale * theta
                            double x[D];
    #pragma unroll
                            for (int j = 0; j < D; ++j) x[j] = sch * th2r[k][j];
                            double n2 = x[0] * x[0];
    #pragma unroll
                            for (int j = 1; j < D; ++j) n2 = n2 + x[j] * x[j];
                            project_unit_ball<D>(x, n2, mask);
                            const double df = dot_seq<D>(z, x) - y;
                            Sr[k] = Sr[k] + fabs(df);
                            const int dhi = __double2hiint(df);
                            okr[k] &= dhi ^ yhi;
                            if (df > 0.0 || df < 0.0) {
                                const int sgn = dhi & 0x80000000;
    #pragma unroll
                                for (int j = 0; j < D; ++j)
                                    th2r[k][j] = th2r[k][j] + __hiloint2double(__double2hiint(z[j]) ^ sgn, __double2loint(z[j]));
                            } else zero_df = true;
                        }
                        }
                        if (t + kRing - 1 < T) {                        // row t + 3 goes into the slot row t - 1 used
    #pragma unroll
                            for (int j = 0; j < D; ++j) cp_async8a(ringa + slotn + j * (kSweepBlock * 8), zrn + j * kSweepBlock);
                            zrn += D * kSweepBlock;
                        }
                        cp_async_commit();
                        slot = slot + kSlotBytes == kRingBytes ? 0u : slot + kSlotBytes;
                        slotn = slotn + kSlotBytes == kRingBytes ? 0u : slotn + kSlotBytes;
                        if constexpr (NF == 1 && !kStageTab) { if (t + 1 < T) scn1 = __ldg(tp + t + 1); }
                    }
                }
            }
        };
        if (doA && doB) run_steps(std::true_type{}); else run_steps(std::false_type{});

        // ---- labels of the generated item: numpy's choice([-1, 1], T) continues the stream after the last normal; one
        // 32-bit draw per label, low half of each raw first (rng_numpy.cuh::LabelCursor), 16 raws per packed word
        if (doA) {
            uint32_t* __restrict__ lwp = reinterpret_cast<uint32_t*>(slabW + row_bytes) + tid;
            for (int w = 0; w < nw; ++w) {
                uint32_t word = 0;
#pragma unroll 4
                for (int i = 0; i < kLabelsPerWord / 2; ++i) {      // (draws past label T - 1 are harmless: the stream ends here)
                    const uint64_t raw = la_take(genA, incA);
                    const uint32_t lo = (uint32_t)raw, hi = (uint32_t)(raw >> 32);
                    word = (word >> 2) | ((~lo >> 1) & 0x40000000u) | (~hi & 0x80000000u);
                }
                __stcg(lwp + (int64_t)w * kSweepBlock, word);
            }
        }
        // ---- comparator, results and per-chunk statistics of the played item
        if (r > 0) {                                            // (uniform over the CTA: the reduction below synchronises)
            const int nf = min(NF, p.plan.n_ftrl - gB * NF);
            double reg_r[NF], reg_l = 0.0;
#pragma unroll
            for (int k = 0; k < NF; ++k) reg_r[k] = 0.0;
            const double halfT = 0.5 * (double)T;
            auto finish = [&](const double (&th2)[D], double S, int ok, double& cum, double& comp, double& reg) {
                double n2 = th2[0] * th2[0];
#pragma unroll
                for (int j = 1; j < D; ++j) n2 = n2 + th2[j] * th2[j];
                const double nth = 0.5 * sqrt(n2);              // ||theta_T||
                cum = 0.5 * S;
                if constexpr (D == 1) {
                    // a == 0 (theta == 0): every term is |+-0 - y| = 1, the sum is T exactly
                    double a1[1], C = 0.0; bool found = true;
                    ftl_action<1>(th2, a1);
                    if (a1[0] == 1.0) C = cand[0]; else if (a1[0] == -1.0) C = cand[1];
                    else if (a1[0] == kW) C = cand[2]; else if (a1[0] == -kW) C = cand[3];
                    else if (a1[0] == 0.0) C = (double)T; else found = false;   // not reachable for finite theta
                    // (theta is finite on g-path data, so one of the five cases holds; a NaN / inf theta gives a NaN comparator
                    //  and a NaN loss in the reference too.  No literal pass is compiled into the d = 1 kernel: its mere presence
                    //  after the step loop cost 5 % there -- registers, spills, moves.  The test hook OCO_SWEEP_FORCE_SECOND_PASS
                    //  is served by the round-1 kernel for d = 1, see launch_sweep.)
                    comp = found ? 0.5 * C : __longlong_as_double(0x7ff8000000000000LL);
                    reg = cum - comp;
                    return;
                } else if (ok < 0 && !zero_df && !p.force_second_pass && !p.literal) {
                    comp = halfT - nth;
                    reg = (cum - halfT) + nth;
                    return;
                }
                if constexpr (D > 1) {                          // literal second pass over the slab (exact)
                    double a[D];
                    ftl_action<D>(th2, a);
                    // (the slab pointers are recomputed from the kernel arguments: anything kept alive across the step loop
                    //  for this rare path costs registers -- and moves -- there)
                    const size_t sb = gsweep_slab_bytes(p.T, D);
                    const char* const sR = reinterpret_cast<const char*>(p.slab_z) + ((size_t)blockIdx.x * 2 + ((r + 1) & 1)) * sb;
                    comp = 0.5 * gsweep_slab_comp<D>(reinterpret_cast<const double*>(sR) + threadIdx.x,
                                                     reinterpret_cast<const uint32_t*>(sR + (size_t)p.T * D * kSweepBlock * sizeof(double)) + threadIdx.x,
                                                     p.T, a);
                    reg = cum - comp;
                }
            };
            if (NFTL > 0 && do_ftl) {
                double cum, comp;
                finish(th2l, Sl, okl, cum, comp, reg_l);
                for (int c = 0; activeB && c < p.plan.n_ftl; ++c) {
                    const int64_t o = bB * A + p.plan.ftl_col[c];
                    if (p.regret) p.regret[o] = reg_l;
                    if (p.cum) p.cum[o] = cum;
                    if (p.comp) p.comp[o] = comp;
                }
            }
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                if (k < nf) {
                    double cum, comp;
                    finish(th2r[k], Sr[k], okr[k], cum, comp, reg_r[k]);
                    const int64_t o = bB * A + p.plan.ftrl_col[gB * NF + k];
                    if (activeB && p.regret) p.regret[o] = reg_r[k];
                    if (activeB && p.cum) p.cum[o] = cum;
                    if (activeB && p.comp) p.comp[o] = comp;
                }
            }
            // per-chunk (max, sum, sumsq) of the regrets, fixed order (same layout as sweep_kernel)
            if (p.partials) {
                                const int nslots = (nf > 0 ? nf : 0) + (do_ftl ? 1 : 0);
                for (int sidx = 0; sidx < nslots; ++sidx) {
                    const bool is_ftl = do_ftl && sidx == nslots - 1;
                    double rg = reg_l;
                    if (!is_ftl) {
#pragma unroll
                        for (int k = 0; k < NF; ++k) if (k == sidx) rg = reg_r[k];
                    }
                    double mx = activeB ? rg : -__longlong_as_double(0x7ff0000000000000LL);
                    double sm = activeB ? rg : 0.0, sq = activeB ? rg * rg : 0.0;
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
                            const int col = is_ftl ? p.plan.ftl_col[c] : p.plan.ftrl_col[gB * NF + sidx];
                            double* o = p.partials + (chunkB * A + col) * 3;
                            o[0] = m; o[1] = s1; o[2] = s2;
                        }
                    }
                    __syncthreads();
                }
            }
        }

        // ---- hand over: the sequence just generated is played next
        chunkB = chunkA; gB = gA; bB = bA; activeB = activeA;
    }
}

// persistent launch; two slabs per resident CTA, allocated stream-ordered
template <int D, int NFTL, int NF>
static cudaError_t launch_gsweep_cfg(SweepArgs a, cudaStream_t st) {
    auto kern = gsweep_kernel<D, NFTL, NF>;
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSweepBlock, 0);
    if (e != cudaSuccess) return e;
    const int64_t chunks = (a.B + kSweepBlock - 1) / kSweepBlock;
    a.n_items = chunks * a.plan.groups;
    int64_t grid = std::min<int64_t>(a.n_items, (int64_t)sms * std::max(per_sm, 1));
    const size_t per_cta = 2 * gsweep_slab_bytes(a.T, D);
    if (per_cta > 0) {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            free_b += scratch_cached_free_bytes();     // what the library's own pool holds idle is reusable by this call
            const int64_t fit = (int64_t)((free_b / 10) * 7 / per_cta);
            if (fit < 1) return cudaErrorMemoryAllocation;
            grid = std::min<int64_t>(grid, fit);
        }
    }
    void* slab = nullptr;
    e = scratch_alloc(&slab, (size_t)grid * per_cta + 256, st);
    if (e != cudaSuccess) return e;
    a.slab_z = slab; a.slab_y = nullptr;
    kern<<<(unsigned)grid, kSweepBlock, 0, st>>>(a);
    e = cudaGetLastError();
    cudaError_t e2 = cudaFreeAsync(slab, st);
    return e != cudaSuccess ? e : e2;
}

template <int D>
static cudaError_t launch_gsweep_d(const SweepArgs& a, cudaStream_t st) {
    const bool ftl = a.plan.n_ftl > 0;
    switch (a.plan.nf_tpl) {
    case 1: return ftl ? launch_gsweep_cfg<D, 1, 1>(a, st) : launch_gsweep_cfg<D, 0, 1>(a, st);
    case 4: return ftl ? launch_gsweep_cfg<D, 1, 4>(a, st) : launch_gsweep_cfg<D, 0, 4>(a, st);
    case 8: return ftl ? launch_gsweep_cfg<D, 1, 8>(a, st) : launch_gsweep_cfg<D, 0, 8>(a, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace oco
