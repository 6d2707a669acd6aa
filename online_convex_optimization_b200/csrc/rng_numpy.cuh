// rng_numpy.cuh -- device-side generator that reproduces the reference's numpy draws.
//
// The reference keys every random stream as
//   np.random.Generator(np.random.PCG64(np.random.SeedSequence([base_seed, T, run])))
// (fast_algorithms.py:254-257, algorithms.py:177-180) and draws standard_normal /
// choice / random from it (fast_algorithms.py:231-239, sequence_generation.py:54-90).
// This header restates, for one CUDA thread per stream:
//   SeedSequence.generate_state   (numpy/random/bit_generator.pyx)
//   PCG64 setseq-128 XSL-RR       (numpy/random/src/pcg64/pcg64.h)
//   random_standard_normal        (256-layer ziggurat, distributions.c)
//   next_double, buffered next_uint32 (Generator.random / Generator.choice)
//   glibc 2.39 log1p (x86_64 FMA variant) used by the ziggurat tail
// All of it is integer work plus a handful of fp64 ops; nothing is fused except
// where glibc fuses (explicit fma()).  Compile with -fmad=false.
#pragma once
#include <stdint.h>
#include "ziggurat_tables.h"

namespace oco {

// ---------------------------------------------------------------- ziggurat tables
// Kept in global memory (read-only, 6 KB) and staged to shared memory by each CTA.
__device__ const unsigned long long g_zig_wi[256] = OCO_ZIG_WI_BITS_INIT;
__device__ const unsigned long long g_zig_ki[256] = OCO_ZIG_KI_INIT;
__device__ const unsigned long long g_zig_fi[256] = OCO_ZIG_FI_BITS_INIT;

// constant-memory copies for the (single-lane, divergent) slow path: no shared-memory pointer has
// to cross the call boundary and the accesses are uniform per executing lane
__constant__ unsigned long long c_zig_wi[256] = OCO_ZIG_WI_BITS_INIT;
__constant__ unsigned long long c_zig_ki[256] = OCO_ZIG_KI_INIT;
__constant__ unsigned long long c_zig_fi[256] = OCO_ZIG_FI_BITS_INIT;

struct ZigTables {          // shared-memory copy, 6 KB
    double wi[256];
    unsigned long long ki[256];
    double fi[256];
};

__device__ __forceinline__ void zig_tables_load(ZigTables& s) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s.wi[i] = __longlong_as_double((long long)g_zig_wi[i]);
        s.ki[i] = g_zig_ki[i];
        s.fi[i] = __longlong_as_double((long long)g_zig_fi[i]);
    }
    __syncthreads();
}

// ---------------------------------------------------------------- SeedSequence
__device__ __forceinline__ uint32_t ss_hashmix(uint32_t v, uint32_t& hc) {
    v ^= hc; hc *= 0x931e8875u; v *= hc; v ^= v >> 16; return v;
}
__device__ __forceinline__ uint32_t ss_mix(uint32_t x, uint32_t y) {
    uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y; r ^= r >> 16; return r;
}
// SeedSequence([a, b, c]).generate_state(4, uint64); every int contributes its
// little-endian uint32 words (at least one).
__device__ inline void seedseq3(uint64_t a, uint64_t b, uint64_t c, uint64_t out[4]) {
    uint32_t ent[6]; int n = 0;
    ent[n++] = (uint32_t)a; if (a >> 32) ent[n++] = (uint32_t)(a >> 32);
    ent[n++] = (uint32_t)b; if (b >> 32) ent[n++] = (uint32_t)(b >> 32);
    ent[n++] = (uint32_t)c; if (c >> 32) ent[n++] = (uint32_t)(c >> 32);
    uint32_t pool[4], hc = 0x43b0d7e5u;
#pragma unroll
    for (int i = 0; i < 4; ++i) pool[i] = ss_hashmix(i < n ? ent[i] : 0u, hc);
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
    for (int s = 4; s < n; ++s)
#pragma unroll
        for (int d = 0; d < 4; ++d) pool[d] = ss_mix(pool[d], ss_hashmix(ent[s], hc));
    uint32_t w[8]; hc = 0x8b51f9ddu;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t v = pool[i & 3];
        v ^= hc; hc *= 0x58f38dedu; v *= hc; v ^= v >> 16; w[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
}

// ---------------------------------------------------------------- PCG64
#define OCO_PCG_MULT_HI 0x2360ED051FC65DA4ULL
#define OCO_PCG_MULT_LO 0x4385DF649FCCF645ULL

struct Pcg64 {              // 128-bit LCG state; the increment is kept separately
    uint64_t hi, lo;        // (it is shared by the normal and the label cursor)
};
struct PcgInc { uint64_t hi, lo; };

__device__ __forceinline__ void pcg_step(Pcg64& s, const PcgInc& inc) {
    // state = state * MULT + inc  (mod 2^128); written on unsigned __int128 so that nvcc shares the
    // partial products and keeps one carry chain (18 SASS instructions vs 24 for the 64-bit spelling)
    const unsigned __int128 m = ((unsigned __int128)OCO_PCG_MULT_HI << 64) | OCO_PCG_MULT_LO;
    unsigned __int128 st = ((unsigned __int128)s.hi << 64) | s.lo;
    st = st * m + (((unsigned __int128)inc.hi << 64) | inc.lo);
    s.lo = (uint64_t)st; s.hi = (uint64_t)(st >> 64);
}
__device__ __forceinline__ uint64_t pcg_output(const Pcg64& s) {
    // XSL-RR: rotr64(hi ^ lo, hi >> 58), as two 32-bit funnel shifts (halves swapped when rot >= 32)
    const uint64_t x = s.hi ^ s.lo;
    const unsigned rot = (unsigned)(s.hi >> 58);
    const uint32_t xl = (uint32_t)x, xh = (uint32_t)(x >> 32);
    const uint32_t a = (rot & 32u) ? xh : xl, b = (rot & 32u) ? xl : xh;
    const uint32_t lo = __funnelshift_r(a, b, rot), hi = __funnelshift_r(b, a, rot);   // shift amount taken mod 32
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t pcg_next64(Pcg64& s, const PcgInc& inc) {
    pcg_step(s, inc);
    return pcg_output(s);
}
__device__ inline void pcg_seed(Pcg64& s, PcgInc& inc, const uint64_t w[4]) {
    // initstate = w0:w1, initseq = w2:w3 ; inc = (initseq << 1) | 1
    inc.hi = (w[2] << 1) | (w[3] >> 63);
    inc.lo = (w[3] << 1) | 1ull;
    s.hi = 0; s.lo = 0;
    pcg_step(s, inc);
    uint64_t lo = s.lo + w[1];
    s.hi = s.hi + w[0] + (lo < s.lo ? 1ull : 0ull);
    s.lo = lo;
    pcg_step(s, inc);
}
__device__ inline void pcg_seed3(Pcg64& s, PcgInc& inc, uint64_t a, uint64_t b, uint64_t c) {
    uint64_t w[4]; seedseq3(a, b, c, w); pcg_seed(s, inc, w);
}
// Generator.random(): (raw >> 11) * 2^-53
__device__ __forceinline__ double pcg_double(Pcg64& s, const PcgInc& inc) {
    return __ull2double_rn(pcg_next64(s, inc) >> 11) * (1.0 / 9007199254740992.0);
}

// Label cursor for Generator.choice([-1.0, 1.0], size=T): one 32-bit draw per
// label, low half of each 64-bit raw first, then the high half; label = top bit.
struct LabelCursor {
    Pcg64 s;
    uint64_t cur;
    __device__ __forceinline__ double next(int64_t t, const PcgInc& inc) {
        uint32_t w;
        if ((t & 1) == 0) { cur = pcg_next64(s, inc); w = (uint32_t)cur; }
        else w = (uint32_t)(cur >> 32);
        return (w >> 31) ? 1.0 : -1.0;
    }
};

// ---------------------------------------------------------------- log1p (glibc 2.39, FMA ifunc variant)
// Mirrors sysdeps/ieee754/dbl-64/s_log1p.c as compiled with -mfma in this image's
// libm (the fused operations were read from its disassembly); the ziggurat tail
// (0.028 % of draws) is the only caller.  Bit-exact vs libm for x in (-1, 0].
__device__ inline double log1p_glibc_fma(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
        Lp1 = 6.666666666666735130e-01, Lp2 = 3.999999999940941908e-01, Lp3 = 2.857142874366239149e-01,
        Lp4 = 2.222219843214978396e-01, Lp5 = 1.818357216161805012e-01, Lp6 = 1.531383769920937332e-01,
        Lp7 = 1.479819860511658591e-01;
    double hfsq, f = 0.0, c = 0.0, s, z, R, u;
    int k, hx, hu = 0, ax;
    hx = __double2hiint(x);
    ax = hx & 0x7fffffff;
    k = 1;
    if (hx < 0x3FDA827A) {
        if (ax >= 0x3ff00000) {
            if (x == -1.0) return -__longlong_as_double(0x7ff0000000000000LL);
            return (x - x) / (x - x);
        }
        if (ax < 0x3e200000) {
            if (ax < 0x3c900000) return x;
            return fma(-0.5, x * x, x);
        }
        if (hx > 0 || hx <= (int)0xbfd2bec3) { k = 0; f = x; hu = 1; }
    } else if (hx >= 0x7ff00000) return x + x;
    if (k != 0) {
        if (hx < 0x43400000) {
            u = 1.0 + x;
            hu = __double2hiint(u);
            k = (hu >> 20) - 1023;
            c = (k > 0) ? 1.0 - (u - x) : x - (u - 1.0);
            c /= u;
        } else {
            u = x; hu = __double2hiint(u); k = (hu >> 20) - 1023; c = 0.0;
        }
        hu &= 0x000fffff;
        if (hu < 0x6a09e) {
            u = __hiloint2double(hu | 0x3ff00000, __double2loint(u));
        } else {
            k += 1;
            u = __hiloint2double(hu | 0x3fe00000, __double2loint(u));
            hu = (0x00100000 - hu) >> 2;
        }
        f = u - 1.0;
    }
    hfsq = (0.5 * f) * f;
    if (hu == 0) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            c = fma((double)k, ln2_lo, c);
            return fma((double)k, ln2_hi, c);
        }
        R = hfsq * fma(-0.66666666666666666, f, 1.0);
        if (k == 0) return f - R;
        return fma((double)k, ln2_hi, -((R - fma((double)k, ln2_lo, c)) - f));
    }
    s = f / (2.0 + f);
    z = s * s;
    {
        double R2 = fma(z, Lp3, Lp2), R3 = fma(z, Lp5, Lp4), R4 = fma(z, Lp7, Lp6);
        double z2 = z * z, z4 = z2 * z2, z6 = z4 * z2;
        R = fma(z6, R4, fma(z4, R3, fma(z, Lp1, z2 * R2)));
    }
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return fma((double)k, ln2_hi, -((hfsq - (s * (hfsq + R) + fma((double)k, ln2_lo, c))) - f));
}

// ---------------------------------------------------------------- ziggurat normal
#define OCO_ZIG_R 3.6541528853610087963519472518
#define OCO_ZIG_INV_R 0.27366123732975827203338247596

// Slow path of random_standard_normal: entered with a raw that failed the `rabs < ki` test and
// left only with an accepted normal, i.e. the whole rejection loop (wedge test, tail sampler,
// redraws) lives in this one non-inlined function.  The caller's per-step code therefore
// contains no loop at all -- nvcc 12.9 otherwise merges the rejection loop into the caller's
// uniform-counter time loop and lanes that redraw lose a step (observed on sm_100a).
// State goes in and out by value so that the fast path keeps it in registers.
struct ZigSlowRet { double x; uint64_t hi, lo; };

static __device__ __noinline__ ZigSlowRet zig_slow(uint64_t s_hi, uint64_t s_lo, uint64_t inc_hi, uint64_t inc_lo,
                                                   int idx, uint64_t rabs, double x) {
    Pcg64 s; s.hi = s_hi; s.lo = s_lo;
    PcgInc inc; inc.hi = inc_hi; inc.lo = inc_lo;
    ZigSlowRet out;
    for (;;) {
        if (idx == 0) {
            for (;;) {
                double xx = -OCO_ZIG_INV_R * log1p_glibc_fma(-pcg_double(s, inc));
                double yy = -log1p_glibc_fma(-pcg_double(s, inc));
                if (yy + yy > xx * xx) {
                    out.x = ((rabs >> 8) & 1) ? -(OCO_ZIG_R + xx) : OCO_ZIG_R + xx;
                    out.hi = s.hi; out.lo = s.lo;
                    return out;
                }
            }
        }
        const double f1 = __longlong_as_double((long long)c_zig_fi[idx]);
        const double f0 = __longlong_as_double((long long)c_zig_fi[idx - 1]);
        const double yv = (f0 - f1) * pcg_double(s, inc) + f1;
        const double arg = -0.5 * x * x;
        // Wedge test yv < exp(arg).  Only the comparison matters, so it is first bracketed with a
        // float exponential (relative error < 2e-6): outside a +-1e-5 band the outcome is certain and
        // identical to the double-precision test; inside (probability ~2e-5 per wedge test) the double
        // exp decides.  (CUDA's exp is <= 1 ulp, glibc's is correctly rounded in all but ~1e-3 of
        // cases: the two can disagree on the comparison with probability ~2^-52)
        const double ef = (double)__expf((float)arg);
        bool accept;
        if (yv < ef * (1.0 - 1e-5)) accept = true;
        else if (yv > ef * (1.0 + 1e-5)) accept = false;
        else accept = yv < exp(arg);
        if (accept) break;
        // rejected: a fresh attempt, exactly the top of numpy's for(;;)
        uint64_t r = pcg_next64(s, inc);
        idx = (int)(r & 0xff);
        rabs = (r >> 9) & 0x000fffffffffffffULL;
        x = (__longlong_as_double((long long)(0x4330000000000000ULL | rabs)) - 4503599627370496.0) *
            __longlong_as_double((long long)c_zig_wi[idx]);
        if ((r >> 8) & 1) x = -x;
        if (rabs < c_zig_ki[idx]) break;
    }
    out.x = x; out.hi = s.hi; out.lo = s.lo;
    return out;
}

// One standard normal (Generator.standard_normal), advancing the stream.
__device__ __forceinline__ double zig_normal(const ZigTables& tb, Pcg64& s, const PcgInc& inc) {
    uint64_t r = pcg_next64(s, inc);
    int idx = (int)(r & 0xff);
    uint64_t rabs = (r >> 9) & 0x000fffffffffffffULL;
    // rabs < 2^52: exact int -> double through the 2^52 magic constant (one DADD)
    double x = (__longlong_as_double((long long)(0x4330000000000000ULL | rabs)) - 4503599627370496.0) * tb.wi[idx];
    if ((r >> 8) & 1) x = -x;
    if (rabs < tb.ki[idx]) return x;
    ZigSlowRet o = zig_slow(s.hi, s.lo, inc.hi, inc.lo, idx, rabs, x);
    s.hi = o.hi; s.lo = o.lo;
    return o.x;
}

// Advance the stream past one normal without forming its value on the fast path
// (used to find where the label draws start: choice() comes after all normals).
__device__ __forceinline__ void zig_skip(const ZigTables& tb, Pcg64& s, const PcgInc& inc) {
    uint64_t r = pcg_next64(s, inc);
    int idx = (int)(r & 0xff);
    uint64_t rabs = (r >> 9) & 0x000fffffffffffffULL;
    if (rabs < tb.ki[idx]) return;
    double x = (__longlong_as_double((long long)(0x4330000000000000ULL | rabs)) - 4503599627370496.0) * tb.wi[idx];
    if ((r >> 8) & 1) x = -x;
    ZigSlowRet o = zig_slow(s.hi, s.lo, inc.hi, inc.lo, idx, rabs, x);
    s.hi = o.hi; s.lo = o.lo;
}

}  // namespace oco
