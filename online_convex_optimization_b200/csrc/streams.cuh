// streams.cuh -- on-device sequence families: one CUDA thread produces exactly the (z_t, y_t) rows
// the reference's numpy generators produce for one stream key.
//   GPathStream  fast_algorithms.py:231-239      float64 rows, Rademacher labels via choice()
//   IidStream    sequence_generation.py:54-90    float32 rows, labels sign(z.u), optional Massart flips
// numpy draws every normal of a sequence first and the labels / flip uniforms afterwards from the
// same stream, and the ziggurat consumes a data-dependent number of raw draws, so the label
// cursor is only known once all T*d normals have been generated: the stream protocol below
// (row / begin_labels / label) generates each row exactly once, in numpy's order.
#pragma once
#include "oco_device.cuh"
#include "rng_numpy.cuh"

namespace oco {

// numpy add.reduce pairwise summation (loops_utils.h.src) of squares, compile-time length.
template <int D, typename F>
__device__ __forceinline__ F np_pairwise_sumsq(const F (&z)[D]) {
    if constexpr (D < 8) {
        F r = z[0] * z[0];                 // 0.0 + s0 is exact
#pragma unroll
        for (int j = 1; j < D; ++j) r = r + z[j] * z[j];
        return r;
    } else {
        static_assert(D <= 128, "pairwise block only");
        F r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = z[j] * z[j];
        constexpr int NB = D - (D % 8);
#pragma unroll
        for (int i = 8; i < NB; i += 8)
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = r[j] + z[i + j] * z[i + j];
        F res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
#pragma unroll
        for (int i = NB; i < D; ++i) res = res + z[i] * z[i];
        return res;
    }
}

// Stream protocol (one CUDA thread per stream key):
//   init(...)            seed the generator
//   row(tb, z, ybit)     next row of normals, clipped as the reference does; ybit = provisional label
//   begin_labels()       call once after all T rows: numpy draws labels / flip uniforms AFTER the normals
//   label(t, ybit)       final label of round t (consumes label / flip draws in order t = 0, 1, ...)
// Rows are produced exactly once per sequence; consumers that need them again (the play pass and
// the comparator pass of the sweep kernel) keep them in a scratch buffer instead of regenerating.

template <int D>
struct GPathStream {                      // fast_algorithms.py:231-239
    typedef double elem_t;
    static constexpr bool kProvLabels = false;   // labels do not depend on the rows (row() leaves ybit = 0)
    PcgInc inc;
    Pcg64 s;
    LabelCursor lab;

    // key = SeedSequence([a, T, c])
    __device__ __forceinline__ void init(const ZigTables& tb, uint64_t a, int64_t T, uint64_t c, double /*p_flip*/, int /*family*/) {
        pcg_seed3(s, inc, a, (uint64_t)T, c);
    }
    __device__ __forceinline__ void row(const ZigTables& tb, double (&z)[D], int& ybit) {
#pragma unroll
        for (int j = 0; j < D; ++j) z[j] = zig_normal(tb, s, inc);
        double nrm;
        if constexpr (D == 1) {
            // sqrt(fl(z*z)) == |z| for 1 < |z| < 1e145, and both sides are <= 1 together otherwise, which is
            // all the clip below looks at: no square root (and nothing for the compiler to speculate)
            nrm = fabs(z[0]);
        } else {
            nrm = sqrt_of_sumsq<D>(np_pairwise_sumsq<D, double>(z), z[0]);
        }
        double inv = rcp_inrange(nrm > 1.0 ? nrm : 1.0); // z *= 1.0 / np.maximum(norms, 1.0); 1 <= norm < 1e3
#pragma unroll
        for (int j = 0; j < D; ++j) z[j] = z[j] * inv;
        ybit = 0;
    }
    __device__ __forceinline__ void begin_labels() { lab.s = s; lab.cur = 0; }
    __device__ __forceinline__ double label(int t, int /*ybit*/) { return lab.next(t, inc); }   // choice([-1, 1], T)
};

template <int D>
struct IidStream {                        // sequence_generation.py:54-90
    typedef float elem_t;
    static constexpr bool kProvLabels = true;    // label = sign(z.u) from row(), then possibly flipped by label()
    PcgInc inc;
    Pcg64 s;
    float u[D];
    double p_flip;       // < 0: no flips (separable family)

    __device__ __forceinline__ void init(const ZigTables& tb, uint64_t run_seed, int64_t T, uint64_t stream_id, double p, int family) {
        {   // task vector u: _rng(run_seed, 0, 11 | 21), float32, normalised (sequence_generation.py:55-59)
            Pcg64 us; PcgInc ui;
            pcg_seed3(us, ui, run_seed, 0, family == 2 ? 21ull : 11ull);
#pragma unroll
            for (int j = 0; j < D; ++j) u[j] = (float)zig_normal(tb, us, ui);
            double acc = 0.0;        // OpenBLAS sdot tail: float products, double accumulator
#pragma unroll
            for (int j = 0; j < D; ++j) acc = acc + (double)(u[j] * u[j]);
            float n = sqrtf((float)acc);
            if (n > 0.0f) {
#pragma unroll
                for (int j = 0; j < D; ++j) u[j] = u[j] / n;
            }
        }
        p_flip = (family == 2) ? p : -1.0;
        pcg_seed3(s, inc, run_seed, (uint64_t)T, stream_id);
    }
    __device__ __forceinline__ void row(const ZigTables& tb, float (&z)[D], int& ybit) {
#pragma unroll
        for (int j = 0; j < D; ++j) z[j] = (float)zig_normal(tb, s, inc);
        float nrm = sqrtf(np_pairwise_sumsq<D, float>(z));
        if (!(nrm >= 1.0f)) nrm = 1.0f;                  // np.maximum(norms, 1.0, out=norms)
        float inv = 1.0f / nrm;
#pragma unroll
        for (int j = 0; j < D; ++j) z[j] = z[j] * inv;
        float q = z[0] * u[0];                           // y = sign(z @ u), zeros -> +1
#pragma unroll
        for (int j = 1; j < D; ++j) q = q + z[j] * u[j];
        ybit = (q < 0.0f) ? 0 : 1;
    }
    __device__ __forceinline__ void begin_labels() {}    // the flip uniforms continue the same stream
    __device__ __forceinline__ float label(int /*t*/, int ybit) {
        float y = ybit ? 1.0f : -1.0f;
        if (p_flip >= 0.0) { if (pcg_double(s, inc) < p_flip) y = -y; }   // flips = gen.random(T) < p
        return y;
    }
};

}  // namespace oco
