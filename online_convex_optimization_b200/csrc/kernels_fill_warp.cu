// kernels_fill_warp.cu -- warp-cooperative generator for the g-path family (fast_algorithms.py:231-239):
// ONE WARP per sequence writes z[b][T][d] (rows clipped to the unit ball) and y[b][T] in the reference layout, any
// d <= 128, bit-identical to numpy's draws.
//
// fill_kernel (kernels_seeded.cu) gives every sequence one THREAD: a sequence of BASELINE config 4 is 6.4e5 normals
// long, a resident chunk holds only 14,208 sequences (74 GB of rows), so the GPU ran 444 warps on 592 schedulers and
// each lane wrote rows T*d*8 bytes apart from its neighbours.  Here the 32 lanes of a warp work on one stream:
//   * PCG64 is an LCG: state_{n+32} = A32 * state_n + C32 (mod 2^128), so lane l keeps the state of raw number
//     32 r + l and every round the warp draws 32 CONSECUTIVE raws with one 128-bit multiply-add per lane;
//   * the raw stream is fixed, only the ROLE of a raw (start of a normal / wedge uniform / tail uniform) depends on
//     the accept tests before it.  Every lane evaluates its raw as "start of a normal" (table lookup + the
//     rabs < ki test); a ballot gives the 32 fast-accept bits, and a warp-uniform walk over the (rare: 1.5 % of
//     raws) failing positions assigns the roles, all lanes evaluating the same wedge / tail test on broadcast
//     values -- no divergence, no serial pass over the 98.5 %;
//   * accepted normals are compacted (prefix popcount) into a shared-memory row buffer; as soon as a row is
//     complete its squared norm is summed in numpy's order (sequential for d < 8, the 8-accumulator pairwise block
//     of add.reduce for 8 <= d <= 128), and the row is scaled and stored with coalesced 8-byte (or 4-byte) writes;
//   * the last < 34 normals of a sequence are drawn by lane 0 alone (so that the walk never runs past the last
//     normal into the label draws); choice([-1, 1], T) then continues the stream 32 raws = 64 labels per round.
// Instruction cost per raw is the same as the thread-per-sequence generator (~2.7 warp-instructions); what changes is
// that a chunk of 14,208 sequences now runs 14,208 warps, and every store is coalesced.
#include "gsweep_impl.cuh"

namespace oco {
namespace {

typedef unsigned __int128 u128;
constexpr int kFillWarps = 4;                 // warps (sequences) per CTA
constexpr int kStageCap = 512;                // staged normals per warp: up to 4 rows + one round (power of two)

__device__ __forceinline__ u128 mk128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }
__device__ __forceinline__ u128 pcg_mult() { return mk128(OCO_PCG_MULT_HI, OCO_PCG_MULT_LO); }

struct WarpGenArgs {
    uint64_t key_a, key_c0; const uint64_t* keys;
    int64_t B; int T; int d; int out_f32;
    void* z; void* y;
};

// wedge acceptance for candidate (x, idx) with the uniform drawn from raw r -- the same decision procedure as
// zig_slow_la_a (float bracket, exact exp only inside the +-1e-5 band)
__device__ __forceinline__ bool wedge_accept(const ZigTab2& tb, double x, int idx, uint64_t r) {
    const double u = __ull2double_rn(r >> 11) * (1.0 / 9007199254740992.0);
    const double f1 = tb.fi[idx], f0 = tb.fi[idx - 1];
    const double yv = (f0 - f1) * u + f1;
    const double arg = -0.5 * x * x;
    const double ef = (double)__expf((float)arg);
    if (yv < ef * (1.0 - 1e-5)) return true;
    if (yv > ef * (1.0 + 1e-5)) return false;
    return zig_wedge_exact(yv, arg);
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    return ((uint64_t)__shfl_sync(0xffffffffu, (unsigned)(v >> 32), src) << 32) | __shfl_sync(0xffffffffu, (unsigned)v, src);
}
__device__ __forceinline__ double shfld(double v, int src) { return __longlong_as_double((long long)shfl64((uint64_t)__double_as_longlong(v), src)); }

// carry state of the role walk across rounds (identical in all lanes)
enum { MODE_NONE = 0, MODE_WEDGE = 1, MODE_TAIL1 = 2, MODE_TAIL2 = 3 };

template <typename TOUT, int DC>      // DC > 0: d is the compile-time constant DC (loops unroll); DC == 0: any d <= 128
__global__ void __launch_bounds__(32 * kFillWarps) fill_warp_kernel(const WarpGenArgs p) {
    __shared__ ZigTab2 tb;
    __shared__ double stage_all[kFillWarps][kStageCap];
    zig_tab2_load(tb);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * kFillWarps + wib;
    if (b >= p.B) return;
    double* stage = stage_all[wib];
    const int T = p.T, d = DC > 0 ? DC : p.d;
    const int64_t N = (int64_t)T * d;                      // normals of the sequence
    TOUT* __restrict__ zout = (TOUT*)p.z + b * N;
    TOUT* __restrict__ yout = (TOUT*)p.y + b * (int64_t)T;

    // ---- seed, per-lane jump constants: A_k = M^k, G_k = 1 + M + ... + M^(k-1)  (state_{n+k} = A_k state_n + inc G_k)
    const uint64_t ka = p.keys ? p.keys[2 * b] : p.key_a;
    const uint64_t kc = p.keys ? p.keys[2 * b + 1] : p.key_c0 + (uint64_t)b;
    Pcg64 s0; PcgInc inc0;
    pcg_seed3(s0, inc0, ka, (uint64_t)T, kc);
    const u128 inc = mk128(inc0.hi, inc0.lo);
    u128 Ak = 1, Gk = 0, Alane = 1, Glane = 0, Alane1 = 1, Glane1 = 0;
    for (int k = 0; k <= 32; ++k) {                        // after the body: Ak = M^k, Gk = G_k
        if (k == lane) { Alane = Ak; Glane = Gk; }          // jump by `lane` (labels: lane l draws raw number l of a round)
        if (k == lane + 1) { Alane1 = Ak; Glane1 = Gk; }    // jump by lane + 1 (state AFTER drawing raw number l)
        if (k < 32) { Gk = Gk * pcg_mult() + 1; Ak = Ak * pcg_mult(); }
    }
    const u128 A32 = Ak, C32 = inc * Gk;                    // one round ahead
    u128 base = mk128(s0.hi, s0.lo);                        // stream state before the next undrawn raw of the round
    u128 st = Alane1 * base + inc * Glane1;                 // this lane's state after drawing its raw of round 0
    u128 st_prev = 0; bool any_round = false;
    (void)Alane; (void)Glane;

    int mode = MODE_NONE, c_idx = 0;                        // carry of the role walk
    double c_x = 0.0, c_xx = 0.0; uint64_t c_rabs = 0;
    int64_t done = 0;                                       // normals emitted so far
    int cnt = 0, head = 0;                                  // staged normals: stage[(head + i) & (kStageCap - 1)], i < cnt
    int t_out = 0;                                          // rows written

    // ---- emit rows from the stage, up to 4 at a time: the 8-lane group g = lane / 8 sums row g's squares in numpy's
    // order (np.linalg.norm -> sqrt(add.reduce(z*z)): sequential for d < 8, else 8 running sums over i = j, j + 8, ...
    // combined as ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)), then the d % 8 leftovers one by one)
    const int batch = max(1, min(4, (kStageCap - 34) / d));            // rows per flush that fit next to one more round
    auto flush_rows = [&](bool all) {
        const int grp = lane >> 3, j8 = lane & 7;
        while (cnt >= (all ? d : batch * d)) {
            const int nrows = min(cnt / d, 4);
            const int rb = head + grp * d;                               // this group's row (valid if grp < nrows)
            double n2 = 0.0;
            if (grp < nrows) {
                if (d < 8) {
                    for (int j = 0; j < d; ++j) { const double v = stage[(rb + j) & (kStageCap - 1)]; n2 = (j == 0) ? v * v : n2 + v * v; }
                } else {
                    const int nb = d - (d & 7);
                    const double v0 = stage[(rb + j8) & (kStageCap - 1)];
                    double r = v0 * v0;
#pragma unroll
                    for (int i = 8 + j8; i < nb; i += 8) { const double v = stage[(rb + i) & (kStageCap - 1)]; r = r + v * v; }
                    n2 = r;
                }
            }
            if (d >= 8) {
                n2 = n2 + __shfl_xor_sync(0xffffffffu, n2, 1);           // (r0 + r1), (r2 + r3), ...
                n2 = n2 + __shfl_xor_sync(0xffffffffu, n2, 2);           // ((r0 + r1) + (r2 + r3)), ...
                n2 = n2 + __shfl_xor_sync(0xffffffffu, n2, 4);
                if (grp < nrows) {
                    for (int i = d - (d & 7); i < d; ++i) { const double v = stage[(rb + i) & (kStageCap - 1)]; n2 = n2 + v * v; }
                }
            }
            const double nrm = sqrt(n2);
            const double inv_mine = 1.0 / (nrm > 1.0 ? nrm : 1.0);     // z *= 1.0 / np.maximum(norms, 1.0)
            for (int g = 0; g < nrows; ++g) {
                const double inv = shfld(inv_mine, 8 * g);
                TOUT* __restrict__ zr = zout + (int64_t)(t_out + g) * d;
                const int rg = head + g * d;
#pragma unroll
                for (int j = lane; j < d; j += 32) zr[j] = (TOUT)(stage[(rg + j) & (kStageCap - 1)] * inv);
            }
            head = (head + nrows * d) & (kStageCap - 1); cnt -= nrows * d; t_out += nrows;
        }
        __syncwarp();
    };
    auto push = [&](double v) {                              // uniform: every lane calls it with the same v; lane 0 stores
        if (lane == 0) stage[(head + cnt) & (kStageCap - 1)] = v;
        ++cnt; ++done;
    };

    // ---- parallel rounds while a whole round cannot overshoot the last normal
    while (N - done > 34) {
        any_round = true;
        const uint64_t r = pcg_output(Pcg64{(uint64_t)(st >> 64), (uint64_t)st});
        double x; int idx; uint64_t mag;
        const bool fast = zig_try(tb, r, x, idx, mag);
        const unsigned F = __ballot_sync(0xffffffffu, fast);
        unsigned W = 0;                                      // positions whose (possibly replaced) value x is an accepted normal
        int pos = 0;
        int carry_emit = 0; double carry_val = 0.0;          // a normal decided here whose candidate sat in an earlier round
        int cand_pos = -1;                                   // position of the pending candidate inside THIS round (-1: earlier round)
        if (F == 0xffffffffu && mode == MODE_NONE) W = 0xffffffffu;      // 62 % of the rounds: 32 fast normals, no walk
        else for (;;) {
            if (mode == MODE_NONE) {
                const unsigned rest = pos >= 32 ? 0u : (~F & (0xffffffffu << pos));
                const int f = rest ? __ffs(rest) - 1 : 32;
                if (f > pos) W |= (f >= 32 ? 0xffffffffu : ((1u << f) - 1u)) & (0xffffffffu << pos);   // fast normals [pos, f)
                if (f >= 32) break;
                c_idx = __shfl_sync(0xffffffffu, idx, f);
                c_x = shfld(x, f);
                c_rabs = shfl64(mag, f) & 0x000fffffffffffffULL;
                mode = (c_idx == 0) ? MODE_TAIL1 : MODE_WEDGE;
                cand_pos = f; pos = f + 1;
            } else {
                if (pos >= 32) break;                        // the next uniform lives in the next round: carry the mode
                const uint64_t ru = shfl64(r, pos);
                ++pos;
                bool emit = false; double v = c_x;
                if (mode == MODE_WEDGE) {
                    emit = wedge_accept(tb, c_x, c_idx, ru); // rejected: nothing is emitted, the next raw is a fresh attempt
                    mode = MODE_NONE;
                } else if (mode == MODE_TAIL1) {
                    const double u = __ull2double_rn(ru >> 11) * (1.0 / 9007199254740992.0);
                    c_xx = -OCO_ZIG_INV_R * log1p_glibc_fma(-u);
                    mode = MODE_TAIL2;
                } else {
                    const double u = __ull2double_rn(ru >> 11) * (1.0 / 9007199254740992.0);
                    const double yy = -log1p_glibc_fma(-u);
                    if (yy + yy > c_xx * c_xx) {
                        emit = true; v = ((c_rabs >> 8) & 1) ? -(OCO_ZIG_R + c_xx) : OCO_ZIG_R + c_xx;
                        mode = MODE_NONE;
                    } else mode = MODE_TAIL1;
                }
                if (emit) {
                    if (cand_pos >= 0) { if (lane == cand_pos) x = v; W |= 1u << cand_pos; }   // emitted at its candidate's position
                    else { carry_emit = 1; carry_val = v; }
                }
                if (mode == MODE_NONE) cand_pos = -1;
            }
        }
        // ---- compact: [carried normal] + the positions in W, in stream order
        const int n_emit = __popc(W) + carry_emit;
        if (carry_emit && lane == 0) stage[(head + cnt) & (kStageCap - 1)] = carry_val;
        if ((W >> lane) & 1u) stage[(head + cnt + carry_emit + __popc(W & ((1u << lane) - 1u))) & (kStageCap - 1)] = x;
        cnt += n_emit; done += n_emit;
        st_prev = st;
        st = st * A32 + C32;                                 // this lane's raw of the next round
        __syncwarp();
        flush_rows(false);
    }

    // ---- the last normals: lane 0 alone, from the exact stream position, honouring the carried mode
    {
        // state before the first undrawn raw = lane 31's state after its draw of the last completed round
        uint64_t s_hi = (uint64_t)(base >> 64), s_lo = (uint64_t)base;
        if (any_round) { s_hi = shfl64((uint64_t)(st_prev >> 64), 31); s_lo = shfl64((uint64_t)st_prev, 31); }
        int64_t rem = N - done;
        double vals[36];
        int nv = 0;
        if (lane == 0) {
            Pcg64 s; s.hi = s_hi; s.lo = s_lo;
            PcgInc ic; ic.hi = inc0.hi; ic.lo = inc0.lo;
            if (mode != MODE_NONE && rem > 0) {              // finish the pending candidate first
                double x = c_x; int idx = c_idx; uint64_t rabs = c_rabs;
                bool got = false; double v = 0.0;
                if (mode == MODE_WEDGE) {
                    if (wedge_accept(tb, x, idx, pcg_next64(s, ic))) { got = true; v = x; }
                } else {
                    double xx = c_xx;
                    int m = mode;
                    for (;;) {
                        if (m == MODE_TAIL1) { xx = -OCO_ZIG_INV_R * log1p_glibc_fma(-pcg_double(s, ic)); m = MODE_TAIL2; }
                        const double yy = -log1p_glibc_fma(-pcg_double(s, ic));
                        if (yy + yy > xx * xx) { got = true; v = ((rabs >> 8) & 1) ? -(OCO_ZIG_R + xx) : OCO_ZIG_R + xx; break; }
                        m = MODE_TAIL1;
                    }
                }
                if (got) { vals[nv++] = v; }
                else {                                       // rejected wedge: a fresh attempt for the same normal
                }
            }
            while (nv < rem) vals[nv++] = zig_normal2(tb, s, ic);
            s_hi = s.hi; s_lo = s.lo;
        }
        s_hi = shfl64(s_hi, 0); s_lo = shfl64(s_lo, 0);
        for (int i = 0; i < (int)rem; ++i) {
            const double v = shfld(lane == 0 ? vals[i < 36 ? i : 35] : 0.0, 0);
            push(v);
            if (cnt >= batch * d) { __syncwarp(); flush_rows(false); }
        }
        __syncwarp();
        flush_rows(true);
        base = mk128(s_hi, s_lo);                            // choice([-1, 1], T) continues the stream here
    }

    // ---- labels: one 32-bit draw per label, low half of each raw first, then the high half; label = top bit
    {
        u128 sl = Alane1 * base + inc * Glane1;              // state after drawing raw number `lane` of the label stream
        const int n_raw = (T + 1) / 2;
        for (int r0 = 0; r0 < n_raw; r0 += 32) {
            const uint64_t r = pcg_output(Pcg64{(uint64_t)(sl >> 64), (uint64_t)sl});
            const int t0 = 2 * (r0 + lane);
            if (t0 < T) yout[t0] = (TOUT)(((uint32_t)r >> 31) ? 1.0 : -1.0);
            if (t0 + 1 < T) yout[t0 + 1] = (TOUT)((r >> 63) ? 1.0 : -1.0);
            sl = sl * A32 + C32;
        }
    }
}

}  // namespace

bool fill_warp_eligible(int family, int d) { return family == FAM_GPATH && d >= 1 && d <= 128; }

cudaError_t launch_fill_warp(uint64_t key_a, uint64_t key_c0, const uint64_t* keys, int64_t B, int64_t T, int d, int out_f32,
                             void* z, void* y, cudaStream_t st) {
    if (B <= 0 || T <= 0) return cudaSuccess;
    WarpGenArgs a;
    a.key_a = key_a; a.key_c0 = key_c0; a.keys = keys; a.B = B; a.T = (int)T; a.d = d; a.out_f32 = out_f32; a.z = z; a.y = y;
    const unsigned gx = (unsigned)((B + kFillWarps - 1) / kFillWarps);
#define OCO_FILL_LAUNCH(DC) do { if (out_f32) fill_warp_kernel<float, DC><<<gx, 32 * kFillWarps, 0, st>>>(a); \
                                 else fill_warp_kernel<double, DC><<<gx, 32 * kFillWarps, 0, st>>>(a); } while (0)
    switch (d) {
    case 1: OCO_FILL_LAUNCH(1); break;
    case 5: OCO_FILL_LAUNCH(5); break;
    case 8: OCO_FILL_LAUNCH(8); break;
    case 16: OCO_FILL_LAUNCH(16); break;
    case 32: OCO_FILL_LAUNCH(32); break;
    case 64: OCO_FILL_LAUNCH(64); break;
    default: OCO_FILL_LAUNCH(0); break;
    }
#undef OCO_FILL_LAUNCH
    return cudaGetLastError();
}

}  // namespace oco
