// kernels_wide_ipm.cu -- the general exact-FTL prefix solver of kernels_exact.cu for 8 < d <= 32.
//
// Same problem, same method and same stopping rule as exact_ftl_barrier_kernel (kernels_exact.cu: path-following
// Newton on the reduced log-barrier function F_mu of
//        min_{x in C} 0.5 * sum_{i<k} |z_i.x - y_i|,     C = unit l2 or linf ball          (exact_ftl.py:83-105),
// epigraph variables eliminated in closed form, step length from the root of the directional derivative, mu
// shrinking by 0.15 per stage down to tol * max(1, f(0)) / (k + nu)); see the header of that file for the maths.
// PARITY UNPINNED (the reference's arithmetic is CVXPY's conic solver), validated against scipy in the tests.
//
// What changes is the mapping.  The small-d kernel keeps the d + d(d+1)/2 sums of a Newton step in every lane's
// registers and solves the d x d system redundantly per lane; at d = 32 that is 560 accumulators.  Here one warp
// still owns one (sequence, prefix) item, but LANE a OWNS VARIABLE a:
//   * the prefix rows are staged 32 at a time in shared memory; lane i forms the residual, the gradient weight c_i
//     and the Hessian weight w_i of row i of the tile (sequential dot with x), then lane a accumulates g[a] and ROW a
//     of the Hessian in registers, reading z[i][b] as shared-memory broadcasts;
//   * the Hessian goes to shared memory (row a written by lane a), Jacobi-scaled to a unit diagonal, and is factored
//     by a left-looking, rank-revealing Cholesky with lane i producing L[i][j] of column j (columns whose pivot falls
//     below 1e-13 are dropped: at d = 32 the barrier of a nearly active bound makes the unscaled matrix span > 1e13);
//     the two triangular solves run column by column with one shuffle per column;
//   * the line search evaluates phi'(a) with lanes striding over the rows (x and the direction in shared memory).
// The l1 ball (2^d facets in the d <= 8 kernel's barrier) is solved here in SPLIT VARIABLES for 8 < d <= 16:
// x = u - w, u, w >= 0, sum(u + w) <= 1 -- 2d <= 32 variables (one per lane), 2d + 1 barrier terms, data rows [z, -z];
// same Newton / line-search machinery, action = u - w.
#include "oco_device.cuh"
#include "oco_internal.h"

namespace oco {

namespace {

constexpr int kIpmD = 32;            // max variables (one per lane)
constexpr int kIpmWarps = 2;         // items per CTA
constexpr int kIpmLd = kIpmD + 1;    // padded leading dimension (conflict-free column access)
constexpr double kTheta = 0.15, kCenter = 0.05;
constexpr int kLineEvals = 12;

struct WideIpmArgs {
    const void* z; const void* y;
    int64_t B, T; int d;
    int norm;                 // 0 = l2, 1 = linf, 2 = l1 (split variables, d <= 16)
    int max_iter; double tol;
    int64_t k0;
    const int* only_flagged;  // optional [B]
    double* actions;          // [B][T+1][d]
    int* status;              // [B]
};

__device__ __forceinline__ double isum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ double imin(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ double imax(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

struct IpmSmem {
    double tile[32][kIpmLd];     // 32 staged rows (columns >= d are zero)
    double H[kIpmD][kIpmLd];     // Hessian / Cholesky factor (lower triangle)
    double x[kIpmD], dir[kIpmD], xa[kIpmD];
    double xe[kIpmD], de[kIpmD];  // point / direction as seen by the data rows (l1: u - w; otherwise copies)
    double c[32], w[32];
};

template <typename TIN>
__global__ void __launch_bounds__(kIpmWarps * 32) exact_ftl_wide_barrier_kernel(const WideIpmArgs p) {
    __shared__ IpmSmem sm_all[kIpmWarps];
    const int lane = threadIdx.x & 31;
    IpmSmem& sm = sm_all[threadIdx.x >> 5];
    const int64_t K = p.T - p.k0 + 1;
    const int64_t item = (int64_t)blockIdx.x * kIpmWarps + (threadIdx.x >> 5);
    if (item >= p.B * K) return;
    const int64_t b = item / K;
    const int64_t k = p.T - (item - b * K);                     // longest prefixes first
    if (p.only_flagged && p.only_flagged[b] == 0) return;
    const int64_t T = p.T;
    const int d = p.d;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * T * d;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * T;
    double* __restrict__ act = p.actions + b * (T + 1) * d;
    if (k == p.k0)
        for (int64_t j = lane; j < p.k0 * d; j += 32) act[j] = 0.0;

    double sum_abs_y = 0.0;
    for (int64_t i = lane; i < k; i += 32) sum_abs_y += fabs((double)yb[i]);
    sum_abs_y = isum(sum_abs_y);
    const bool split = p.norm == 2;                             // l1 ball: variables (u, w), x = u - w
    const int nv = split ? 2 * d : d;                           // variables (one per lane)
    const double nu = p.norm == 0 ? 1.0 : (split ? 2.0 * d + 1.0 : 2.0 * d);
    const double mu_f = p.tol * fmax(1.0, 0.5 * sum_abs_y) / ((double)k + nu);
    double mu = fmax(fmax(sum_abs_y / (double)k, mu_f), 1e-3);

    const bool mine = lane < nv;                                // this lane owns a variable
    double xv = (split && mine) ? 1.0 / (4.0 * d) : 0.0;        // x[lane]; l1: strictly feasible start, u - w = 0
    sm.x[lane] = xv; sm.dir[lane] = 0.0; sm.xa[lane] = 0.0; sm.xe[lane] = 0.0; sm.de[lane] = 0.0;
#pragma unroll 1
    for (int i = 0; i < 32; ++i) sm.tile[i][lane] = 0.0;
    __syncwarp();

    // mu * grad B(xq) . dir for the ball's barrier (all lanes return the same value)
    auto barrier_dir = [&](const double* xq, const double* dq) -> double {
        const double xl = mine ? xq[lane] : 0.0, dl = mine ? dq[lane] : 0.0;
        if (p.norm == 0) {
            const double n2 = isum(xl * xl), xp = isum(xl * dl);
            return 2.0 * mu * xp / (1.0 - n2);
        }
        if (split) {                                            // -sum log v_a - log(1 - sum v)
            const double sv = isum(xl), sd = isum(dl), a1 = isum(mine ? -dl / xl : 0.0);
            return mu * (a1 + sd / (1.0 - sv));
        }
        return isum(mine ? 2.0 * mu * xl * dl / (1.0 - xl * xl) : 0.0);
    };
    auto strictly_inside = [&](const double* xq) -> bool {
        const double xl = mine ? xq[lane] : 0.0;
        if (split) return imin(mine ? xl : 1.0) > 1e-300 && isum(xl) < 1.0 - 1e-15;
        const double m = p.norm == 0 ? isum(xl * xl) : imax(fabs(xl));
        return m < 1.0 - 1e-15;
    };

    int newton = 0;
    bool done = false;
    while (newton < p.max_iter) {
        // ---- gradient g[lane] and row `lane` of the Hessian of F_mu at x
        double g = 0.0, Hrow[kIpmD];
#pragma unroll
        for (int j = 0; j < kIpmD; ++j) Hrow[j] = 0.0;
        const double mu2 = mu * mu;
        for (int64_t i0 = 0; i0 < k; i0 += 32) {
            const int nrow = (int)((k - i0) < 32 ? (k - i0) : 32);
            __syncwarp();
            for (int e = lane; e < nrow * d; e += 32) {          // the tile's rows are one contiguous run: coalesced
                const int ri = e / d, cj = e - ri * d;
                const double zv = (double)zb[i0 * d + e];
                sm.tile[ri][cj] = zv;
                if (split) sm.tile[ri][d + cj] = -zv;           // data row of the split variables: [z, -z]
            }
            __syncwarp();
            double cw = 0.0, ww = 0.0;
            if (lane < nrow) {
                double r = -(double)yb[i0 + lane];
                for (int j = 0; j < nv; ++j) r += sm.tile[lane][j] * sm.x[j];
                const double s = sqrt(mu2 + r * r), t = mu + s;
                cw = r / t; ww = mu / (t * s);
            }
            sm.c[lane] = cw; sm.w[lane] = ww;
            __syncwarp();
            if (mine) {
                for (int i = 0; i < nrow; ++i) {
                    const double za = sm.tile[i][lane];
                    g += sm.c[i] * za;
                    const double wz = sm.w[i] * za;
#pragma unroll
                    for (int j = 0; j < kIpmD; ++j) Hrow[j] += wz * sm.tile[i][j];      // columns >= d hold zeros
                }
            }
        }
        // ---- barrier of the ball
        if (p.norm == 0) {
            const double n2 = isum(xv * xv);
            const double q = 1.0 - n2, a = 2.0 * mu / q, bq = 4.0 * mu / (q * q);
            if (mine) {
                g += a * xv;
#pragma unroll
                for (int j = 0; j < kIpmD; ++j) Hrow[j] += bq * xv * sm.x[j] + (j == lane ? a : 0.0);
            }
        } else if (split) {
            const double sl = 1.0 - isum(mine ? xv : 0.0), c1 = mu / sl, c2 = mu / (sl * sl);
            if (mine) {
                g += c1 - mu / xv;
#pragma unroll
                for (int j = 0; j < kIpmD; ++j) Hrow[j] += (j < nv ? c2 : 0.0) + (j == lane ? mu / (xv * xv) : 0.0);
            }
        } else if (mine) {
            const double q = 1.0 - xv * xv;
            g += 2.0 * mu * xv / q;
#pragma unroll
            for (int j = 0; j < kIpmD; ++j) if (j == lane) Hrow[j] += 2.0 * mu * (1.0 + xv * xv) / (q * q);
        }
        // ---- H -> shared memory, Jacobi-scaled to a unit diagonal (lanes that own no variable carry the identity), then
        // a RANK-REVEALING Cholesky: a pivot below 1e-13 (relative to the unit diagonal) marks a direction without
        // curvature at working precision -- flat directions of prefixes shorter than d, or coordinates whose barrier
        // term (1/q^2 near an active bound) dwarfs everything else at small mu -- and that column is dropped (no
        // step along it) instead of dividing rounding noise by a floored pivot.  Solve (S H S) u = -S g, dir = S u.
        double diag = 1.0;
#pragma unroll
        for (int j = 0; j < kIpmD; ++j) if (j == lane && mine) diag = Hrow[j];
        const double sc = 1.0 / sqrt(fmax(diag, 1e-300));
        __syncwarp();
        sm.c[lane] = sc;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kIpmD; ++j)
            sm.H[lane][j] = mine ? (j < nv ? Hrow[j] * sc * sm.c[j] : 0.0) : (j == lane ? 1.0 : 0.0);
        __syncwarp();
        for (int j = 0; j < kIpmD; ++j) {                       // left-looking: lane i >= j forms column j
            double s = 0.0;
            if (lane >= j) {
                s = sm.H[lane][j];
                for (int c = 0; c < j; ++c) s -= sm.H[lane][c] * sm.H[j][c];
            }
            const double piv = __shfl_sync(0xffffffffu, s, j);
            const bool drop = !(piv >= 1e-13);
            const double l = drop ? 1e150 : sqrt(piv), inv = drop ? 0.0 : 1.0 / l;
            if (lane == j) sm.H[j][j] = l;
            else if (lane > j) sm.H[lane][j] = s * inv;
            __syncwarp();
        }
        double rhs = mine ? -g * sc : 0.0;                       // forward: L yv = rhs
        for (int c = 0; c < kIpmD; ++c) {
            const double yc = __shfl_sync(0xffffffffu, rhs, c) / sm.H[c][c];
            if (lane == c) rhs = yc;
            else if (lane > c) rhs -= sm.H[lane][c] * yc;
        }
        for (int c = kIpmD - 1; c >= 0; --c) {                    // backward: L' u = yv
            const double xc = __shfl_sync(0xffffffffu, rhs, c) / sm.H[c][c];
            if (lane == c) rhs = xc;
            else if (lane < c) rhs -= sm.H[c][lane] * xc;
        }
        rhs *= sc;
        const double dv = mine ? rhs : 0.0;                      // dir[lane]
        const double gp = isum(mine ? g * dv : 0.0);
        if (!(-gp > kCenter * mu)) {                            // centred at this mu (also catches NaN)
            if (mu <= mu_f) { done = (gp == gp); break; }
            mu = fmax(mu * kTheta, mu_f);
            continue;
        }
        ++newton;
        sm.dir[lane] = dv;
        __syncwarp();
        sm.de[lane] = split ? (lane < d ? dv - sm.dir[lane + d] : 0.0) : dv;
        __syncwarp();
        // ---- step length: largest a <= 1 that stays strictly inside, then the root of phi'(a)
        double a = 1.0;
        for (int h = 0; h < 60; ++h) {
            __syncwarp();
            sm.xa[lane] = xv + a * dv;
            __syncwarp();
            if (strictly_inside(sm.xa)) break;
            a *= 0.5;
        }
        double lo = 0.0, hi = -1.0, dlo = gp, dhi = 0.0;
        for (int e = 0; e < kLineEvals; ++e) {
            __syncwarp();
            sm.xa[lane] = xv + a * dv;
            __syncwarp();
            sm.xe[lane] = split ? (lane < d ? sm.xa[lane] - sm.xa[lane + d] : 0.0) : sm.xa[lane];
            __syncwarp();
            double da = 0.0;
            for (int64_t i = lane; i < k; i += 32) {
                const TIN* __restrict__ zi = zb + i * d;
                double r = -(double)yb[i], zp = 0.0;
                for (int j = 0; j < d; ++j) { const double zj = (double)zi[j]; r += zj * sm.xe[j]; zp += zj * sm.de[j]; }
                da += r / (mu + sqrt(mu2 + r * r)) * zp;
            }
            da = isum(da) + barrier_dir(sm.xa, sm.dir);
            if (fabs(da) <= 0.2 * fabs(gp) || (da <= 0.0 && hi < 0.0)) break;
            if (da > 0.0) { hi = a; dhi = da; } else { lo = a; dlo = da; }
            const double sec = lo - dlo * (hi - lo) / (dhi - dlo), w = 0.1 * (hi - lo);
            a = fmin(fmax(sec, lo + w), hi - w);
        }
        xv += a * dv;
        __syncwarp();
        sm.x[lane] = xv;
        __syncwarp();
    }
    if (!done && lane == 0) atomicAdd(p.status + b, 1);
    const double wv = __shfl_sync(0xffffffffu, xv, (lane + d) & 31);     // l1: the w half of the split variables
    if (lane < d) act[k * d + lane] = split ? xv - wv : xv;
}

}  // namespace

// status must hold 0 for every sequence this launch solves (the kernel adds one per unconverged prefix).
cudaError_t launch_exact_ftl_wide_general(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d, int norm,
                                          int max_iter, double tol, int64_t first_prefix, const int* only_flagged,
                                          double* actions, int* status, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (d < 1 || d > kIpmD || norm < 0 || norm > 2 || (norm == 2 && 2 * d > kIpmD)) return cudaErrorInvalidValue;
    if (T == 0) return cudaMemsetAsync(actions, 0, sizeof(double) * (size_t)B * d, st);
    WideIpmArgs a;
    a.z = z; a.y = y; a.B = B; a.T = T; a.d = d; a.norm = norm; a.max_iter = max_iter; a.tol = tol;
    a.k0 = first_prefix < 1 ? 1 : (first_prefix > T ? T : first_prefix);
    a.only_flagged = only_flagged; a.actions = actions; a.status = status;
    const int64_t items = B * (T - a.k0 + 1);
    if (items <= 0) return cudaSuccess;
    const int64_t blocks = (items + kIpmWarps - 1) / kIpmWarps;
    if (blocks > 0x7fffffff) return cudaErrorInvalidValue;
    if (in_f32) exact_ftl_wide_barrier_kernel<float><<<(unsigned)blocks, kIpmWarps * 32, 0, st>>>(a);
    else exact_ftl_wide_barrier_kernel<double><<<(unsigned)blocks, kIpmWarps * 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace oco
