// kernels_exact.cu -- general exact-FTL prefix solver (outside the linear regime, and for the l1 / linf
// balls): for every prefix k of every sequence solve
//        min_{x in C}  f(x) = 0.5 * sum_{i<k} |z_i.x - y_i|,      C = {||x||_p <= 1},  p in {2, inf, 1}
// (the problem exact_ftl.py:83-105 hands to CVXPY after every appended row, :140-149, :280-303).
//
// The reference's arithmetic for this path lives in a third-party interior-point solver that is not
// installed here and is not pinned by the reference: PARITY UNPINNED.  This kernel is therefore not a
// restatement but an independent interior-point solver with a stated tolerance.
//
// Method (path-following on a reduced log barrier).  Write the problem with epigraph variables,
// min sum t_i  s.t. |r_i(x)| <= t_i, x in C, and attach the log barrier with parameter mu.  For fixed x
// the optimal t_i is closed form, t_i = mu + sqrt(mu^2 + r_i^2), which leaves a smooth convex function
// of the d <= 8 variables only:
//        F_mu(x) = sum_i [t_i - mu*log(2*mu*t_i)] + mu*B_C(x)
//        grad    = sum_i (r_i/t_i) z_i                 + mu*grad B_C
//        Hess    = sum_i mu/(t_i*s_i) z_i z_i'          + mu*Hess B_C,        s_i = sqrt(mu^2 + r_i^2)
// with B_C = -log(1-||x||^2) (l2), -sum_j log(1-x_j^2) (linf), -sum_sigma log(1-sigma.x) over the 2^d
// facets (l1).  Newton steps p = -H^-1 g (modified Cholesky: pivots floored at 1e-13*max diag, which
// only matters on flat directions of degenerate prefixes, k < d); the step length is the root of the
// directional derivative phi'(a) = grad F(x+a*p).p found by a safeguarded secant search (no function
// values: at small mu they drown in rounding, derivatives do not).  A stage ends when the Newton
// decrement lambda^2 = -g.p/mu <= 0.05, then mu <- 0.15*mu down to mu_f = tol*max(1, f(0))/(k + nu),
// nu = 1 / 2d / 2^d barrier terms, so the duality gap on exit is <= tol*max(1, f(0)).  Typically
// 20-30 Newton steps per prefix, independent of conditioning (tools/ipm_numpy_prototype.py is the
// host prototype this was tuned on).
//
// Mapping: one warp per (sequence, prefix) work item, every prefix cold-started at x = 0 -- prefixes
// are independent, so B*T warps are in flight instead of B, and a single requested prefix
// (`first_prefix`) gives the same bits as the batched call.  Lanes own rows i = lane, lane+32, ...;
// the d + d(d+1)/2 sums are xor-butterflies (identical in every lane) and every lane solves the
// d x d system redundantly.  Validated against scipy (HiGHS / SLSQP) in the tests.
#include "oco_device.cuh"
#include "oco_internal.h"

namespace oco {

namespace {

constexpr int kExWarps = 4;
constexpr double kTheta = 0.15;      // mu reduction per stage
constexpr double kCenter = 0.05;     // Newton decrement^2 that ends a stage
constexpr int kLineEvals = 12;

struct ExactArgs {
    const void* z; const void* y;
    int64_t B, T;
    int norm;                 // 0 = l2, 1 = linf, 2 = l1
    int max_iter; double tol;
    int64_t k0;               // shortest prefix that is solved (>= 1); shorter ones get zero actions
    const int* only_flagged;  // optional [B]
    double* actions;          // [B][T+1][D]
    int* status;              // [B]: += 1 for every prefix that used up max_iter Newton steps
};

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// mu * (gradient, Hessian lower triangle) of the ball's barrier at x (strictly inside), this lane's
// share: the caller adds it BEFORE the warp reduction.  l2 / linf are closed form (lane 0 carries
// them); the 2^D facets of the l1 ball are spread over the lanes.
template <int D>
__device__ __forceinline__ void add_barrier(const double (&x)[D], int norm, double mu, int lane, double (&g)[D], double (&H)[D][D]) {
    if (norm == 0) {
        if (lane != 0) return;
        double n2 = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) n2 += x[j] * x[j];
        const double q = 1.0 - n2, a = 2.0 * mu / q, b = 4.0 * mu / (q * q);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            g[i] += a * x[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) H[i][j] += b * x[i] * x[j] + (i == j ? a : 0.0);
        }
    } else if (norm == 1) {
        if (lane != 0) return;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const double q = 1.0 - x[i] * x[i];
            g[i] += 2.0 * mu * x[i] / q;
            H[i][i] += 2.0 * mu * (1.0 + x[i] * x[i]) / (q * q);
        }
    } else {
        for (int f = lane; f < (1 << D); f += 32) {
            double s = 1.0;
#pragma unroll
            for (int j = 0; j < D; ++j) s -= ((f >> j) & 1) ? x[j] : -x[j];
            const double a = mu / s, b = a / s;
#pragma unroll
            for (int i = 0; i < D; ++i) {
                const bool pi = (f >> i) & 1;
                g[i] += pi ? a : -a;
#pragma unroll
                for (int j = 0; j <= i; ++j) H[i][j] += (pi == (bool)((f >> j) & 1)) ? b : -b;
            }
        }
    }
}

// mu * grad B(x) . p, this lane's share (sum over lanes = the full value)
template <int D>
__device__ __forceinline__ double barrier_dir(const double (&x)[D], const double (&p)[D], int norm, double mu, int lane) {
    double acc = 0.0;
    if (norm == 0) {
        if (lane != 0) return 0.0;
        double n2 = 0.0, xp = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) { n2 += x[j] * x[j]; xp += x[j] * p[j]; }
        acc = 2.0 * mu * xp / (1.0 - n2);
    } else if (norm == 1) {
        if (lane != 0) return 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) acc += 2.0 * mu * x[j] * p[j] / (1.0 - x[j] * x[j]);
    } else {
        for (int f = lane; f < (1 << D); f += 32) {
            double s = 1.0, sp = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const bool pj = (f >> j) & 1;
                s -= pj ? x[j] : -x[j];
                sp += pj ? p[j] : -p[j];
            }
            acc += mu * sp / s;
        }
    }
    return acc;
}

template <int D>
__device__ __forceinline__ bool strictly_inside(const double (&x)[D], int norm) {
    double m = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        if (norm == 0) m += x[j] * x[j];
        else if (norm == 1) m = fmax(m, fabs(x[j]));
        else m += fabs(x[j]);
    }
    return m < 1.0 - 1e-15;
}

// In-place modified Cholesky of the lower triangle of H, then b <- H^-1 b.
template <int D>
__device__ __forceinline__ void chol_solve_inplace(double (&H)[D][D], double (&b)[D]) {
    double dmax = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i) dmax = fmax(dmax, H[i][i]);
    const double floor_ = 1e-13 * fmax(dmax, 1e-300);
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double piv = H[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) piv -= H[j][k] * H[j][k];
        piv = piv < floor_ ? floor_ : piv;
        const double l = sqrt(piv), inv = 1.0 / l;
        H[j][j] = l;
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
            double s = H[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= H[i][k] * H[j][k];
            H[i][j] = s * inv;
        }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= H[i][k] * b[k];
        b[i] = s / H[i][i];
    }
#pragma unroll
    for (int i = D - 1; i >= 0; --i) {
        double s = b[i];
#pragma unroll
        for (int k = i + 1; k < D; ++k) s -= H[k][i] * b[k];
        b[i] = s / H[i][i];
    }
}

template <int D, typename TIN>
__global__ void __launch_bounds__(kExWarps * 32) exact_ftl_barrier_kernel(const ExactArgs p) {
    const int lane = threadIdx.x & 31;
    const int64_t K = p.T - p.k0 + 1;                           // prefixes solved per sequence
    const int64_t item = (int64_t)blockIdx.x * kExWarps + (threadIdx.x >> 5);
    if (item >= p.B * K) return;
    const int64_t b = item / K;
    const int64_t k = p.T - (item - b * K);                     // longest prefixes first
    if (p.only_flagged && p.only_flagged[b] == 0) return;
    const int64_t T = p.T;
    const TIN* __restrict__ zb = (const TIN*)p.z + b * T * D;
    const TIN* __restrict__ yb = (const TIN*)p.y + b * T;
    double* __restrict__ act = p.actions + b * (T + 1) * D;
    if (k == p.k0)                                              // prefix 0 (exact_ftl.py:294-297) and the skipped ones
        for (int64_t j = lane; j < p.k0 * D; j += 32) act[j] = 0.0;

    double sum_abs_y = 0.0;
    for (int64_t i = lane; i < k; i += 32) sum_abs_y += fabs((double)yb[i]);
    sum_abs_y = wsum(sum_abs_y);
    const double nu = p.norm == 0 ? 1.0 : (p.norm == 1 ? 2.0 * D : (double)(1 << D));
    const double mu_f = p.tol * fmax(1.0, 0.5 * sum_abs_y) / ((double)k + nu);
    double mu = fmax(fmax(sum_abs_y / (double)k, mu_f), 1e-3);

    double x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = 0.0;
    int newton = 0;
    bool done = false;
    while (newton < p.max_iter) {
        // gradient and Hessian of F_mu at x
        double g[D], H[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            g[i] = 0.0;
#pragma unroll
            for (int j = 0; j <= i; ++j) H[i][j] = 0.0;
        }
        const double mu2 = mu * mu;
        for (int64_t i = lane; i < k; i += 32) {
            double zr[D], r = -(double)yb[i];
#pragma unroll
            for (int j = 0; j < D; ++j) { zr[j] = (double)zb[i * D + j]; r += zr[j] * x[j]; }
            const double s = sqrt(mu2 + r * r), t = mu + s;
            const double c = r / t, w = mu / (t * s);
#pragma unroll
            for (int a = 0; a < D; ++a) {
                g[a] += c * zr[a];
                const double wz = w * zr[a];
#pragma unroll
                for (int j = 0; j <= a; ++j) H[a][j] += wz * zr[j];
            }
        }
        add_barrier<D>(x, p.norm, mu, lane, g, H);
        double dir[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            g[i] = wsum(g[i]);
            dir[i] = -g[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) H[i][j] = wsum(H[i][j]);
        }
        chol_solve_inplace<D>(H, dir);
        double gp = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) gp += g[j] * dir[j];
        if (!(-gp > kCenter * mu)) {                            // centred at this mu (also catches NaN)
            if (mu <= mu_f) { done = (gp == gp); break; }
            mu = fmax(mu * kTheta, mu_f);
            continue;
        }
        ++newton;
        // step length: largest a <= 1 that stays strictly inside, then the root of phi'(a)
        double a = 1.0, xa[D];
        for (int h = 0; h < 60; ++h) {
#pragma unroll
            for (int j = 0; j < D; ++j) xa[j] = x[j] + a * dir[j];
            if (strictly_inside<D>(xa, p.norm)) break;
            a *= 0.5;
        }
        double lo = 0.0, hi = -1.0, dlo = gp, dhi = 0.0;
        for (int e = 0; e < kLineEvals; ++e) {
#pragma unroll
            for (int j = 0; j < D; ++j) xa[j] = x[j] + a * dir[j];
            double da = 0.0;
            for (int64_t i = lane; i < k; i += 32) {
                double r = -(double)yb[i], zp = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) { const double zj = (double)zb[i * D + j]; r += zj * xa[j]; zp += zj * dir[j]; }
                da += r / (mu + sqrt(mu2 + r * r)) * zp;
            }
            da = wsum(da + barrier_dir<D>(xa, dir, p.norm, mu, lane));
            if (fabs(da) <= 0.2 * fabs(gp) || (da <= 0.0 && hi < 0.0)) break;
            if (da > 0.0) { hi = a; dhi = da; } else { lo = a; dlo = da; }
            const double sec = lo - dlo * (hi - lo) / (dhi - dlo), w = 0.1 * (hi - lo);
            a = fmin(fmax(sec, lo + w), hi - w);
        }
#pragma unroll
        for (int j = 0; j < D; ++j) x[j] += a * dir[j];
    }
    if (!done && lane == 0) atomicAdd(p.status + b, 1);
    if (lane < D) {
        double out = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) if (j == lane) out = x[j];
        act[k * D + lane] = out;
    }
}

template <typename TIN>
cudaError_t launch_exact_t(const ExactArgs& a, int d, cudaStream_t st) {
    const int64_t items = a.B * (a.T - a.k0 + 1);
    if (items <= 0) return cudaSuccess;
    const unsigned gx = (unsigned)((items + kExWarps - 1) / kExWarps);
    switch (d) {
    case 1: exact_ftl_barrier_kernel<1, TIN><<<gx, kExWarps * 32, 0, st>>>(a); break;
    case 2: exact_ftl_barrier_kernel<2, TIN><<<gx, kExWarps * 32, 0, st>>>(a); break;
    case 3: exact_ftl_barrier_kernel<3, TIN><<<gx, kExWarps * 32, 0, st>>>(a); break;
    case 4: exact_ftl_barrier_kernel<4, TIN><<<gx, kExWarps * 32, 0, st>>>(a); break;
    case 5: exact_ftl_barrier_kernel<5, TIN><<<gx, kExWarps * 32, 0, st>>>(a); break;
    case 6: exact_ftl_barrier_kernel<6, TIN><<<gx, kExWarps * 32, 0, st>>>(a); break;
    case 7: exact_ftl_barrier_kernel<7, TIN><<<gx, kExWarps * 32, 0, st>>>(a); break;
    case 8: exact_ftl_barrier_kernel<8, TIN><<<gx, kExWarps * 32, 0, st>>>(a); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace

// status must hold 0 for every sequence this launch solves (the kernel adds one per unconverged prefix).
cudaError_t launch_exact_ftl_general(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d, int norm,
                                     int max_iter, double tol, int64_t first_prefix, const int* only_flagged,
                                     double* actions, int* status, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (T == 0) return cudaMemsetAsync(actions, 0, sizeof(double) * (size_t)B * d, st);
    ExactArgs a;
    a.z = z; a.y = y; a.B = B; a.T = T; a.norm = norm; a.max_iter = max_iter; a.tol = tol;
    a.k0 = first_prefix < 1 ? 1 : (first_prefix > T ? T : first_prefix);
    a.only_flagged = only_flagged; a.actions = actions; a.status = status;
    return in_f32 ? launch_exact_t<float>(a, d, st) : launch_exact_t<double>(a, d, st);
}

}  // namespace oco
