// oco_internal.h -- host-side launcher declarations shared by the .cu files and abi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oco {

constexpr int kMaxVariants = 64;   // A, (algorithm, eta0) variants per sequence
constexpr int kMaxSmallD = 8;      // thread-per-trajectory kernels (registers hold theta)

enum Family {                      // sequence families
    FAM_GPATH = 0,                 // fast_algorithms.py:231-239  float64
    FAM_IID = 1,                   // sequence_generation.py:54-70   float32
    FAM_NOISY = 2,                 // sequence_generation.py:72-90   float32, Massart flips
    FAM_FLIP = 3,                  // sequence_generation.py:24-28
    FAM_SWITCH = 4,                // sequence_generation.py:36-47
};

// ---- scratch: a private stream-ordered memory pool per device, owned by the library (kernels_seeded.cu).
// Nothing is released back to the driver between calls (release threshold = max) -- the sweep kernels' row slab is
// tens of GB and re-reserving it costs more than a short sweep -- but the process-wide DEFAULT pool is never touched.
cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st);
cudaError_t scratch_trim();                      // current device: give cached scratch back to the driver
size_t scratch_cached_free_bytes();              // current device: reserved by the pool but not in use (reusable by the next call)

struct Variants {                  // passed by value to kernels
    int A;
    int alg_flag[kMaxVariants];    // 0 = FTRL, non-zero = FTL (fast_algorithms.py:100-103)
};

struct Etas { double e[kMaxVariants]; };

// tab[t*A + a] = -(eta0[a] / sqrt(max(1, t+1)))
cudaError_t launch_ftrl_scale_table(double* tab, const Etas& eta0, int64_t T, int A, cudaStream_t st);

// Canonical layout of the variants of one sweep: FTL ignores eta0 (fast_algorithms.py:102-103), so
// one FTL trajectory serves all FTL columns; FTRL variants are processed nf_tpl at a time.
struct SweepPlan {
    int A;                              // output columns
    int n_ftl, n_ftrl;                  // number of FTL / FTRL columns
    int nf_tpl;                         // FTRL variants per thread (template NF: 1, 4 or 8)
    int groups;                         // grid.y = max(1, ceil(n_ftrl / nf_tpl))
    int ftl_col[kMaxVariants];
    int ftrl_col[kMaxVariants + 8];     // padded to groups * nf_tpl
    double ftrl_eta[kMaxVariants + 8];  // padding: 0.0
};
void make_sweep_plan(const Variants& v, const Etas& e, SweepPlan* plan);
inline int64_t sweep_table_elems(const SweepPlan& pl, int64_t T) { return (int64_t)pl.groups * pl.nf_tpl * T; }
// tab[(g*T + t)*nf_tpl + k] = -(eta[g*nf_tpl + k] / sqrt(max(1, t+1)))
cudaError_t launch_sweep_table(double* tab, const SweepPlan& plan, int64_t T, cudaStream_t st);

// Fused generate + simulate + compare (no HBM traffic for inputs).  Sequence b has
// stream key SeedSequence([a, T, c]) with (a, c) = keys[b] when keys != NULL (device
// [B][2]), else (key_a, key_c0 + b).  regret/cum/comp: device [B*A] or NULL.
// partials: device [blocks * A * 3] (max, sum, sumsq per block) or NULL.
cudaError_t launch_sweep(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys, int64_t T, int64_t B, int d,
                         double p_flip, const SweepPlan& plan, const double* sweep_tab,
                         double* regret, double* cum, double* comp, double* partials, int comparator_mode, cudaStream_t st);
cudaError_t launch_sweep_gpath(const struct SweepArgs& a, int d, cudaStream_t st);        // round-1 kernel, d = 1 .. 5 (OCO_SWEEP_LEGACY=1)
// g-path, 2 <= d <= 5, single-pass comparator, rows regenerated instead of parked in HBM (gsweep_impl.cuh)
cudaError_t launch_gsweep_lo(const struct SweepArgs& a, int d, cudaStream_t st);   // d = 1, 2, 3
cudaError_t launch_gsweep_hi(const struct SweepArgs& a, int d, cudaStream_t st);   // d = 4, 5
cudaError_t launch_gsweep_wide(const struct SweepArgs& a, int d, cudaStream_t st); // d = 6, 7, 8
cudaError_t launch_sweep_iid(const struct SweepArgs& a, int d, cudaStream_t st);
int sweep_gpath_blocks(int64_t n_runs);
// candidates for the exact maximum: indices i with regret[i * stride] >= stats_max[0] - margin (at most cap are stored)
cudaError_t launch_select_candidates(const double* regret, int64_t n, int stride, const double* stats_max, double margin,
                                     int64_t* idx, int* count, int cap, cudaStream_t st);
// stats[a*4 + {0,1,2,3}] = max, sum, sumsq, count over all runs (fixed reduction order)
cudaError_t launch_stats_finalize(const double* partials, int nblocks, int A, int64_t n_runs, double* stats, cudaStream_t st);

// Fill kernels: write the sequences of one family to device memory, row-major
// z[B][T][d], y[B][T], as float64 (out_f32 == 0) or float32.  Keys as for launch_sweep.
cudaError_t launch_fill(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys, int64_t B, int64_t T, int d,
                        double p_flip, int64_t block_len, int out_f32, void* z, void* y, cudaStream_t st);

// warp-per-sequence generator for the g-path family, any d <= 128 (kernels_fill_warp.cu)
bool fill_warp_eligible(int family, int d);
cudaError_t launch_fill_warp(uint64_t key_a, uint64_t key_c0, const uint64_t* keys, int64_t B, int64_t T, int d, int out_f32,
                             void* z, void* y, cudaStream_t st);

// Streamed simulators on device-resident sequences (z[B][T][d], y[B][T]; f32 or f64).
cudaError_t launch_sim_stream(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                              const Variants& v, const double* scale_tab, const double* comparator,
                              double* regret, double* cum, double* comp, double* theta_final, double* x_last,
                              double* x_trace, int comparator_mode, cudaStream_t st);
// float32 compute mode (algorithms.py semantics), float32 inputs, d <= 8
cudaError_t launch_sim_stream_f32(const void* z, const void* y, int64_t B, int64_t T, int d,
                                  const Variants& v, const double* scale_tab,
                                  double* regret, double* cum, double* comp, double* theta_final, double* x_last,
                                  cudaStream_t st);
// exact FTL prefix actions + replay in the linear regime (||z||<=1, y=+-1), d <= 8
cudaError_t launch_exact_ftl_linear(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                                    double* actions, double* cum, double* comp, double* regret, int* status,
                                    cudaStream_t st);
cudaError_t launch_sim_wide(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                            const Variants& v, const double* scale_tab,
                            double* regret, double* cum, double* comp, double* theta_final, int comparator_mode,
                            cudaStream_t st);
// TMA (1-D bulk async copy + mbarrier) version of launch_sim_wide; requires wide_tma_eligible()
bool wide_tma_eligible(const void* z, const void* y, int in_f32, int64_t T, int d);
cudaError_t launch_sim_wide_tma(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                                const Variants& v, const double* scale_tab,
                                double* regret, double* cum, double* comp, double* theta_final, int comparator_mode,
                                cudaStream_t st);
// general exact-FTL prefix solver (ADMM), any unit ball (0 = l2, 1 = linf, 2 = l1), d <= 8
cudaError_t launch_exact_ftl_general(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d, int norm,
                                     int max_iter, double tol, int64_t first_prefix, const int* only_flagged,
                                     double* actions, int* status, cudaStream_t st);
// the same solver for 8 < d <= kWideIpmMaxD, l2 / linf balls (kernels_wide_ipm.cu: lane per variable, Hessian in shared memory)
constexpr int kWideIpmMaxD = 32;
cudaError_t launch_exact_ftl_wide_general(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d, int norm,
                                          int max_iter, double tol, int64_t first_prefix, const int* only_flagged,
                                          double* actions, int* status, cudaStream_t st);
cudaError_t launch_replay_actions(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                                  const double* actions, double* cum, double* comp, double* regret, cudaStream_t st);
// 8 < d <= 256 twins of the three launchers above (kernels_wide_exact.cu: one warp per sequence / per (sequence, variant))
cudaError_t launch_wide_exact_prefix(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                                     double* actions, double* cum, double* comp, double* regret, int* status, cudaStream_t st);
cudaError_t launch_wide_replay(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                               const double* actions, double* cum, double* comp, double* regret, cudaStream_t st);
cudaError_t launch_wide_fixed_comp(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                                   const Variants& v, const double* scale_tab, const double* comparator,
                                   double* regret, double* cum, double* comp, double* theta_final, double* x_last, cudaStream_t st);
cudaError_t launch_status_scale(int* status, int64_t B, int factor, cudaStream_t st);   // status[b] *= factor
// SMART: all A <= 4 thresholds of a sequence in one warp; scale_tab is [T][A] (eta0 per threshold).
cudaError_t launch_smart(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                         const double* thresh_dev, const double* scale_tab, int A,
                         double* regret, int64_t* switch_t, int mode, int* flags, cudaStream_t st);

// SMART for 8 < d <= 256 (kernels_wide_smart.cu): lane-distributed states, exact pruning of the prefix loss, 1e-12 class
cudaError_t launch_smart_wide(const void* z, const void* y, int in_f32, int64_t B, int64_t T, int d,
                              const double* thresh_dev, const double* scale_tab, int A,
                              double* regret, int64_t* switch_t, cudaStream_t st);

// non-FMA fp64 issue-rate microbenchmark (16 ops per thread per iteration, 256 threads per block)
cudaError_t launch_fp64_probe(double* out, int blocks, int iters, cudaStream_t st);

// RNG probes for known-answer tests (device out[n])
cudaError_t launch_rng_probe(int what, uint64_t a, uint64_t b, uint64_t c, int64_t n_skip, int64_t n, void* out, cudaStream_t st);

}  // namespace oco
