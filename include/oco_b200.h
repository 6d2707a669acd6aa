/*
 * oco_b200.h -- C ABI of the B200-native engine for the online-convex-optimization hot path
 * of revvu/online_convex_optimization (FTL / FTRL / SMART trajectory loops + the numpy-keyed
 * sequence generators).  Plain pointers and sizes only; no torch / numpy types.
 *
 * The reference is pure Python: its "FFI" for this path is the module seam at
 * fast_driver.py:23-29 (names imported from fast_algorithms / sequence_generation).  Each
 * entry point below names the reference function(s) whose per-call work it batches; the
 * Python module online_convex_optimization_b200.fast_algorithms binds them with ctypes and
 * re-exports the reference's names and signatures.
 *
 * Conventions
 *   - every function returns an int status (OCO_OK == 0); oco_last_error() gives the message
 *     of the calling thread's last failure.  OCO_ERR_VALUE maps to Python ValueError,
 *     OCO_ERR_RUNTIME (CUDA failures) to RuntimeError, mirroring exact_ftl.py:125-126,142-143.
 *   - sequences are row-major z[B][T][d], y[B][T] (the reference's (T, d) / (T,) arrays,
 *     fast_algorithms.py:171-177, stacked over B), dtype OCO_F32 or OCO_F64; float32 is
 *     widened to float64 on load exactly like np.ascontiguousarray(z, dtype=float64).
 *   - alg_flag: 0 = FTRL, non-zero = FTL (fast_algorithms.py:100-103).  A = number of
 *     (alg_flag, eta0) variants run on every sequence; results are [B][A] row-major.
 *   - functions without the _host suffix take DEVICE pointers for bulk data and a
 *     cudaStream_t (as void*); they enqueue work and return without synchronising.
 *     alg_flag / eta0 / theta_thresh / keys / T_grid are always small HOST arrays.
 *   - *_host functions take HOST pointers, perform their own (chunked, double-buffered)
 *     host<->device copies on `device`, and return after synchronising.
 *   - the caller owns every buffer; the library frees all scratch it allocates.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *     with OCO_ERR_RUNTIME.
 */
#ifndef OCO_B200_H
#define OCO_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define OCO_API __attribute__((visibility("default")))
#else
#define OCO_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define OCO_OK 0
#define OCO_ERR_VALUE 1
#define OCO_ERR_RUNTIME 2

#define OCO_F32 0
#define OCO_F64 1

#define OCO_ALG_FTRL 0
#define OCO_ALG_FTL 1

/* sequence families (sequence_generation.py / fast_algorithms.py:231-239) */
#define OCO_FAM_GPATH 0   /* standard_normal((T,d)) f64, rows clipped to unit norm, choice([-1,1],T) */
#define OCO_FAM_IID 1     /* make_random_iid_stream: f32 rows, y = sign(z@u)            (:54-70)  */
#define OCO_FAM_NOISY 2   /* make_noisy_iid_stream: same + Massart label flips p        (:72-90)  */
#define OCO_FAM_FLIP 3    /* flip_sequence                                              (:24-28)  */
#define OCO_FAM_SWITCH 4  /* switching_two_leaders_sequence                             (:36-47)  */

#define OCO_MAX_VARIANTS 64
#define OCO_MAX_SMART_THRESHOLDS 4

/* comparator_mode for oco_simulate_batch*:
 * 0 = second pass over the sequence, the literal fast_algorithms.py:113-114 (bit-exact for d <= 8);
 * 1 = single pass using 0.5|z.a - y| = 0.5(1 - y z.a) when ||z_t|| <= 1, y = +-1 (checked per
 *     row; sequences that violate it fall back to the second pass).  Tolerance 1e-12 relative. */
#define OCO_COMPARATOR_TWO_PASS 0
#define OCO_COMPARATOR_LINEAR 1
/* oco_g_sweep_mode_host only: the EXACT maximum at single-pass speed.  For labels +-1, |q - y| >= 1 - y q, so the
 * single-pass regret (OCO_COMPARATOR_LINEAR) is an upper bound of the literal one for every run; only the runs whose
 * bound reaches the largest bound (minus a margin covering worst-case rounding) can attain the maximum, and only
 * those -- typically one run -- take the literal fast_algorithms.py:113-114 second pass.  g_max is the reference's
 * value bit for bit; the sum / sum of squares in `stats` are the single-pass ones (1e-12 class). */
#define OCO_COMPARATOR_EXACT_MAX 2

OCO_API const char* oco_last_error(void);
OCO_API int oco_abi_version(void);
/* number of visible CUDA devices (0 and OCO_OK when there is none) */
OCO_API int oco_device_count(int* count);

/* ---- fast_algorithms.simulate_alg (fast_algorithms.py:171-177 -> _simulate_alg_core :88-115),
 * batched over B sequences and A variants; also exact_ftl._simulate_ftrl (exact_ftl.py:230-277)
 * when `comparator` is given.  comparator: optional [B*d] fixed comparator action per sequence
 * (exact_ftl.run_ftrl(..., comparator_action=...)); NULL = the reference's FTL action of the
 * final theta (d <= 8 only).  Outputs ([B*A] unless noted, any may be NULL): regret =
 * cum_loss - comp_loss, cum_loss, comp_loss, theta_final [B*A*d], x_last [B*A*d] (the action
 * played at the last round, RunResult.x_last; d <= 8), x_trace [B*A*T*d] (all played x_t; d <= 8).
 * compute_dtype: OCO_F64 = the numba path fast_algorithms.py (float64 state, the default and the
 * parity target); OCO_F32 = the NumPy path algorithms.py:28-54 (float32 theta / x, float64 loss
 * accumulator; needs in_dtype == OCO_F32, d <= 8, no comparator / x_trace; tolerance 1e-5). */
OCO_API int oco_simulate_batch(const void* z, const void* y, int in_dtype, int compute_dtype, int64_t B, int64_t T, int d,
                       const int* alg_flag, const double* eta0, int A, const double* comparator,
                       double* regret, double* cum_loss, double* comp_loss,
                       double* theta_final, double* x_last, double* x_trace, int comparator_mode, void* stream);
OCO_API int oco_simulate_batch_host(const void* z, const void* y, int in_dtype, int compute_dtype, int64_t B, int64_t T, int d,
                            const int* alg_flag, const double* eta0, int A, const double* comparator,
                            double* regret, double* cum_loss, double* comp_loss,
                            double* theta_final, double* x_last, double* x_trace, int comparator_mode, int device);

/* ---- exact_ftl.compute_prefix_actions + replay_exact_ftl (exact_ftl.py:280-333) for the unit
 * l2 ball, batched, d <= 256, in the LINEAR REGIME only (every row ||z_t||_2 <= 1, y_t = +-1;
 * there the SOCP of exact_ftl.py:83-105 has the closed-form minimiser v_k/||v_k||,
 * v_k = sum_{i<k} y_i z_i, and 0 on ties).  actions [B*(T+1)*d] (may be NULL), cum_loss /
 * comp_loss / regret [B] (may be NULL), status [B]: 0 = ok, 1 = some row is outside the regime
 * (results for that sequence are NOT the exact optimum and must be discarded).
 * PARITY UNPINNED: the reference's arithmetic for this path lives in CVXPY's conic solver. */
OCO_API int oco_exact_ftl_prefix_batch(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d,
                               double* actions, double* cum_loss, double* comp_loss, double* regret,
                               int* status, void* stream);
OCO_API int oco_exact_ftl_prefix_batch_host(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d,
                                    double* actions, double* cum_loss, double* comp_loss, double* regret,
                                    int* status, int device);

/* ---- general exact-FTL prefix solver: any rows / labels, unit ball norm 0 = l2, 1 = linf, 2 = l1
 * (exact_ftl.py:67-105 `norm=`), d <= 32 (l1: d <= 16).  For l2, sequences in the linear regime take the closed
 * form above; all others are solved, every prefix independently, by a path-following interior-point
 * method (Newton on the log-barrier function with the epigraph variables eliminated, see
 * csrc/kernels_exact.cu; d > 8: lane per variable, csrc/kernels_wide_ipm.cu) until the duality gap is <= tol * max(1, f(0)), f(0) = 0.5*sum|y_i|
 * (suggested tol 1e-8, what conic solvers deliver by default), using at most `max_iter` Newton steps
 * per prefix (suggested 300; 20-30 are typical).  status [B]: 0 = all prefixes converged, n > 0 =
 * n prefixes stopped at max_iter (treat like the reference's solver failure, exact_ftl.py:125-126).
 * Prefixes shorter than `first_prefix` are not solved (their actions are zeros; 0 or 1 = solve all,
 * T = only the full sequence, which is what one ExactFTLNoClip.solve_prefix call needs -- prefixes
 * are solved independently, so the result is the same bits either way).  actions [B*(T+1)*d] is
 * required; cum_loss / comp_loss / regret [B] (may be NULL) are the replay of those actions.  An
 * independent solver, not a restatement of the reference's CVXPY stack: PARITY UNPINNED, validated
 * against scipy (HiGHS / SLSQP) in the tests. */
OCO_API int oco_exact_ftl_general_batch(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d, int norm,
                                int max_iter, double tol, int64_t first_prefix, double* actions, double* cum_loss, double* comp_loss,
                                double* regret, int* status, void* stream);
OCO_API int oco_exact_ftl_general_batch_host(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d, int norm,
                                     int max_iter, double tol, int64_t first_prefix, double* actions, double* cum_loss, double* comp_loss,
                                     double* regret, int* status, int device);

/* ---- exact_ftl.replay_exact_ftl (exact_ftl.py:306-333) for caller-supplied actions [B*(T+1)*d]:
 * cum_loss = sum_t 0.5|z_t.actions[t] - y_t|, comp_loss = 0.5 sum_t |z_t.actions[T] - y_t|. d <= 256
 * (bit-exact summation order for d <= 8, warp-parallel 1e-12 above). */
OCO_API int oco_replay_actions_batch_host(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d,
                                  const double* actions, double* cum_loss, double* comp_loss, double* regret, int device);

/* ---- fast_algorithms.simulate_SMART_like (fast_algorithms.py:184-195 -> :118-164), batched;
 * A <= OCO_MAX_SMART_THRESHOLDS (theta_thresh[a], eta0[a]) pairs per sequence, d <= 256 (d > 8: lane-distributed
 * states, csrc/kernels_wide_smart.cu, 1e-12 class, both modes run the exact kernel with its pruning bound).
 * switch_t [B*A] (may be NULL) = step at which the switch test fired, -1 if never.
 * smart_mode 0: the reference's O(T^2) prefix-loss recomputation (:157-158), spread over a warp.
 * smart_mode 1: O(T) -- in the linear regime (||z_t|| <= 1, y_t = +-1) the leader's prefix loss is
 *   0.5*((t+1) - s.sum_i y_i z_i); sequences with a violating row are re-run by mode 0. */
OCO_API int oco_smart_batch(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d,
                    const double* theta_thresh, const double* eta0, int A,
                    double* regret, int64_t* switch_t, int smart_mode, void* stream);
OCO_API int oco_smart_batch_host(const void* z, const void* y, int in_dtype, int64_t B, int64_t T, int d,
                         const double* theta_thresh, const double* eta0, int A,
                         double* regret, int64_t* switch_t, int smart_mode, int device);

/* ---- sequence generators (sequence_generation.py:24-100, fast_algorithms.py:231-239).
 * Sequence b is drawn from np.random.Generator(PCG64(SeedSequence([a_b, T, c_b]))) with
 * (a_b, c_b) = (keys[2b], keys[2b+1]) when keys != NULL, else (key_a, key_c0 + b).
 * g-path: a = base_seed, c = run.  iid / noisy: a = run_seed, c = 13 + rep / 23 + rep.
 * Writes z[B][T][d], y[B][T] in out_dtype (device / host). */
OCO_API int oco_gen_sequences(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                      int64_t B, int64_t T, int d, double p_flip, int64_t block_len,
                      int out_dtype, void* z, void* y, void* stream);
OCO_API int oco_gen_sequences_host(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                           int64_t B, int64_t T, int d, double p_flip, int64_t block_len,
                           int out_dtype, void* z, void* y, int device);

/* ---- fused generate + simulate + compare for seeded families GPATH (d <= 8) / IID / NOISY (d <= 5):
 * the sequences never exist in memory.  This is the inner double loop of
 * fast_algorithms.empirical_worst_case_thresholds (:226-243) and of
 * fast_driver.evaluate_stream_with_stats (:87-111, FTRL / FTL columns) as one launch.
 * regret / cum_loss / comp_loss: [B*A] or NULL.  stats: [A*4] = (max, sum, sum of squares,
 * count) of the regrets per variant, reduced in a fixed order (bit-reproducible), or NULL. */
OCO_API int oco_regret_sweep(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                     int64_t B, int64_t T, int d, double p_flip,
                     const int* alg_flag, const double* eta0, int A,
                     double* regret, double* cum_loss, double* comp_loss, double* stats, void* stream);
OCO_API int oco_regret_sweep_host(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                          int64_t B, int64_t T, int d, double p_flip,
                          const int* alg_flag, const double* eta0, int A,
                          double* regret, double* cum_loss, double* comp_loss, double* stats, int device);

/* Same with an explicit comparator_mode (the entry points above are OCO_COMPARATOR_TWO_PASS, bit-exact):
 * OCO_COMPARATOR_LINEAR applies to the g-path family with 2 <= d <= 8 (d = 6, 7, 8 exist in this mode only; other
 * shapes ignore it and stay exact):
 * rows are regenerated in registers instead of being parked in a scratch slab (no DRAM traffic) and the
 * comparator loss of fast_algorithms.py:113-114 is T/2 - ||theta_T|| -- valid because the kernel verifies
 * sign(q_t - y_t) == -y_t on every step; a sequence that ever violates it takes the literal second pass.
 * cum_loss stays bit-exact; regret / comp_loss agree with the reference to 1e-12 * max(1, cum_loss). */
OCO_API int oco_regret_sweep_mode(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                          int64_t B, int64_t T, int d, double p_flip,
                          const int* alg_flag, const double* eta0, int A,
                          double* regret, double* cum_loss, double* comp_loss, double* stats,
                          int comparator_mode, void* stream);
OCO_API int oco_regret_sweep_mode_host(int family, uint64_t key_a, uint64_t key_c0, const uint64_t* keys,
                               int64_t B, int64_t T, int d, double p_flip,
                               const int* alg_flag, const double* eta0, int A,
                               double* regret, double* cum_loss, double* comp_loss, double* stats,
                               int comparator_mode, int device);

/* ---- fast_algorithms.empirical_worst_case_thresholds (fast_algorithms.py:211-247) for runs
 * [run_lo, run_lo + n_runs): g_max[i] = max(0, max_r regret_FTRL(eta0)) on g-path sequences of
 * length T_grid[i]; stats [nT*4] as above (may be NULL).  d <= 8 (the reference uses 5). */
OCO_API int oco_g_sweep_host(const int64_t* T_grid, int nT, uint64_t run_lo, int64_t n_runs, uint64_t base_seed,
                     int d, double eta0, double* g_max, double* stats, int device);
OCO_API int oco_g_sweep_mode_host(const int64_t* T_grid, int nT, uint64_t run_lo, int64_t n_runs, uint64_t base_seed,
                          int d, double eta0, double* g_max, double* stats, int comparator_mode, int device);

/* ---- known-answer probes of the device generator (tests): what = 0 raw uint64, 1 standard
 * normals, 2 random() doubles, 3 choice([-1,1]) labels, 4 SeedSequence state words, 5 log1p(-U);
 * n_skip normals are consumed first.  out: HOST buffer of n 8-byte items.
 * what = 6 checks the kernels' branch-free reciprocal against IEEE division: n threads x n_skip
 * pseudo-random doubles each (seed a); out[0] (uint64) = number of differing results (must be 0). */
OCO_API int oco_rng_probe_host(int what, uint64_t a, uint64_t b, uint64_t c, int64_t n_skip, int64_t n, void* out, int device);

/* ---- scratch (the sweep kernels' row slab, scale tables, staging buffers of the *_host calls) is allocated
 * stream-ordered from a PRIVATE memory pool per device owned by this library and kept cached there between calls
 * (the process-wide default pool and its release threshold are never touched); this returns it to the driver. */
OCO_API int oco_release_scratch(int device);

/* ---- roofline denominator for the generated-on-device kernels: measured issue rate of
 * independent non-fused fp64 multiplies and adds (ops / second) on `device`. */
OCO_API int oco_fp64_peak_probe_host(int device, double* ops_per_sec);

#ifdef __cplusplus
}
#endif
#endif /* OCO_B200_H */
