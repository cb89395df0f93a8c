/*
 * b200id.h -- C ABI of the B200-native frequency-domain identification hot path
 *             (FRF estimation -> batched GP model selection -> Hermitian-IDFT FIR + validation).
 *
 * Drop-in boundary.  The reference (AtsushiYanaigsawa768/gauss-process) is pure Python and has
 * no FFI of its own; its boundary is the import surface of src/frequency_transform, src/gpr and
 * src/fir_model.  Each entry point below replaces the NumPy/SciPy arithmetic of the reference
 * lines cited next to it; the Python mirror in gauss-process_b200/src/ binds them with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - Every data pointer is a DEVICE pointer owned by the caller (FP64, contiguous) unless the
 *     parameter name ends in _host.  The library never allocates or frees caller-visible memory;
 *     scratch comes from the caller-provided workspace (query *_workspace_bytes first).
 *   - `stream` is a cudaStream_t passed as void*.  No call synchronises the host.
 *   - Return value: 0 = ok, < 0 = error (B200ID_E_*); text via b200id_last_error().
 *   - Complex arrays are interleaved (re, im) doubles.
 */
#ifndef B200ID_H
#define B200ID_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ID_VERSION 100

#define B200ID_OK            0
#define B200ID_E_ARG        -1   /* bad argument (null pointer, size <= 0, unknown id) */
#define B200ID_E_WORKSPACE  -2   /* workspace too small */
#define B200ID_E_CUDA       -3   /* CUDA runtime error (launch failed, wrong arch) */
#define B200ID_E_UNSUPPORTED -4  /* size outside what the kernels support */

/* kernel ids: src/gpr/kernels.py:386-411 (create_kernel names) */
enum b200id_kernel_id {
    B200ID_K_RBF = 0,            /* kernels.py:90   [length_scale, variance]        */
    B200ID_K_MATERN12 = 1,       /* kernels.py:113  nu = 0.5                        */
    B200ID_K_MATERN32 = 2,       /*                 nu = 1.5                        */
    B200ID_K_MATERN52 = 3,       /*                 nu = 2.5                        */
    B200ID_K_EXP = 4,            /* kernels.py:175  [omega, variance]               */
    B200ID_K_DC = 5,             /* kernels.py:227  [alpha, beta, rho]              */
    B200ID_K_DI = 6,             /* kernels.py:253  [beta, alpha]                   */
    B200ID_K_SS1 = 7,            /* kernels.py:282  [beta, variance]                */
    B200ID_K_SS2 = 8,            /* kernels.py:304                                  */
    B200ID_K_SSHF = 9,           /* kernels.py:332                                  */
    B200ID_K_STABLE_SPLINE = 10, /* kernels.py:358                                  */
    B200ID_K_RQ = 11,            /* kernels.py:150  [length_scale, variance, alpha] */
    B200ID_K_TC = 12,            /* kernels.py:201  [omega, variance]               */
    B200ID_K_MATERN_NU = 13,     /* kernels.py:131  general nu: [length_scale, variance] optimised,
                                    theta[2] = nu rides along as a fixed shape parameter (scipy.special.kv branch) */
    B200ID_K_COUNT = 14
};
#define B200ID_MAX_PARAMS 3

/* candidate metric: src/gpr/grid_search.py:91-133 */
#define B200ID_METRIC_NLL      0
#define B200ID_METRIC_VAL_RMSE 1
#define B200ID_FAIL_METRIC     1e10   /* grid_search.py:132-133, gpr_fitting.py:122-123 */

/* ------------------------------------------------------------------ library */
int         b200id_version(void);
const char* b200id_last_error(void);
/* Fills SM count and compute capability of the current device. */
int         b200id_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ FRF (K1)
 * Replaces the per-bin Python loop of synchronous_coefficients_trapz,
 * src/frequency_transform/frf_estimator.py:194-232 (hot loop :226-231), for u and y in one pass.
 *
 * Samples [0, n) are resident; the call integrates the OWNED samples [own_begin, own_end) with
 * trapezoid weights w_i = (t[min(i+1,n-1)] - t[max(i-1,0)]) / 2, so a time-segment shard passes
 * its slice plus a one-sample halo on each interior side.  Phases are fl(w_k * t_i) exactly as
 * the reference forms them (frf_estimator.py:227-228).
 */
size_t b200id_frf_workspace_bytes(int nd);

/* sums_out[0] = sum u[own), sums_out[1] = sum y[own) (y may be NULL -> 0). frf_estimator.py:219-220 */
int b200id_frf_moments(const double* u, const double* y, int64_t n,
                       int64_t own_begin, int64_t own_end,
                       double* sums_out /*[2]*/, void* ws, size_t ws_bytes, void* stream);

/* partial_out[4*nd] = { S_uc[nd], S_us[nd], S_yc[nd], S_ys[nd] } with
 *   S_xc[k] = sum_i w_i (x_i - mean_x) cos(fl(w_k t_i)),  S_xs likewise with sin.
 * mean_x = sums[.] * inv_count when sums != NULL (subtract_mean), else 0.  y may be NULL. */
int b200id_frf_project(const double* t, const double* u, const double* y, int64_t n,
                       int64_t own_begin, int64_t own_end,
                       const double* w, int nd,
                       const double* sums /*[2] or NULL*/, double inv_count,
                       double* partial_out /*[4*nd]*/, void* ws, size_t ws_bytes, void* stream);

/* Uniform-timebase fast path of b200id_frf_project (same outputs, same reference lines).  Valid when every
 * resident sample obeys t_i = t0 + i*dt + delta_i with w_max * |delta_i| < ~2e-6 (sampled data whose time
 * stamps differ from the grid only by FP64 rounding): the phase fl(w_k t_i) the reference forms is then
 * reproduced to first order from a rotation recurrence plus the exactly computed rounding terms, at about
 * half the FP64 work.  b200id_frf_uniformity returns max |delta_i| over i in [own_begin, own_end) (i counts
 * from the start of the resident slice, as everywhere in this section) so the caller can decide -- per
 * chunk, when a record is streamed to the device piecewise. */
int b200id_frf_uniformity(const double* t, int64_t n, int64_t own_begin, int64_t own_end, double t0, double dt,
                          double* max_abs_delta_out /*[1]*/, void* ws, size_t ws_bytes, void* stream);
int b200id_frf_project_uniform(const double* t, const double* u, const double* y, int64_t n,
                               int64_t own_begin, int64_t own_end, const double* w, int nd,
                               const double* sums /*[2] or NULL*/, double inv_count, double t0, double dt,
                               double* partial_out /*[4*nd]*/, void* ws, size_t ws_bytes, void* stream);

/* Exactly regenerable time base (what `np.arange(n) * dt`, MATLAB's `(0:n-1) * dt` and files saved from them hold):
 * b200id_frf_timebase_check writes to mismatches_out[0] the number of samples with t_i != fl(fl(i dt) + t0)
 * (mismatches_out needs TWO doubles; [1] is scratch) and b200id_frf_timebase_fill writes exactly that progression.
 * A host record whose stamps passed the check once can skip the upload of t afterwards (the load_time_u_y ->
 * synchronous_coefficients_trapz sequence of frf_estimator.py:70-134,194-232 run again on the same arrays):
 * u and y are the only per-call inputs.  Workspace: b200id_frf_workspace_bytes(1). */
int b200id_frf_timebase_check(const double* t, int64_t n, double t0, double dt, double* mismatches_out /*[2]*/,
                              void* ws, size_t ws_bytes, void* stream);
int b200id_frf_timebase_fill(double* t_out, int64_t n, double t0, double dt, void* stream);

/* Which recurrence b200id_frf_project_uniform runs for the well-conditioned bins: 0 = three-term recurrence on
 * the basis values (frf_project_uniform3_kernel), 1 = Goertzel recurrence on the accumulators with the rounding
 * term in packed FP32 (frf_project_uniform_g_kernel), 2 = bin-adaptive (the default; frf_adaptive.cuh: exact-phase
 * main term of every bin as FP64 matrix products, the reference's per-sample phase-rounding term only for the bins a
 * device-side test finds weakly excited; whole-record calls with y, anything else runs form 1), -1 = back to the
 * default (environment B200ID_FRF_FORM, else the build's default).  Process-wide, not stream-ordered: meant for
 * tests and for timing one form against the other.  Returns the previous setting. */
int b200id_frf_select_recurrence(int which);

/* Bin-adaptive form only.  b200id_frf_adaptive_force(1) lists every bin for the rounding term (0 = decide per bin,
 * 2 = list none: the main term alone, for timing its kernel -- NOT a valid result at weak bins, -1 = environment
 * B200ID_FRF_ADAPTIVE=all / default); returns the previous setting.  b200id_frf_adaptive_diag copies
 * {number of bins that took the rounding term, largest (16 sigma_k + rigorous_k) / |S_k| among the skipped bins} of
 * the LAST b200id_frf_project_uniform call on this workspace to diag_out[2] (device), stream-ordered. */
int b200id_frf_adaptive_force(int all_bins);
int b200id_frf_adaptive_diag(const void* ws, size_t ws_bytes, int nd, double* diag_out /*[2]*/, void* stream);

/* EXPERIMENTAL alternative evaluation of b200id_frf_project_uniform (same arguments, same outputs, same
 * reference lines frf_estimator.py:226-231): the exact-phase part as a blocked FP64 matrix product on
 * mma.sync.m8n8k4.f64 and the reference's per-sample phase rounding as a fixed-point / FP32 correction
 * (csrc/frf_mma.cuh).  Needs y != NULL and nd <= 160, else B200ID_E_UNSUPPORTED.  Parity-tested against the
 * recurrence kernels; currently slower than them, so b200id_frf_project_uniform only takes
 * this route when the environment has B200ID_FRF_MMA=1. */
int b200id_frf_project_uniform_mma(const double* t, const double* u, const double* y, int64_t n,
                                   int64_t own_begin, int64_t own_end, const double* w, int nd,
                                   const double* sums /*[2] or NULL*/, double inv_count, double t0, double dt,
                                   double* partial_out /*[4*nd]*/, void* ws, size_t ws_bytes, void* stream);

/* Selected DFT bins of a windowed, mean-removed record: what the Fourier estimator needs of its two n-point
 * FFTs (fourier_estimator.py:120-223 feeding np.interp at transform.py:160-181 -- only the FFT bins either
 * side of each grid frequency are ever read).  For k = bins[b] (0 <= k < n_fft):
 *   partial_out[0*nb+b] = sum_j a_j cos(2 pi j k / n_fft),  [1*nb+b] = sum_j a_j sin(...)   (a from u)
 *   partial_out[2*nb+b],  [3*nb+b]                                                        (a from y)
 * a_j = win_j (x_j - mean_x), mean from sums*inv_count (NULL -> 0), j over [own_begin, own_end) of the resident
 * arrays; X[k] = (cos sum) - i (sin sum).  window_mode 0 none, 1 periodic Hann (scipy get_window('hann', n)),
 * 2 window_vals[n].  window_sum_out [2] (Hann only, may be NULL) receives sum_j win_j in [0].
 * Workspace: b200id_frf_workspace_bytes(nb). */
int b200id_dft_bins(const double* u, const double* y /*may be NULL*/, int64_t n, int64_t own_begin, int64_t own_end,
                    const int64_t* bins /*[nb] device*/, int nb, int64_t n_fft, int window_mode,
                    const double* window_vals, const double* sums /*[2] or NULL*/, double inv_count,
                    double* partial_out /*[4*nb]*/, double* window_sum_out /*[2] or NULL*/,
                    void* ws, size_t ws_bytes, void* stream);

/* U[k] = scale * (S_uc[k] - j S_us[k]), Y likewise; scale = 2/T. frf_estimator.py:225-231 */
int b200id_frf_finalize(const double* partial /*[4*nd]*/, int nd, double scale,
                        double* U_out /*[2*nd]*/, double* Y_out /*[2*nd] or NULL*/, void* stream);

/* G[k] = sum_r Y_r conj(U_r) / sum_r |U_r|^2, NaN+NaNj where the denominator <= eps.
 * frf_cross_power_average, frf_estimator.py:239-252.  U, Y: [R][nd] complex. */
int b200id_cross_power(const double* U, const double* Y, int n_records, int nd, double eps,
                       double* G_out /*[2*nd]*/, void* stream);

/* The same average in two steps, for records spread over ranks and for the resident pipeline:
 *   b200id_cross_power_sums : sums_out [3*nd] = { Re sum_r Y conj(U), Im sum_r Y conj(U), sum_r |U|^2 }
 *                             (accumulate != 0 adds to what sums_out holds) -- the quantity that is all-reduced;
 *   b200id_gp_targets       : G_out [2*nd] = num / den (NaN + NaN j where den <= eps) and, optionally, what
 *                             gp_pipeline.py:130-132 builds from it: stats_out [4] = { mean Re G, mean Im G,
 *                             std Re G, std Im G } (population std, 0 -> 1: sklearn StandardScaler) and the
 *                             standardised targets y_re_out [nd], y_im_out [nd] (both NULL to skip; with
 *                             stats_out NULL they receive the raw real / imaginary parts instead). */
int b200id_cross_power_sums(const double* U, const double* Y, int n_records, int nd, int accumulate,
                            double* sums_out /*[3*nd]*/, void* stream);
int b200id_gp_targets(const double* sums /*[3*nd]*/, int nd, double eps, double* G_out /*[2*nd]*/,
                      double* stats_out /*[4] or NULL*/, double* y_re_out, double* y_im_out, void* stream);

/* ------------------------------------------------------------------ GP (K2, K3)
 * Gram matrices of the 13 kernels for 1-D inputs: src/gpr/kernels.py:90-382. */
int b200id_gp_gram(int kernel_id, const double* theta /*[p] device*/, const double* X1, int n1,
                   const double* X2 /*NULL -> X1*/, int n2, double* K_out /*[n1*n2] row-major*/,
                   void* stream);

/* Batched candidate evaluation, one CTA per candidate: Gram + noise*I (no jitter), in-shared-memory
 * Cholesky, two triangular solves, NLL or validation RMSE.  grid_search.py:91-133 and
 * gpr_fitting.py:104-123.  Non-positive pivot -> 1e10; NaN pivot -> NaN metric.
 *   theta [C][B200ID_MAX_PARAMS] (unused slots ignored); kernel_ids [C] or NULL (-> kernel_id);
 *   metric == VAL_RMSE uses Xv, yv (raw scale) and pred = (K* alpha) * y_scale + y_mean.
 *   metric_out [C]. */
size_t b200id_gp_batch_workspace_bytes(int64_t n_candidates, int n);
int b200id_gp_batch_eval(int kernel_id, const int* kernel_ids, const double* theta, int64_t n_candidates,
                         const double* X, const double* y, int n, double noise, int metric,
                         const double* Xv, const double* yv, int nv, double y_mean, double y_scale,
                         double* metric_out, void* ws, size_t ws_bytes, void* stream);
/* Two regression targets on the same inputs, candidates and noise (the Re and Im searches of the pipeline,
 * gp_pipeline.py:125-156): K + noise I and its Cholesky factor are common, only the right-hand side differs, so
 * one factorisation per candidate serves both (n <= 126; larger n runs the two evaluations one after the other).
 * Outputs metric_out_a [C], metric_out_b [C]; a failed factorisation gives 1e10 in both. */
int b200id_gp_batch_eval_pair(int kernel_id, const int* kernel_ids, const double* theta, int64_t n_candidates,
                              const double* X, const double* y_a, const double* y_b, int n, double noise, int metric,
                              const double* Xv, const double* yv_a, const double* yv_b, int nv,
                              double y_mean_a, double y_scale_a, double y_mean_b, double y_scale_b,
                              double* metric_out_a, double* metric_out_b, void* ws, size_t ws_bytes, void* stream);
/* Candidate lists that are products of 1-D grids (grid_search.py:173: itertools.product of the per-parameter grids):
 * in every kernel but DI the amplitude theta[1] is the last factor of a covariance entry, k = fl(theta[1] * s(x, x')),
 * so candidates that differ only in theta[1] share the shape values s.  shape_theta [S][B200ID_MAX_PARAMS] holds one
 * theta per distinct shape (its amplitude slot is ignored), shape_of [C] the shape index of every candidate; the S
 * shape tables are evaluated first (one launch) and every candidate forms its entries from its table by the same
 * final multiplication as the direct evaluation: metrics bit-identical to b200id_gp_batch_eval(_pair).
 * y_b / yv_b / metric_out_b NULL -> one target.  n <= 127 (126 with two targets); one kernel id per call.
 * Workspace: b200id_gp_shared_workspace_bytes(S, n, nv (0 for NLL), two_targets). */
size_t b200id_gp_shared_workspace_bytes(int n_shapes, int n, int nv, int two_targets);
int b200id_gp_batch_eval_shared(int kernel_id, const double* theta, int64_t n_candidates, const int* shape_of,
                                const double* shape_theta, int n_shapes,
                                const double* X, const double* y_a, const double* y_b, int n, double noise, int metric,
                                const double* Xv, const double* yv_a, const double* yv_b, int nv,
                                double y_mean_a, double y_scale_a, double y_mean_b, double y_scale_b,
                                double* metric_out_a, double* metric_out_b, void* ws, size_t ws_bytes, void* stream);
/* Same, with one noise variance PER CANDIDATE (noise [C], device): the finite-difference points of the
 * L-BFGS-B hyperparameter optimisation perturb log(noise) as well, gpr_fitting.py:104-123,130,147 -- the
 * value and all forward-difference points of one gradient are one launch. */
int b200id_gp_batch_eval_noise(int kernel_id, const int* kernel_ids, const double* theta, const double* noise,
                               int64_t n_candidates, const double* X, const double* y, int n, int metric,
                               const double* Xv, const double* yv, int nv, double y_mean, double y_scale,
                               double* metric_out, void* ws, size_t ws_bytes, void* stream);

/* First strict minimum of metric[0..C) scanning in index order, NaNs skipped (grid_search.py:186-196).
 * best_out[0] = metric (+inf if none), idx_out[0] = index (-1 if none). */
int b200id_first_strict_min(const double* metric, int64_t n_candidates,
                            double* best_out, int64_t* idx_out, void* ws, size_t ws_bytes, void* stream);

/* The same scan per segment [offsets[s], offsets[s+1]) of one metric array (offsets: n_segments + 1 device int64), one
 * launch for all segments: the per-kernel minima of a mixed-kernel candidate population (BASELINE config 3, the
 * restart rule of gpr_fitting.py:132-145 applied to all 11 kernels of run_baseline.py:37-41 at once).
 * pend_out [n_segments][2] = (best metric or +inf, index into metric[] as a double or -1): what b200id_argmin_pack takes. */
int b200id_first_strict_min_segments(const double* metric, const int64_t* offsets, int n_segments,
                                     double* pend_out /*[n_segments][2]*/, void* stream);

/* The two small steps either side of the ranks' argmin exchange (candidates are sharded round
 * robin, every rank reports (best metric, global index), the first strict minimum in GLOBAL scan order wins --
 * grid_search.py:186-196 over the whole list), for n_searches independent searches at once, as one launch each instead
 * of a few dozen tensor operations:
 *   argmin_pack : pend [S][2] = (best metric, LOCAL index as a double, -1 = none) -> mine [S][2] = (best metric,
 *                 global index = global_index_map[local] (or local itself when the map is NULL), -1 = none)
 *   argmin_pick : gathered [W][S][2] (rank-major; W = 1 without an exchange) -> out [S][5] = (theta[0..2] of the winner
 *                 from full_theta [C][B200ID_MAX_PARAMS], global index or -1, best metric or +inf); ties go to the
 *                 smallest global index, NaN metrics and empty shards are skipped. */
int b200id_argmin_pack(const double* pend, int n_searches, const int64_t* global_index_map, double* mine_out, void* stream);
int b200id_argmin_pick(const double* gathered, int n_ranks, int n_searches, const double* full_theta, double* out,
                       void* stream);

/* Final fit: K + diag_add*I, Cholesky, alpha = K^-1 y, K_inv (from the factor).  gpr_fitting.py:73-83.
 * Kinv_out may be NULL when only alpha is needed (n <= 127).
 * info_out[0] = 0 ok, k > 0 = non-positive pivot at column k (caller then uses b200id_gp_fit_pinv). */
size_t b200id_gp_fit_workspace_bytes(int n);
int b200id_gp_fit(int kernel_id, const double* theta, const double* X, const double* y, int n,
                  double diag_add, double* alpha_out /*[n]*/, double* Kinv_out /*[n*n]*/,
                  int* info_out, void* ws, size_t ws_bytes, void* stream);

/* B final fits that share X, the kernel family and the diagonal term (the Re and Im models of one
 * identification, gp_pipeline.py:125-156 -> gpr_fitting.py:73-83) in one launch: theta [B][B200ID_MAX_PARAMS],
 * y [B][n] -> alpha_out [B][n], info_out [B] (0 = ok, >0 = failing pivot -> use b200id_gp_fit_pinv for that
 * problem).  K_inv is not produced.  Workspace: B * b200id_gp_fit_workspace_bytes(n). */
int b200id_gp_fit_batch(int kernel_id, const double* theta, const double* X, const double* y, int n,
                        int n_problems, double diag_add, double* alpha_out, int* info_out,
                        void* ws, size_t ws_bytes, void* stream);

/* Pseudo-inverse of the symmetric matrix K + diag_add*I by cyclic Jacobi eigendecomposition,
 * dropping |lambda| <= rcond * max|lambda|; alpha = pinv @ y.  gpr_fitting.py:84-87 (np.linalg.pinv). */
int b200id_gp_fit_pinv(int kernel_id, const double* theta, const double* X, const double* y, int n,
                       double diag_add, double rcond, double* alpha_out, double* Kinv_out,
                       void* ws, size_t ws_bytes, void* stream);

/* mean = K(Xs, X) alpha; std = sqrt(max(k(xs,xs) - sum_j (K* K_inv)_j K*_j, 0)) (std_out may be NULL).
 * gpr_fitting.py:89-100. */
int b200id_gp_predict(int kernel_id, const double* theta, const double* X, int n,
                      const double* alpha, const double* Kinv,
                      const double* Xs, int m, double* mean_out, double* std_out, void* stream);

/* Blocked FP64 Cholesky (lower, in place, row-major) for large n (N_d = 2048): panel factorisation in
 * shared memory, trailing update as FP64 DGEMM tiles.  np.linalg.cholesky at gpr_fitting.py:80.
 * info_out[0] as in b200id_gp_fit. */
size_t b200id_chol_blocked_workspace_bytes(int n);
int b200id_chol_blocked(double* A, int n, int lda, int* info_out, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ kernel-regularised FIR (time domain)
 * Reference src/fir_model/kernel_regularized.py.  kernel: 0 = 'dc' (theta = [log c, logit lam, atanh rho]),
 * 1 = 'ss' ([log c, logit lam]), 2 = 'si' ([logit r, logit(omega / pi), log q, logit barlam]); every theta row
 * carries log sigma^2 as its last entry (kernel_regularized.py:192-217, :236-238).
 *
 * b200id_kr_summaries: G = U^T U [L*L], b = U^T y [L], yy = y^T y [1] of the zero-padded Toeplitz regressor
 * (:84-92, :337-340), all on the device.
 * b200id_kr_nll_batch: negative log marginal likelihood (:224-261) of `batch` hyper-parameter points, one CTA each.
 *   info_out[b]: 0 ok, 2 = S took the reference's 1e-6 fallback jitter (:246-250), 1 = K + jitter I not positive
 *   definite, 3 = S not positive definite even so; nll_out[b] is NaN for 1 and 3 (the reference raises LinAlgError).
 * b200id_kr_posterior: g_hat [L], sigma^2 [1] and (optionally) K [L*L] for one point (:264-291, jitter 1e-10);
 *   with g_out == NULL only the covariance K of theta's kernel part is built (_build_kernel, :192-217).
 * Workspace: b200id_kr_workspace_bytes(L, batch) (batch = 1 for summaries and posterior). */
size_t b200id_kr_workspace_bytes(int L, int batch);
int b200id_kr_summaries(const double* u, const double* y, int64_t n, int L, double* G_out, double* b_out,
                        double* yy_out, void* ws, size_t ws_bytes, void* stream);
int b200id_kr_nll_batch(int kernel, const double* theta /*[batch][kdim+1]*/, int batch, const double* G,
                        const double* b, double yy, double n_samples, int L, double jitter, double* nll_out,
                        int* info_out, void* ws, size_t ws_bytes, void* stream);
int b200id_kr_posterior(int kernel, const double* theta /*[kdim+1]*/, const double* G, const double* b, int L,
                        double* g_out, double* K_out /*may be NULL*/, double* sig2_out, int* info_out, void* ws,
                        size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ FIR (K5, K6)
 * h[m] = Re( (1/M) sum_k X[k] e^{+2 pi j k m / M} ), m < n_taps, with X the two-sided Hermitian
 * extension of G_pos (M = 2 nd - 1, X[0] = Re G[0]).  fir_helpers.py:75-137. */
int b200id_idft_hermitian(const double* G_pos /*[2*nd]*/, int nd, double* h_out, int n_taps, void* stream);

/* Organisation of b200id_fir_eval's overlap-save transform: 0 = radix-16 passes (fir_fft16_kernel, default),
 * 1 = radix-4 passes (fir_fft_kernel; same butterflies in the same order, hence the same values), -1 = back to the
 * default (environment B200ID_FIR_RADIX4).  Process-wide, for tests and timing comparisons.  Returns the previous
 * setting. */
int b200id_fir_select_fft(int radix4);

/* Causal FIR y_pred[i] = sum_{j<n_taps, j<=i+lead} g[j] u[i-j] over OWNED outputs [own_begin, own_end)
 * of a resident slice (u has `lead` samples of history before index 0 of y: u points at the slice start,
 * u[-lead..-1] must be readable when lead > 0), fused with the error sums over i >= skip_global:
 *   sums_out = { sum e^2, sum y^2, sum (y - y0)^2, count }   with e = y - y_pred.
 * np.convolve(u, g, 'full')[:N] at fir_fitting.py:285 / fir_validation.py:190 and the metrics at
 * fir_fitting.py:288-306 / fir_validation.py:115-128.  y_pred_out may be NULL. */
size_t b200id_fir_workspace_bytes(int64_t n);
int b200id_fir_eval(const double* u, const double* y, int64_t n, int64_t lead,
                    const double* g, int n_taps, int64_t skip, double y0,
                    double* y_pred_out, double* sums_out /*[4]*/, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ measurement helpers */
/* Runs an FP64 FMA (mode 0) or FP64 mma.sync (mode 1) issue-bound loop; returns flops executed in
 * flops_out_host and leaves timing to the caller's CUDA events.  Used by bench.py for the FP64 roofline. */
int b200id_fp64_peak_probe(int mode, int iters, double* sink, double* flops_out_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200ID_H */
