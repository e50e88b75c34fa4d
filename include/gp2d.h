/* gp2d -- C ABI of the B200-native Helmholtz Gaussian-process hot path.
 *
 * Drop-in boundary for rafaelcgon/2D-GP's kernel / fit / predict / likelihood path
 * (SURVEY.md §8b).  Plain C: pointers and sizes only, no torch or C++ types.  Unless a
 * function name ends in _host, every data pointer is a CALLER-OWNED DEVICE pointer to
 * IEEE fp64, `stream` is a cudaStream_t passed as void*, calls are asynchronous on that
 * stream, allocate nothing (workspace is sized by the *_workspace_bytes queries), keep no
 * global state and are re-entrant across streams (the _host entry points cache one device buffer per
 * host thread, see the end of this file).  The library also exports bring-up hooks gp2d_dbg_* that are not
 * declared here (tests and bench.py use them: forced GEMM tile shape, forced predict split, FP64 peak probe);
 * the overrides they set are per host thread and off by default, so they never couple two callers.
 *
 * Return value: 0 = launched OK; -k = argument k (1-based) is invalid;
 * <= -1000 = -(1000 + cudaError_t).  Numerical failure of the factorisation is reported
 * LAPACK-style through an `int* info` in device memory: 0, or the 1-based index of the
 * first non-positive pivot (the matrix is not positive definite).
 *
 * Layouts (the reference's): X is [N,2] row-major (x, y); y is the stacked observation
 * vector [first component; second component] of length 2N (GP_laser.py:98,174);
 * covariance matrices are component-major blocks, row c*N+i, column c'*M+j, row-major
 * with leading dimension ld (myKernel.py:40-43; GP_scripts.py:89-95); predictive outputs
 * are [component][point].  theta = (l_df, l_cf, ratio); `noise` is a variance.
 */
#ifndef GP2D_H
#define GP2D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP2D_VERSION 100

int gp2d_version(void);
/* Human-readable text for a negative return code (static storage). */
const char* gp2d_error_string(int code);

/* ---- covariance ------------------------------------------------------------------ */

/* K[2N,2M] = ratio*K_divfree + (1-ratio)*K_curlfree between X[N,2] and X2[M,2].
 * X2 == NULL means X2 = X (then diag_add is added on the diagonal).
 * Replaces myKernel.myKernel.K(X,X2) (myKernel.py:27-53), nonDivK.K (ratio=1,
 * myKernel.py:159-176), nonRotK.K (ratio=0, myKernel.py:255-271), GP_scripts.myKernel
 * (GP_scripts.py:6-42), compute_K (GP_scripts.py:74-95) and, with the roles of X and X2
 * swapped, compute_Ks (GP_scripts.py:97-123). */
int gp2d_kernel_build(const double* X, int N, const double* X2, int M,
                      double l_df, double l_cf, double ratio, double diag_add,
                      double* K, int64_t ldk, void* stream);

/* out[2M] = ratio/l_df^2 + (1-ratio)/l_cf^2.  Replaces myKernel.Kdiag (myKernel.py:55-57),
 * nonDivK.Kdiag (:178-180), nonRotK.Kdiag (:273-275). */
int gp2d_kdiag(int M, double l_df, double l_cf, double ratio, double* out, void* stream);

/* out3 = sum(dK/dtheta * dL_dK) for theta = (l_df, l_cf, ratio); dL_dK is [2N,2M] in the
 * block layout.  reference_compat != 0 reproduces the integrands of
 * myKernel.update_gradients_full (myKernel.py:59-106) verbatim; 0 gives the analytic
 * derivative (the reference's length-scale terms are not the derivative of its K). */
size_t gp2d_kernel_grad_workspace_bytes(int N, int M);
int gp2d_kernel_grad(const double* X, int N, const double* X2, int M,
                     double l_df, double l_cf, double ratio, int reference_compat,
                     const double* dL_dK, int64_t ld, void* ws, size_t ws_bytes,
                     double* out3, void* stream);

/* ---- dense factorisation --------------------------------------------------------- */

/* In-place lower Cholesky of the row-major SPD matrix A[n,n] (upper triangle untouched).
 * Replaces scipy/LAPACK dpotrf as used by GPy and sklearn. */
size_t gp2d_potrf_workspace_bytes(int n);
int gp2d_potrf(double* A, int n, int64_t lda, void* ws, size_t ws_bytes, int* info, void* stream);

/* In-place inverse of the row-major SPD matrix A[n,n] (full symmetric result): Cholesky,
 * L^-1 and K^-1 = L^-T L^-1 on the DMMA pipe.  Replaces np.linalg.inv(K) at
 * GP_laser.py:118,180 / GP_scripts.py:50 and GPy's dpotri.  The factor panels are refined
 * against the factor (one step of iterative refinement per panel), so the result keeps the
 * quality of a LAPACK inverse on ill-conditioned input.
 *
 * Conditioning of the fit entry points (gp2d_fit, gp2d_lml_grad and their _st / _hsum / _rbf
 * twins): the factorisation forms panels of L with explicit block inverses, which is accurate to
 * the 1e-8 parity bar while n k(x,x) / (noise + jitter) stays below ~1e9; past 1e7 the fit switches
 * to the same refined panels by itself (same results to rounding for well-conditioned input, ~55 %
 * more time) and iterates alpha against the matrix, so alpha and the predictive MEAN keep the accuracy
 * of a backward-stable solve.  The fused predictive VARIANCE applies the explicit inverse factor: for covariances beyond
 * that bound use the iterated solve the Python engine builds from gp2d_*_kernel_build,
 * gp2d_spd_inverse and gp2d_dgemm (engine.refined_predict). */
size_t gp2d_spd_inverse_workspace_bytes(int n);
int gp2d_spd_inverse(double* A, int n, int64_t lda, void* ws, size_t ws_bytes, int* info, void* stream);

/* C[M,N] = alpha * op(A) * op(B) + beta * C, row-major, fp64, on the DMMA pipe.
 * op(A) is M x K: A[m*lda + k] (transa = 0) or A[k*lda + m] (transa = 1); op(B) is K x N:
 * B[k*ldb + n] (transb = 0) or B[n*ldb + k] (transb = 1).  M and N must be multiples of
 * 128, K of 16, leading dimensions even, pointers 16-byte aligned (callers pad).
 * Replaces the np.dot calls of GP_scripts.getMean/getCov (GP_scripts.py:44-54). */
int gp2d_dgemm(int transa, int transb, int M, int N, int K, double alpha, const double* A, int64_t lda,
               const double* B, int64_t ldb, double beta, double* C, int64_t ldc, void* stream);

/* ---- GP fit / predict / likelihood ----------------------------------------------- */

/* Fit: build K + (noise+jitter) I, factorise, alpha = K^-1 y,
 * LML = -1/2 y'alpha - sum log L_ii - N log 2pi.  The workspace then holds the fit state
 * consumed by gp2d_predict (layout private, a pure function of N).
 * alpha_out (2N, block order) may be NULL.  lml_out: 1 double, info: 1 int (device).
 * Replaces GPy GPRegression(X,Y,k) (GP_plots.py:763; krig.py:411), the numpy
 * K + noise*I ; inv path (GP_laser.py:113-118,177-180) and sklearn .fit (krig.py:182-185). */
size_t gp2d_fit_workspace_bytes(int N);
int gp2d_fit(const double* X, int N, const double* y,
             double l_df, double l_cf, double ratio, double noise, double jitter,
             void* ws, size_t ws_bytes, double* alpha_out, double* lml_out, int* info,
             void* stream);

/* The part of a fit workspace that gp2d_predict reads (L^-1 tiles, alpha, X, LML, info) is one
 * contiguous byte range [offset, offset + bytes): copy or broadcast it into a workspace of the
 * same N on another GPU and gp2d_predict works there (grid shards of one snapshot on several
 * GPUs; the factorisation itself stays on one GPU). */
int gp2d_fit_predict_state(int N, size_t* offset, size_t* bytes);

/* Predict at Xs[M,2] from a fit state: mean[c*out_stride + j], var[c*out_stride + j],
 * c in {0,1}, j in [0,M).  var = k** - |L^-1 k*|^2 clamped at 0, plus var_add (pass the
 * noise variance to reproduce GPy's predict, 0 for sklearn/GP_laser).
 * Replaces GPy model.predict (krig.py:543-544,600-601), getMean + diag of Cov
 * (GP_scripts.py:44-46; GP_laser.py:129-131) and sklearn predict(return_std=True)
 * (krig.py:194).  Independent grid shards may be issued on different streams / GPUs.
 *
 * ws is a per-call scratch (not shared between concurrent calls): one [npad x 128] K*
 * panel per streaming multiprocessor; gp2d_predict_workspace_bytes(N, M) sizes it for the
 * current device.  A smaller buffer is accepted down to one panel (fewer resident CTAs). */
size_t gp2d_predict_workspace_bytes(int N, int M);
int gp2d_predict(const void* fit_ws, int N, double l_df, double l_cf, double ratio,
                 const double* Xs, int M, int64_t out_stride, double var_add,
                 double* mean, double* var, void* ws, size_t ws_bytes, void* stream);

/* LML and its gradient in one pass: out5 = (LML, dLML/dl_df, dLML/dl_cf, dLML/dratio,
 * dLML/dnoise).  Leaves a valid fit state in ws.  Replaces one objective evaluation of
 * optimize_restarts (krig.py:450; GP_plots.py:765): kernel build + Cholesky + dL_dK +
 * update_gradients_full + noise gradient (sklearn _gpr.py:583-656). */
int gp2d_lml_grad(const double* X, int N, const double* y,
                  double l_df, double l_cf, double ratio, double noise, double jitter,
                  int reference_compat, void* ws, size_t ws_bytes, double* out5, int* info,
                  void* stream);

/* ---- batches of independent problems of one size ----------------------------------------------
 * B problems with the same N in ONE chain of launches (every kernel of the fit covers the whole batch:
 * B diagonal blocks factorise side by side instead of one SM working while 147 idle).  This is what the
 * reference's independent units map to: the restarts of optimize_restarts (krig.py:450; GP_plots.py:765 --
 * same X, y, different theta: x_stride = y_stride = 0) and the time slices / snapshots of the predict loop
 * (krig.py:541-557 -- different X, y: strides >= 2N doubles).
 * Problem b reads X + b x_stride, y + b y_stride, theta4[4 b .. 4 b + 3] = (l_df, l_cf, ratio, noise) (HOST
 * array, read before the call returns) and owns the workspace ws + b gp2d_fit_workspace_bytes(N); after the
 * call that workspace is a fit state like gp2d_fit's (pass it to gp2d_predict / gp2d_fit_predict_state).
 * alpha_out [B][2N] may be NULL; lml_out [B]; out5 [B][5] as gp2d_lml_grad; info [B] (device).
 * Every problem gets the arithmetic of its single-problem call: results are bit-identical to gp2d_fit /
 * gp2d_lml_grad (ill-conditioned problems, which need the refined factorisation, run one by one). */
size_t gp2d_fit_batched_workspace_bytes(int N, int B);
int gp2d_fit_batched(const double* X, int64_t x_stride, int N, const double* y, int64_t y_stride, int B,
                     const double* theta4, double jitter, void* ws, size_t ws_bytes, double* alpha_out,
                     double* lml_out, int* info, void* stream);
int gp2d_lml_grad_batched(const double* X, int64_t x_stride, int N, const double* y, int64_t y_stride, int B,
                          const double* theta4, double jitter, int reference_compat, void* ws, size_t ws_bytes,
                          double* out5, int* info, void* stream);

/* ---- space-time product kernel ---------------------------------------------------------------
 * K = tvar * exp(-dt^2 / (2 lt^2)) * Helmholtz(l_df, l_cf, ratio)(a - a', b - b'): the product
 * Kt(t) * nonDivK(y, x) of scratch.py:506-508 with Kt of myKernel.py:337-363 (an RBF in time tiled
 * over the 2x2 blocks), generalised to the combined Helmholtz kernel.  Points are [N,3] row-major
 * (t, a, b); everything else (block layout, stacked observations, workspaces, return codes) is as
 * in the functions above.  gp2d_st_kernel_grad: out5 = sums for (l_df, l_cf, ratio, tvar, lt).
 * gp2d_st_lml_grad: out7 = (LML, d/dl_df, d/dl_cf, d/dratio, d/dtvar, d/dlt, d/dnoise).
 * gp2d_kernel_grad_workspace_bytes and gp2d_predict_workspace_bytes size the scratch here too. */
int gp2d_st_kernel_build(const double* X3, int N, const double* X3b, int M, double l_df, double l_cf, double ratio,
                         double tvar, double lt, double diag_add, double* K, int64_t ldk, void* stream);
int gp2d_st_kernel_grad(const double* X3, int N, const double* X3b, int M, double l_df, double l_cf, double ratio,
                        double tvar, double lt, const double* dL_dK, int64_t ld, void* ws, size_t ws_bytes,
                        double* out5, void* stream);
size_t gp2d_st_fit_workspace_bytes(int N);
int gp2d_st_fit_predict_state(int N, size_t* offset, size_t* bytes);
int gp2d_st_fit(const double* X3, int N, const double* y, double l_df, double l_cf, double ratio, double tvar,
                double lt, double noise, double jitter, void* ws, size_t ws_bytes, double* alpha_out,
                double* lml_out, int* info, void* stream);
int gp2d_st_predict(const void* fit_ws, int N, double l_df, double l_cf, double ratio, double tvar, double lt,
                    const double* Xs3, int M, int64_t out_stride, double var_add, double* mean, double* var,
                    void* ws, size_t ws_bytes, void* stream);
int gp2d_st_lml_grad(const double* X3, int N, const double* y, double l_df, double l_cf, double ratio, double tvar,
                     double lt, double noise, double jitter, void* ws, size_t ws_bytes, double* out7, int* info,
                     void* stream);

/* ---- scalar ARD-RBF sum family --------------------------------------------------------------
 * k(x,x') = sum_{q<Q} var[q] exp(-1/2 sum_{d<D} ((x_d - x'_d) / ls[q*D+d])^2), D <= 4, Q <= 4,
 * scalar observations y[N].  var[Q] and ls[Q*D] are HOST arrays (read before the call returns);
 * all other pointers are device pointers as above.  This is GPy.kern.RBF(input_dim=3, ARD=True)
 * summed nKernels times (krig.py:388,405-407) and scikit-learn's HP[0]*RBF([..]) + HP[4]*RBF([..])
 * (krig.py:174-178); the white-noise term (WhiteKernel krig.py:179 / GPy Gaussian_noise) is the
 * `noise` argument of the fit and `var_add` of the prediction.  Matrices are plain [N,M]
 * row-major. */
int gp2d_rbf_kernel_build(const double* X, int N, const double* X2, int M, int D, int Q, const double* var,
                          const double* ls, double diag_add, double* K, int64_t ldk, void* stream);

/* out[Q*(1+D)] = sum(dK/dtheta * dL_dK), theta ordered (var_q, ls[q*D .. q*D+D-1]) per component:
 * GPy's RBF.update_gradients_full with ARD. */
size_t gp2d_rbf_kernel_grad_workspace_bytes(int N, int M);
int gp2d_rbf_kernel_grad(const double* X, int N, const double* X2, int M, int D, int Q, const double* var,
                         const double* ls, const double* dL_dK, int64_t ld, void* ws, size_t ws_bytes,
                         double* out, void* stream);

/* Fit / predict / likelihood, same contracts as gp2d_fit, gp2d_predict, gp2d_lml_grad.
 * Replaces GPRegression(X, vo, RBF+RBF...) (krig.py:409-412), GaussianProcessRegressor(kernel=k,
 * optimizer=None).fit(XT,u) (krig.py:182-185) and .predict(X, return_std=True) (krig.py:194; pass
 * var_add = noise: WhiteKernel is part of sklearn's predictive variance, and of GPy's).
 * gp2d_rbf_lml_grad: out[2 + Q*(1+D)] = (LML, d/dtheta as in gp2d_rbf_kernel_grad, d/dnoise). */
size_t gp2d_rbf_fit_workspace_bytes(int N, int D);
int gp2d_rbf_fit_predict_state(int N, int D, size_t* offset, size_t* bytes);
int gp2d_rbf_fit(const double* X, int N, int D, const double* y, int Q, const double* var, const double* ls,
                 double noise, double jitter, void* ws, size_t ws_bytes, double* alpha_out, double* lml_out,
                 int* info, void* stream);
size_t gp2d_rbf_predict_workspace_bytes(int N, int M);
int gp2d_rbf_predict(const void* fit_ws, int N, int D, int Q, const double* var, const double* ls,
                     const double* Xs, int M, double var_add, double* mean, double* variance, void* ws,
                     size_t ws_bytes, void* stream);
int gp2d_rbf_lml_grad(const double* X, int N, int D, const double* y, int Q, const double* var, const double* ls,
                      double noise, double jitter, void* ws, size_t ws_bytes, double* out, int* info, void* stream);

/* ---- sum of space-time Helmholtz terms ----------------------------------------------------------
 * K = sum_{q<Q} var_q exp(-dt^2/2lt_q^2 - da^2/2la_q^2 - db^2/2lb_q^2) * G_q(da, db), Q <= 8, with
 *   G (type 0, divergence-free) = [[(1 - db^2/lb^2)/lb^2,  da db/(la^2 lb^2)], [., (1 - da^2/la^2)/la^2]]
 *   G (type 1, curl-free)       = [[(1 - da^2/la^2)/la^2, -da db/(la^2 lb^2)], [., (1 - db^2/lb^2)/lb^2]]
 * These are the kernels krig.kriging(kernelType = 2, 3, 4, nKernels) builds from the module myKernel2
 * that is missing from the reference (krig.py:396-407): divFreeK / curlFreeK(input_dim=3, var, lt, ly, lx),
 * their sum, and nKernels copies of it.  With la == lb, ldx == 2 and the two terms
 * {type 0, var = ratio} + {type 1, var = 1 - ratio} the matrix equals gp2d_kernel_build's
 * (myKernel.py:27-53).  type[Q] and params[Q*4] = (var, lt, la, lb) per term are HOST arrays (read
 * before the call returns); ldx = 3: points are [N,3] rows (t, a, b); ldx = 2: rows (a, b), lt unused.
 * Layouts, stacking, workspaces and return codes as in the Helmholtz functions above;
 * gp2d_predict_workspace_bytes sizes the prediction scratch.
 * gp2d_hsum_kdiag: out[2M] = prior variances (first M: component 0).
 * gp2d_hsum_kernel_grad: out[Q*4] = sum(dK/d(var, lt, la, lb)_q * dL_dK) (0 for lt when ldx == 2).
 * gp2d_hsum_lml_grad: out[2 + 4Q] = (LML, the 4Q derivatives in the same order, d/dnoise). */
int gp2d_hsum_kernel_build(const double* X, int N, const double* X2, int M, int ldx, int Q, const int* type,
                           const double* params, double diag_add, double* K, int64_t ldk, void* stream);
int gp2d_hsum_kdiag(int M, int ldx, int Q, const int* type, const double* params, double* out, void* stream);
size_t gp2d_hsum_kernel_grad_workspace_bytes(int N, int M, int Q);
int gp2d_hsum_kernel_grad(const double* X, int N, const double* X2, int M, int ldx, int Q, const int* type,
                          const double* params, const double* dL_dK, int64_t ld, void* ws, size_t ws_bytes,
                          double* out, void* stream);
size_t gp2d_hsum_fit_workspace_bytes(int N, int ldx, int Q);
int gp2d_hsum_fit_predict_state(int N, int ldx, int Q, size_t* offset, size_t* bytes);
int gp2d_hsum_fit(const double* X, int N, int ldx, const double* y, int Q, const int* type, const double* params,
                  double noise, double jitter, void* ws, size_t ws_bytes, double* alpha_out, double* lml_out,
                  int* info, void* stream);
int gp2d_hsum_predict(const void* fit_ws, int N, int ldx, int Q, const int* type, const double* params,
                      const double* Xs, int M, int64_t out_stride, double var_add, double* mean, double* var,
                      void* ws, size_t ws_bytes, void* stream);
int gp2d_hsum_lml_grad(const double* X, int N, int ldx, const double* y, int Q, const int* type, const double* params,
                       double noise, double jitter, void* ws, size_t ws_bytes, double* out, int* info, void* stream);

/* ---- options (per host thread) ----------------------------------------------------------- */

/* gp2d_set_option(key, value): 0, -1 (unknown key) or -2 (bad value).  Options are thread-local: they apply
 * to the calls the setting thread makes afterwards and never couple two callers.
 * GP2D_OPT_PREDICT_I8 -- the predictive pass of the Helmholtz families (gp2d_predict, gp2d_st_predict,
 *   gp2d_fit_predict_host) has two kernels: the fp64 tensor-pipe kernel and an int8-sliced tcgen05 kernel
 *   (5x faster at the reference configurations; exact integer products of base-256 digit slices, deterministic;
 *   digit slices that are identically zero -- distant points -- are skipped, which is why the fits keep the
 *   observations in a spatial order internally (invisible at this boundary: alpha_out is in the caller's order);
 *   agrees with the fp64 kernel to ~2e-10 relative on the variance with 6 slices, ~1e-12 with 7, at the
 *   prior-to-noise ratio 4 of the reference configurations).  0 (default): gp2d_fit picks the slice count from m = k** / (noise + jitter) *
 *   sqrt(n / 4000) -- 6 up to m = 20, 7 up to m = 2000, the fp64 kernel beyond, in the robust (ill-conditioned)
 *   mode and for N > 32768; 1: fp64 kernel only; 6 / 7: that slice count whenever N <= 32768.
 *   The choice is made by gp2d_fit (it prepares the slices) and honoured by the predict calls on that workspace;
 *   set the option before the fit. */
#define GP2D_OPT_PREDICT_I8 1
int gp2d_set_option(int key, double value);

/* ---- host-buffer convenience (allocates, copies, synchronises) -------------------- */

/* Whole fit + predict with HOST pointers (pageable is fine) on the current device, legacy default
 * stream; returns info (> 0) or an error code (< 0).  lml may be NULL.  This is what the numpy route
 * of the reference binds in one call: GP_laser.simLaser's K + noise I ; inv ; getMean ; diag(Cov)
 * (GP_laser.py:177-185) and GPRegression(X,Y,k).predict(Xnew) (GP_plots.py:763-768).
 * The device buffer (fit workspace, inputs, outputs, K* panels) is cached PER HOST THREAD between
 * calls and grown on demand -- the only state the library keeps; it is freed when the thread exits
 * or by gp2d_host_release() (current thread). */
int gp2d_fit_predict_host(const double* X, int N, const double* y,
                          double l_df, double l_cf, double ratio, double noise, double jitter,
                          const double* Xs, int M, int include_noise,
                          double* mean, double* var, double* lml);
void gp2d_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* GP2D_H */
