// Internal host-side entry points shared between the translation units of libgp2d.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "helmholtz.cuh"
#include "rbf.cuh"
#include "hsum.cuh"

namespace gp2d {

// linalg.cu -----------------------------------------------------------------------------
// Lower Cholesky of the row-major SPD matrix A (n multiple of 128) and, when need_inv,
// Z = L^-1 (lower; upper off-diagonal tiles untouched).  A is destroyed unless keep_L
// (then its lower triangle holds L and W must be an (n/2)^2 scratch).  logdiag[n] receives
// log(L_ii); *info the 1-based index of the first non-positive pivot, else 0.
// t_refine > 0 (needs keep_L and W): every off-diagonal panel of L gets that many steps of iterative
// refinement against the factor (robust mode for ill-conditioned matrices, see potri_rec).
// batch > 1: that many independent problems of the same size, problem b at A + b bstride, Z + b bstride,
// logdiag + b bstride (doubles) and info + 2 b bstride (ints) -- one launch per step of the recursion for
// the whole batch (plain mode only: no keep_L, no refinement).
cudaError_t potri_lower(double* A, long lda, double* Z, long ldz, int n, double* logdiag, int* info,
                        bool need_inv, bool keep_L, double* W, cudaStream_t st, int t_refine = 0, int batch = 1,
                        long bstride = 0);
void set_potri_overlap(bool on);      // bring-up switch: side-stream overlap of the inverse GEMMs
// alpha = Z^T Z y and LML from Z, logdiag.  y_block is the caller's vector: stacked [u;v] of
// length 2N (ncomp = 2) or N scalar observations (ncomp = 1); everything else is internal
// (pair-interleaved for ncomp = 2, padded to npad).
// batch > 1: problem b at every workspace pointer + b bstride doubles, its observations at y_block + b y_bstride.
cudaError_t solve_alpha_lml(const double* Z, long ldz, int npad, int N, int ncomp, const double* y_block,
                            double* y_int, double* w, double* alpha_int, double* partial,
                            const double* logdiag, double* lml_out, cudaStream_t st, int batch = 1, long y_bstride = 0,
                            long bstride = 0, const int* perm = nullptr, long p_bstride = 0);
// robust mode: alpha += Z^T Z (y - K alpha), `steps` times; K = lower tiles of the padded covariance,
// t1 / t2 scratch vectors of npad doubles (t1 may be the w of solve_alpha_lml: it is not needed afterwards)
cudaError_t refine_alpha(const double* K, long ldk, const double* Z, long ldz, int npad, const double* y_int,
                         double* t1, double* t2, double* alpha_int, double* partial, int steps, cudaStream_t st);
// perm (optional, order.cu): internal observation i is the caller's observation perm[i]; batch b at perm + b p_bstride
cudaError_t deinterleave(const double* xi, int N, double* x, cudaStream_t st, int batch = 1, long bstride = 0,
                         const int* perm = nullptr, long p_bstride = 0);
// Z-order (Morton) permutation of N <= spatial_order_max_points() points by the coordinates in columns xo, xo + 1 of
// X[N][ldx]; gather_points writes X in that order.  batch: problem b at X + b x_bstride, perm + b p_bstride (ints),
// out + b o_bstride.
int spatial_order_max_points();
cudaError_t spatial_order(const double* X, int ldx, int xo, int N, int* perm, cudaStream_t st, int batch = 1, long x_bstride = 0,
                          long p_bstride = 0);
cudaError_t gather_points(const double* X, int ldx, int N, const int* perm, double* out, cudaStream_t st, int batch = 1,
                          long x_bstride = 0, long p_bstride = 0, long o_bstride = 0);
// Tile-major, pre-swizzled copy of the lower triangle of Z (what predict_fused streams with
// bulk copies): tile (row block li, k-tile kt) is number 8 li (li+1)/2 + kt, 2048 doubles each.
size_t packed_tiles_doubles(int npad);
cudaError_t pack_lower_tiles(const double* Z, long ldz, int npad, double* Zt, cudaStream_t st, int batch = 1, long bstride = 0);

// kernel_build.cu -------------------------------------------------------------------------
// Reference (component-major block) layout, arbitrary N, M, ld: K[2N, 2M].  X2 == nullptr
// means X2 = X.  diag_add is added where row == col (only meaningful for X2 == X).
cudaError_t build_block_layout(const double* X, int N, const double* X2, int M, const HelmParams& hp,
                               double diag_add, double* K, long ldk, cudaStream_t st);
// Internal layout: pair-interleaved (row 2i+c), padded to npad with identity, lower
// 128x128 tiles only (diagonal tiles complete).
cudaError_t build_interleaved_lower(const double* X, int N, const HelmParams& hp, double diag_add,
                                    double* K, long ldk, int npad, cudaStream_t st);
// Per-problem parameters of a batch, kept in each problem's workspace (set by scatter_batch_params).
struct BatchPar { HelmParams hp; double diag_add; };
// batch of problems bstride doubles apart: problem b reads its points at X + b x_bstride, its BatchPar at
// par + b bstride and writes K + b bstride
cudaError_t build_interleaved_lower_batched(const double* X, long x_bstride, int N, const BatchPar* par, double* K, long ldk,
                                            int npad, int batch, long bstride, cudaStream_t st);
cudaError_t scatter_batch_params(const BatchPar* host, int batch, BatchPar* par, long bstride, const double* X,
                                 long x_bstride, int x_doubles, double* Xws, cudaStream_t st);
// sum(dK/dtheta * dL_dK) for theta = (l_df, l_cf, ratio); dL_dK in block layout [2N,2M].
cudaError_t kernel_grad_sums_block(const double* X, int N, const double* X2, int M, const HelmParams& hp,
                                   int compat, const double* dL_dK, long ld, double* partial,
                                   int partial_cap, double* out, cudaStream_t st);
int grad_sums_block_partials(int N, int M);

// scalar ARD-RBF sum (rbf.cuh): K[N,M] row-major; internal padded lower tiles; sum(dK/dtheta * dL_dK)
// with out[Q (1 + D)] ordered (var_q, l_{q,0..D-1}) per component; partial holds
// (rbf_grad_partials + 1) * RBF_MAXQ * (1 + RBF_MAXD) doubles.
cudaError_t rbf_build(const double* X, int N, const double* X2, int M, const RbfParams& rp, double diag_add,
                      double* K, long ldk, cudaStream_t st);
cudaError_t rbf_build_padded_lower(const double* X, int N, const RbfParams& rp, double diag_add, double* K,
                                   long ldk, int npad, cudaStream_t st);
int rbf_grad_partials(int N, int M);
cudaError_t rbf_grad_sums(const double* X, int N, const double* X2, int M, const RbfParams& rp, const double* dL_dK,
                          long ld, double* partial, int partial_cap, double* out, cudaStream_t st);

// sum of space-time Helmholtz terms (hsum.cuh): same layouts as the Helmholtz entry points above;
// out[Q][4] = sum(dK/d(var, lt, la, lb)_q * dL_dK), partial holds 4 Q grad_sums_block_partials doubles
cudaError_t hsum_build_block_layout(const double* X, int N, const double* X2, int M, const HsumParams& hp,
                                    double diag_add, double* K, long ldk, cudaStream_t st);
cudaError_t hsum_build_interleaved_lower(const double* X, int N, const HsumParams& hp, double diag_add,
                                         double* K, long ldk, int npad, cudaStream_t st);
cudaError_t hsum_grad_sums(const double* X, int N, const double* X2, int M, const HsumParams& hp, const double* dL_dK,
                           long ld, double* partial, size_t partial_doubles, double* out, cudaStream_t st);

// predict.cu ------------------------------------------------------------------------------
// scratch: one [npad x 128] K* panel per resident CTA (at most one CTA per SM is launched;
// fewer panels only reduce the grid).
cudaError_t predict_fused(const double* Zt, int npad, const double* alpha_int,
                          const double* X, int N, const HelmParams& hp, const double* Xs, int M,
                          long out_stride, double var_add, double* mean, double* var,
                          double* scratch, size_t scratch_bytes, cudaStream_t st, const int* gate = nullptr);
size_t predict_panel_bytes(int npad);
size_t predict_scratch_bytes(int npad, int M, int pts_per_tile);   // full parallelism on the current device
// scalar ARD-RBF sum: alpha and the tiles are in plain observation order (npad = N rounded up to 128)
cudaError_t predict_fused_rbf(const double* Zt, int npad, const double* alpha, const double* X, int N,
                              const RbfParams& rp, const double* Xs, int M, double var_add, double* mean,
                              double* var, double* scratch, size_t scratch_bytes, cudaStream_t st);
// sum of space-time Helmholtz terms: same tiles and ordering as predict_fused
cudaError_t predict_fused_hsum(const double* Zt, int npad, const double* alpha_int, const double* X, int N,
                               const HsumParams& sp, const double* Xs, int M, long out_stride, double var_add,
                               double* mean, double* var, double* scratch, size_t scratch_bytes, cudaStream_t st);
int predict_max_ctas();
void set_predict_split(int s);          // bring-up override: 1, 2, 4, 8 (0 = heuristic)

// predict_i8.cu ---------------------------------------------------------------------------
// int8-sliced (tcgen05, TMEM) variant of predict_fused for the Helmholtz family.  Fit side: digit slices of the
// lower triangle of Z as shared-memory tile images (Zq, i8_zq_bytes) with one power-of-two unit per row (zunit);
// zqs is npad doubles of scratch.  *gate holds the slice count chosen for this fit (6, 7, or 0 = fp64 path only).
size_t i8_zq_bytes(int npad);
int i8_max_npad();
cudaError_t i8_quantize_lower(const double* Z, long ldz, int npad, double* zunit, double* zqs, int8_t* Zq, cudaStream_t st,
                              int batch = 1, long bstride = 0);
cudaError_t i8_set_gate(int* gate, int s, cudaStream_t st);
size_t predict_i8_scratch_bytes(int npad);
size_t predict_i8_min_scratch_bytes(int npad);      // below this the fp64 kernel takes the call
void set_i8_debug(int v);                // bring-up timing experiments, per host thread
// launches the kernel of every slice count (only_s = 0) or of one; a kernel whose count is not *gate returns at once
cudaError_t predict_fused_i8(const int8_t* Zq, const double* zunit, const int* gate, int npad, const double* alpha_int,
                             const double* X, int N, const HelmParams& hp, const double* Xs, int M, long out_stride,
                             double var_add, double* mean, double* var, void* scratch, size_t scratch_bytes, cudaStream_t st,
                             int only_s = 0);

// grad.cu ---------------------------------------------------------------------------------
// From Kinv (lower, interleaved, padded) and alpha_int: out6 = d LML / d(l_df, l_cf, ratio, tvar, lt, noise).
cudaError_t lml_grad_reduce(const double* Kinv, long ld, int npad, const double* alpha_int,
                            const double* X, int N, const HelmParams& hp, int compat,
                            double* partial, double* out6, cudaStream_t st);
// batch: every pointer + b bstride doubles, parameters from par + b bstride
cudaError_t lml_grad_reduce_batched(const double* Kinv, long ld, int npad, const double* alpha_int, const double* X, int N,
                                    const BatchPar* par, int compat, double* partial, double* out6, int batch, long bstride,
                                    cudaStream_t st);
int lml_grad_partials(int npad);
// scalar ARD-RBF sum: out[Q (1 + D) + 1] = d LML / d(var_q, l_{q,0..D-1})_q, then d LML / d noise
int rbf_lml_grad_partial_doubles(int npad);
cudaError_t rbf_lml_grad_reduce(const double* Kinv, long ld, int npad, const double* alpha, const double* X, int N,
                                const RbfParams& rp, double* partial, double* out, cudaStream_t st);

// sum of space-time Helmholtz terms: out[4 Q + 1] = d LML / d(var, lt, la, lb)_q, then d LML / d noise
size_t hsum_lml_grad_partial_doubles(int npad, int Q);
cudaError_t hsum_lml_grad_reduce(const double* Kinv, long ld, int npad, const double* alpha_int, const double* X, int N,
                                 const HsumParams& hp, double* partial, double* out, cudaStream_t st);

}  // namespace gp2d
