// Internal host-side entry points shared between the translation units of libgp2d.
#pragma once
#include <cuda_runtime.h>
#include "helmholtz.cuh"

namespace gp2d {

// linalg.cu -----------------------------------------------------------------------------
// Lower Cholesky of the row-major SPD matrix A (n multiple of 128) and, when need_inv,
// Z = L^-1 (lower; upper off-diagonal tiles untouched).  A is destroyed unless keep_L
// (then its lower triangle holds L and W must be an (n/2)^2 scratch).  logdiag[n] receives
// log(L_ii); *info the 1-based index of the first non-positive pivot, else 0.
cudaError_t potri_lower(double* A, long lda, double* Z, long ldz, int n, double* logdiag, int* info,
                        bool need_inv, bool keep_L, double* W, cudaStream_t st);
// alpha = Z^T Z y and LML from Z, logdiag.  y_block is the caller's [u;v] vector (2N);
// everything else is internal (interleaved, padded to npad).
cudaError_t solve_alpha_lml(const double* Z, long ldz, int npad, int N, const double* y_block,
                            double* y_int, double* w, double* alpha_int, double* partial,
                            const double* logdiag, double* lml_out, cudaStream_t st);
cudaError_t deinterleave(const double* xi, int N, double* x, cudaStream_t st);
// Tile-major, pre-swizzled copy of the lower triangle of Z (what predict_fused streams with
// bulk copies): tile (row block li, k-tile kt) is number 8 li (li+1)/2 + kt, 2048 doubles each.
size_t packed_tiles_doubles(int npad);
cudaError_t pack_lower_tiles(const double* Z, long ldz, int npad, double* Zt, cudaStream_t st);

// kernel_build.cu -------------------------------------------------------------------------
// Reference (component-major block) layout, arbitrary N, M, ld: K[2N, 2M].  X2 == nullptr
// means X2 = X.  diag_add is added where row == col (only meaningful for X2 == X).
cudaError_t build_block_layout(const double* X, int N, const double* X2, int M, const HelmParams& hp,
                               double diag_add, double* K, long ldk, cudaStream_t st);
// Internal layout: pair-interleaved (row 2i+c), padded to npad with identity, lower
// 128x128 tiles only (diagonal tiles complete).
cudaError_t build_interleaved_lower(const double* X, int N, const HelmParams& hp, double diag_add,
                                    double* K, long ldk, int npad, cudaStream_t st);
// sum(dK/dtheta * dL_dK) for theta = (l_df, l_cf, ratio); dL_dK in block layout [2N,2M].
cudaError_t kernel_grad_sums_block(const double* X, int N, const double* X2, int M, const HelmParams& hp,
                                   int compat, const double* dL_dK, long ld, double* partial,
                                   int partial_cap, double* out3, cudaStream_t st);
int grad_sums_block_partials(int N, int M);

// predict.cu ------------------------------------------------------------------------------
// scratch: one [npad x 128] K* panel per resident CTA (at most one CTA per SM is launched;
// fewer panels only reduce the grid).
cudaError_t predict_fused(const double* Zt, int npad, const double* alpha_int,
                          const double* X, int N, const HelmParams& hp, const double* Xs, int M,
                          long out_stride, double var_add, double* mean, double* var,
                          double* scratch, size_t scratch_bytes, cudaStream_t st);
size_t predict_panel_bytes(int npad);
size_t predict_scratch_bytes(int npad, int M);      // full parallelism on the current device
int predict_max_ctas();

// grad.cu ---------------------------------------------------------------------------------
// From Kinv (lower, interleaved, padded) and alpha_int: out4 = d LML / d(l_df, l_cf, ratio, noise).
cudaError_t lml_grad_reduce(const double* Kinv, long ld, int npad, const double* alpha_int,
                            const double* X, int N, const HelmParams& hp, int compat,
                            double* partial, double* out4, cudaStream_t st);
int lml_grad_partials(int npad);

}  // namespace gp2d
