// Covariance build: the 2N x 2M matrix-valued Helmholtz kernel, and the plug-in gradient
// reductions sum(dK/dtheta * dL_dK).
//
// Replaces myKernel.myKernel.K / nonDivK.K / nonRotK.K (myKernel.py:27-53,159-176,255-271),
// GP_scripts.myKernel / compute_K / compute_Ks (GP_scripts.py:6-42,74-123) and
// myKernel.update_gradients_full (myKernel.py:59-106).  HBM-write bound: one thread per
// point pair evaluates the 2x2 block once (1-2 exps) and issues coalesced stores into the
// four quadrants.
#include "common.cuh"
#include "linalg.h"
#include "reduce.cuh"
#include "rbf.cuh"
#include "hsum.cuh"

namespace gp2d {

// the two matrix-valued families share the build kernels: points and 2x2 blocks by overload
__device__ __forceinline__ HelmPoint kp_point(const HelmParams& p, const double* __restrict__ X, long i) { return helm_point(p, X, i); }
__device__ __forceinline__ HelmPoint kp_point(const HsumParams& p, const double* __restrict__ X, long i) { return hsum_point(p, X, i); }
__device__ __forceinline__ void kp_block(const HelmParams& p, const HelmPoint& x, const HelmPoint& y, double& k11, double& k12, double& k22) {
    helm_block_pts(p, x, y, k11, k12, k22);
}
__device__ __forceinline__ void kp_block(const HsumParams& p, const HelmPoint& x, const HelmPoint& y, double& k11, double& k12, double& k22) {
    hsum_block_pts(p, x, y, k11, k12, k22);
}

// ---------------------------------------------------------------------------------------
// reference block layout
// CTA: 64 x 4 threads; each thread 2 adjacent columns x 8 rows of point pairs.
// ---------------------------------------------------------------------------------------
constexpr int BB_J = 128, BB_I = 32;

template <bool VEC, class KP>
__global__ void __launch_bounds__(256)
build_block_kernel(const double* __restrict__ X, int N, const double* __restrict__ X2, int M,
                   const __grid_constant__ KP hp, double diag_add, int symmetric, double* __restrict__ K, long ldk) {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j0 = blockIdx.x * BB_J + tx * 2;
    const int i0 = blockIdx.y * BB_I + ty * 8;
    if (j0 >= M) return;
    const bool has1 = (j0 + 1 < M);
    const HelmPoint q0 = kp_point(hp, X2, j0);
    const HelmPoint q1 = has1 ? kp_point(hp, X2, j0 + 1) : q0;
#pragma unroll 2
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + r;
        if (i >= N) break;
        const HelmPoint pa = kp_point(hp, X, i);
        double a11, a12, a22, b11, b12, b22;
        kp_block(hp, pa, q0, a11, a12, a22);
        kp_block(hp, pa, q1, b11, b12, b22);
        if (symmetric) {
            if (i == j0) { a11 += diag_add; a22 += diag_add; }
            if (i == j0 + 1) { b11 += diag_add; b22 += diag_add; }
        }
        double* r0 = K + (long)i * ldk + j0;          // component 0 row
        double* r1 = K + ((long)N + i) * ldk + j0;    // component 1 row
        if (VEC) {
            *reinterpret_cast<double2*>(r0) = make_double2(a11, b11);
            *reinterpret_cast<double2*>(r0 + M) = make_double2(a12, b12);
            *reinterpret_cast<double2*>(r1) = make_double2(a12, b12);
            *reinterpret_cast<double2*>(r1 + M) = make_double2(a22, b22);
        } else {
            r0[0] = a11; r0[M] = a12; r1[0] = a12; r1[M] = a22;
            if (has1) { r0[1] = b11; r0[M + 1] = b12; r1[1] = b12; r1[M + 1] = b22; }
        }
    }
}

template <class KP>
static cudaError_t build_block_layout_t(const double* X, int N, const double* X2, int M, const KP& hp,
                                        double diag_add, double* K, long ldk, cudaStream_t st) {
    if (N <= 0 || M <= 0) return cudaSuccess;
    const int symmetric = (X2 == nullptr);
    if (symmetric) X2 = X;
    dim3 grid((M + BB_J - 1) / BB_J, (N + BB_I - 1) / BB_I);
    const bool vec = (M % 2 == 0) && (ldk % 2 == 0) && ((reinterpret_cast<uintptr_t>(K) & 15) == 0);
    if (vec) build_block_kernel<true, KP><<<grid, 256, 0, st>>>(X, N, X2, M, hp, diag_add, symmetric, K, ldk);
    else build_block_kernel<false, KP><<<grid, 256, 0, st>>>(X, N, X2, M, hp, diag_add, symmetric, K, ldk);
    return cudaGetLastError();
}

cudaError_t build_block_layout(const double* X, int N, const double* X2, int M, const HelmParams& hp,
                               double diag_add, double* K, long ldk, cudaStream_t st) {
    return build_block_layout_t(X, N, X2, M, hp, diag_add, K, ldk, st);
}
cudaError_t hsum_build_block_layout(const double* X, int N, const double* X2, int M, const HsumParams& hp,
                                    double diag_add, double* K, long ldk, cudaStream_t st) {
    return build_block_layout_t(X, N, X2, M, hp, diag_add, K, ldk, st);
}

// ---------------------------------------------------------------------------------------
// internal layout: pair-interleaved, padded, lower tiles.  One CTA per 128x128 tile
// (64x64 point pairs); a warp writes 512 contiguous bytes per matrix row.
// ---------------------------------------------------------------------------------------
template <class KP>
__device__ __forceinline__ void build_interleaved_tile(const double* __restrict__ X, int N, const KP& hp, double diag_add,
                                                       double* __restrict__ K, long ldk) {
    int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(I + 1) * (I + 2) / 2 <= t) ++I;
    while ((long)I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const int jj = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = J * 64 + jj;
    const bool jv = j < N;
    const HelmPoint qb = kp_point(hp, X, jv ? j : 0);
#pragma unroll 4
    for (int ii = ty; ii < 64; ii += 4) {
        const int i = I * 64 + ii;
        double k11, k12, k22;
        if (jv && i < N) {
            kp_block(hp, kp_point(hp, X, i), qb, k11, k12, k22);
            if (i == j) { k11 += diag_add; k22 += diag_add; }
        } else {
            k12 = 0.0;
            k11 = k22 = (i == j) ? 1.0 : 0.0;
        }
        double* r0 = K + (long)(2 * i) * ldk + 2 * j;
        *reinterpret_cast<double2*>(r0) = make_double2(k11, k12);
        *reinterpret_cast<double2*>(r0 + ldk) = make_double2(k12, k22);
    }
}

template <class KP>
__global__ void __launch_bounds__(256)
build_interleaved_kernel(const double* __restrict__ X, int N, const __grid_constant__ KP hp, double diag_add,
                         double* __restrict__ K, long ldk) {
    build_interleaved_tile(X, N, hp, diag_add, K, ldk);
}

// batch (gridDim.y problems): the same tile code with the problem's own points, parameters and matrix
__global__ void __launch_bounds__(256)
build_interleaved_batched_kernel(const double* __restrict__ X, long x_bstride, int N, const BatchPar* __restrict__ par,
                                 double* __restrict__ K, long ldk, long bstride) {
    const BatchPar* bp = reinterpret_cast<const BatchPar*>(reinterpret_cast<const double*>(par) + (long)blockIdx.y * bstride);
    const HelmParams hp = bp->hp;
    build_interleaved_tile(X + (long)blockIdx.y * x_bstride, N, hp, bp->diag_add, K + (long)blockIdx.y * bstride, ldk);
}

cudaError_t build_interleaved_lower(const double* X, int N, const HelmParams& hp, double diag_add,
                                    double* K, long ldk, int npad, cudaStream_t st) {
    const int T = npad / TILE;
    build_interleaved_kernel<HelmParams><<<T * (T + 1) / 2, 256, 0, st>>>(X, N, hp, diag_add, K, ldk);
    return cudaGetLastError();
}
cudaError_t build_interleaved_lower_batched(const double* X, long x_bstride, int N, const BatchPar* par, double* K, long ldk,
                                            int npad, int batch, long bstride, cudaStream_t st) {
    const int T = npad / TILE;
    build_interleaved_batched_kernel<<<dim3(T * (T + 1) / 2, batch), 256, 0, st>>>(X, x_bstride, N, par, K, ldk, bstride);
    return cudaGetLastError();
}

// Parameters of up to 16 problems travel as kernel arguments (no host staging buffer to keep alive, no
// stream synchronisation as a pageable cudaMemcpyAsync would force) and are written to each problem's
// workspace together with its copy of the observation points.
constexpr int SCATTER_CHUNK = 16;
struct BatchParChunk { BatchPar p[SCATTER_CHUNK]; };
__global__ void __launch_bounds__(256)
scatter_batch_params_kernel(const __grid_constant__ BatchParChunk chunk, BatchPar* __restrict__ par, long bstride,
                            const double* __restrict__ X, long x_bstride, int x_doubles, double* __restrict__ Xws) {
    const int b = blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x < (int)(sizeof(BatchPar) / sizeof(double)))
        (reinterpret_cast<double*>(par) + (long)b * bstride)[threadIdx.x] = reinterpret_cast<const double*>(&chunk.p[b])[threadIdx.x];
    const double* src = X + (long)b * x_bstride;
    double* dst = Xws + (long)b * bstride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < x_doubles; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
static_assert(sizeof(BatchPar) % sizeof(double) == 0, "BatchPar is copied as doubles");

cudaError_t scatter_batch_params(const BatchPar* host, int batch, BatchPar* par, long bstride, const double* X,
                                 long x_bstride, int x_doubles, double* Xws, cudaStream_t st) {
    for (int b0 = 0; b0 < batch; b0 += SCATTER_CHUNK) {
        const int nb = batch - b0 < SCATTER_CHUNK ? batch - b0 : SCATTER_CHUNK;
        BatchParChunk c;
        for (int i = 0; i < nb; ++i) c.p[i] = host[b0 + i];
        for (int i = nb; i < SCATTER_CHUNK; ++i) c.p[i] = host[b0];
        scatter_batch_params_kernel<<<dim3(8, nb), 256, 0, st>>>(
            c, reinterpret_cast<BatchPar*>(reinterpret_cast<double*>(par) + (long)b0 * bstride), bstride,
            X + (long)b0 * x_bstride, x_bstride, x_doubles, Xws + (long)b0 * bstride);
    }
    return cudaGetLastError();
}
cudaError_t hsum_build_interleaved_lower(const double* X, int N, const HsumParams& hp, double diag_add,
                                         double* K, long ldk, int npad, cudaStream_t st) {
    const int T = npad / TILE;
    build_interleaved_kernel<HsumParams><<<T * (T + 1) / 2, 256, 0, st>>>(X, N, hp, diag_add, K, ldk);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// scalar ARD-RBF sum (rbf.cuh): K[N,M] row-major.  CTA = 128 columns x 32 rows; a thread owns 2
// adjacent columns x 8 rows.  PADDED: internal lower 128x128 tiles with identity padding.
// ---------------------------------------------------------------------------------------
template <bool PADDED>
__global__ void __launch_bounds__(256)
rbf_build_kernel(const double* __restrict__ X, int N, const double* __restrict__ X2, int M, RbfParams rp,
                 double diag_add, int symmetric, double* __restrict__ K, long ldk) {
    int bi, bj;
    if (PADDED) {
        int t = blockIdx.x;
        int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
        while ((long)(I + 1) * (I + 2) / 2 <= t) ++I;
        while ((long)I * (I + 1) / 2 > t) --I;
        bj = t - I * (I + 1) / 2;
        bi = I * 4 + blockIdx.y;             // 4 row slabs of 32 per 128-tile
    } else {
        bj = blockIdx.x;
        bi = blockIdx.y;
    }
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j0 = bj * 128 + tx * 2;
    const int i0 = bi * 32 + ty * 8;
    if (!PADDED && j0 >= M) return;
    const bool v0 = j0 < M, v1 = j0 + 1 < M;
    double b0[RBF_MAXD], b1[RBF_MAXD];
    rbf_load_point(X2, v0 ? j0 : 0, rp.D, b0);
    rbf_load_point(X2, v1 ? j0 + 1 : 0, rp.D, b1);
#pragma unroll 2
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + r;
        if (!PADDED && i >= N) break;
        double k0 = 0.0, k1 = 0.0;
        if (i < N) {
            double a[RBF_MAXD];
            rbf_load_point(X, i, rp.D, a);
            if (v0) k0 = rbf_eval(rp, a, b0);
            if (v1) k1 = rbf_eval(rp, a, b1);
            if (symmetric) {
                if (i == j0) k0 += diag_add;
                if (i == j0 + 1) k1 += diag_add;
            }
        } else {                              // padding rows: identity
            k0 = (i == j0) ? 1.0 : 0.0;
            k1 = (i == j0 + 1) ? 1.0 : 0.0;
        }
        if (PADDED) {
            *reinterpret_cast<double2*>(K + (long)i * ldk + j0) = make_double2(k0, k1);
        } else {
            double* o = K + (long)i * ldk + j0;
            o[0] = k0;
            if (v1) o[1] = k1;
        }
    }
}

cudaError_t rbf_build(const double* X, int N, const double* X2, int M, const RbfParams& rp, double diag_add,
                      double* K, long ldk, cudaStream_t st) {
    if (N <= 0 || M <= 0) return cudaSuccess;
    const int symmetric = (X2 == nullptr);
    if (symmetric) X2 = X;
    dim3 grid((M + 127) / 128, (N + 31) / 32);
    rbf_build_kernel<false><<<grid, 256, 0, st>>>(X, N, X2, M, rp, diag_add, symmetric, K, ldk);
    return cudaGetLastError();
}

cudaError_t rbf_build_padded_lower(const double* X, int N, const RbfParams& rp, double diag_add, double* K,
                                   long ldk, int npad, cudaStream_t st) {
    const int T = npad / TILE;
    rbf_build_kernel<true><<<dim3(T * (T + 1) / 2, 4), 256, 0, st>>>(X, N, X, N, rp, diag_add, 1, K, ldk);
    return cudaGetLastError();
}

// sum(dK/dtheta * dL_dK) for theta = (var_q, l_{q,0..D-1})_q : out[Q (1 + D)], dL_dK is [N, M]
//   dK/dvar_q = k_q / var_q ;  dK/dl_{q,d} = k_q * delta_d^2 / l_{q,d}^3        (GPy RBF, ARD)
__global__ void __launch_bounds__(256)
rbf_grad_sums_kernel(const double* __restrict__ X, int N, const double* __restrict__ X2, int M, RbfParams rp,
                     const double* __restrict__ W, long ld, double* __restrict__ partial) {
    __shared__ double sh[RBF_NG * 32];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = blockIdx.x * 64 + tx;
    double acc[RBF_NG];
#pragma unroll
    for (int q = 0; q < RBF_NG; ++q) acc[q] = 0.0;
    if (j < M) {
        double b[RBF_MAXD];
        rbf_load_point(X2, j, rp.D, b);
        for (int r = 0; r < 8; ++r) {
            const int i = blockIdx.y * 32 + ty * 8 + r;
            if (i >= N) break;
            double a[RBF_MAXD];
            rbf_load_point(X, i, rp.D, a);
            rbf_grad_terms(rp, a, b, W[(long)i * ld + j], acc);
        }
    }
    block_reduce<RBF_NG>(acc, sh);
    if (threadIdx.x == 0) {
        double* o = partial + RBF_NG * ((long)blockIdx.y * gridDim.x + blockIdx.x);
#pragma unroll
        for (int q = 0; q < RBF_NG; ++q) o[q] = acc[q];
    }
}

__global__ void rbf_grad_compact_kernel(const double* __restrict__ full, int Q, int D, double* __restrict__ out) {
    const int t = threadIdx.x;
    if (t < Q * (1 + D)) {
        const int q = t / (1 + D), r = t % (1 + D);
        out[t] = full[q * (1 + RBF_MAXD) + r];
    }
}

int rbf_grad_partials(int N, int M) { return ((M + 63) / 64) * ((N + 31) / 32); }

cudaError_t rbf_grad_sums(const double* X, int N, const double* X2, int M, const RbfParams& rp, const double* dL_dK,
                          long ld, double* partial, int partial_cap, double* out, cudaStream_t st) {
    if (X2 == nullptr) X2 = X;
    dim3 grid((M + 63) / 64, (N + 31) / 32);
    const int count = grid.x * grid.y;
    if (count + 1 > partial_cap) return cudaErrorInvalidValue;
    double* full = partial + (size_t)RBF_NG * count;
    rbf_grad_sums_kernel<<<grid, 256, 0, st>>>(X, N, X2, M, rp, dL_dK, ld, partial);
    final_reduce_kernel<RBF_NG><<<1, 1024, 0, st>>>(partial, count, full);
    rbf_grad_compact_kernel<<<1, 32, 0, st>>>(full, rp.Q, rp.D, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// plug-in gradient reductions (update_gradients_full): out3 = sum(dK/dtheta * dL_dK)
// ---------------------------------------------------------------------------------------
constexpr int GS_J = 64, GS_I = 32;

__global__ void __launch_bounds__(256)
grad_sums_block_kernel(const double* __restrict__ X, int N, const double* __restrict__ X2, int M,
                       HelmParams hp, int compat, const double* __restrict__ W, long ld,
                       double* __restrict__ partial) {
    __shared__ double sh[HELM_NP * 32];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = blockIdx.x * GS_J + tx;
    double acc[HELM_NP] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (j < M) {
        const HelmPoint qb = helm_point(hp, X2, j);
        for (int r = 0; r < 8; ++r) {
            const int i = blockIdx.y * GS_I + ty * 8 + r;
            if (i >= N) break;
            double g[HELM_NP][3];
            const HelmPoint pa = helm_point(hp, X, i);
            helm_block_grad(hp, compat, pa.a - qb.a, pa.b - qb.b, pa.t - qb.t, g);
            const double* w0 = W + (long)i * ld + j;
            const double* w1 = W + ((long)N + i) * ld + j;
            const double w11 = w0[0], w12 = w0[M] + w1[0], w22 = w1[M];
#pragma unroll
            for (int p = 0; p < HELM_NP; ++p) acc[p] += g[p][0] * w11 + g[p][1] * w12 + g[p][2] * w22;
        }
    }
    block_reduce<HELM_NP>(acc, sh);
    if (threadIdx.x == 0) {
        double* o = partial + HELM_NP * ((long)blockIdx.y * gridDim.x + blockIdx.x);
#pragma unroll
        for (int p = 0; p < HELM_NP; ++p) o[p] = acc[p];
    }
}

int grad_sums_block_partials(int N, int M) {
    return ((M + GS_J - 1) / GS_J) * ((N + GS_I - 1) / GS_I);
}

// out receives 3 sums (l_df, l_cf, ratio), or 5 (+ tvar, lt) when the time factor is on; partial
// holds HELM_NP * (grad_sums_block_partials + 1) doubles.
cudaError_t kernel_grad_sums_block(const double* X, int N, const double* X2, int M, const HelmParams& hp,
                                   int compat, const double* dL_dK, long ld, double* partial,
                                   int partial_cap, double* out, cudaStream_t st) {
    if (X2 == nullptr) X2 = X;
    dim3 grid((M + GS_J - 1) / GS_J, (N + GS_I - 1) / GS_I);
    int count = grid.x * grid.y;
    if (count + 1 > partial_cap) return cudaErrorInvalidValue;
    double* full = partial + (size_t)HELM_NP * count;
    grad_sums_block_kernel<<<grid, 256, 0, st>>>(X, N, X2, M, hp, compat, dL_dK, ld, partial);
    final_reduce_kernel<HELM_NP><<<1, 1024, 0, st>>>(partial, count, full);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(out, full, (hp.has_t ? 5 : 3) * sizeof(double), cudaMemcpyDeviceToDevice, st);
}

// ---------------------------------------------------------------------------------------
// sum of space-time Helmholtz terms (hsum.cuh): out[Q][4] = sum(dK/d(var, lt, la, lb)_q * dL_dK).
// grid (column blocks, row blocks, Q): one term per z-slice, so the accumulator count is fixed.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hsum_grad_sums_kernel(const double* __restrict__ X, int N, const double* __restrict__ X2, int M,
                      const __grid_constant__ HsumParams hp, const double* __restrict__ W, long ld,
                      double* __restrict__ partial) {
    __shared__ double sh[HSUM_NP * 32];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = blockIdx.x * GS_J + tx;
    const HsumTerm term = hp.t[blockIdx.z];
    double acc[HSUM_NP] = {0.0, 0.0, 0.0, 0.0};
    if (j < M) {
        const HelmPoint qb = hsum_point(hp, X2, j);
        for (int r = 0; r < 8; ++r) {
            const int i = blockIdx.y * GS_I + ty * 8 + r;
            if (i >= N) break;
            const HelmPoint pa = hsum_point(hp, X, i);
            const double d1 = pa.a - qb.a, d2 = pa.b - qb.b, dt = pa.t - qb.t;
            const double* w0 = W + (long)i * ld + j;
            const double* w1 = W + ((long)N + i) * ld + j;
            hsum_term_grad(term, dt * dt, d1 * d1, d2 * d2, d1 * d2, w0[0], w0[M] + w1[0], w1[M], acc);
        }
    }
    block_reduce<HSUM_NP>(acc, sh);
    if (threadIdx.x == 0) {
        const long nblk = (long)gridDim.x * gridDim.y;
        double* o = partial + HSUM_NP * (blockIdx.z * nblk + (long)blockIdx.y * gridDim.x + blockIdx.x);
#pragma unroll
        for (int p = 0; p < HSUM_NP; ++p) o[p] = acc[p];
    }
}

// partial holds HSUM_NP * Q * grad_sums_block_partials(N, M) doubles
cudaError_t hsum_grad_sums(const double* X, int N, const double* X2, int M, const HsumParams& hp, const double* dL_dK,
                           long ld, double* partial, size_t partial_doubles, double* out, cudaStream_t st) {
    if (X2 == nullptr) X2 = X;
    dim3 grid((M + GS_J - 1) / GS_J, (N + GS_I - 1) / GS_I, hp.Q);
    const int count = grid.x * grid.y;
    if ((size_t)HSUM_NP * count * hp.Q > partial_doubles) return cudaErrorInvalidValue;
    hsum_grad_sums_kernel<<<grid, 256, 0, st>>>(X, N, X2, M, hp, dL_dK, ld, partial);
    strided_final_reduce_kernel<HSUM_NP><<<hp.Q, 1024, 0, st>>>(partial, count, out, HSUM_NP);
    return cudaGetLastError();
}

}  // namespace gp2d
