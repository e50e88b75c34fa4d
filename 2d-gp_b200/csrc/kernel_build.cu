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

namespace gp2d {

// ---------------------------------------------------------------------------------------
// reference block layout
// CTA: 64 x 4 threads; each thread 2 adjacent columns x 8 rows of point pairs.
// ---------------------------------------------------------------------------------------
constexpr int BB_J = 128, BB_I = 32;

template <bool VEC>
__global__ void __launch_bounds__(256)
build_block_kernel(const double* __restrict__ X, int N, const double* __restrict__ X2, int M,
                   HelmParams hp, double diag_add, int symmetric, double* __restrict__ K, long ldk) {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j0 = blockIdx.x * BB_J + tx * 2;
    const int i0 = blockIdx.y * BB_I + ty * 8;
    if (j0 >= M) return;
    const bool has1 = (j0 + 1 < M);
    const double bx0 = X2[2 * (long)j0], by0 = X2[2 * (long)j0 + 1];
    const double bx1 = has1 ? X2[2 * (long)j0 + 2] : bx0, by1 = has1 ? X2[2 * (long)j0 + 3] : by0;
#pragma unroll 2
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + r;
        if (i >= N) break;
        const double ax = X[2 * (long)i], ay = X[2 * (long)i + 1];
        double a11, a12, a22, b11, b12, b22;
        helm_block(hp, ax - bx0, ay - by0, a11, a12, a22);
        helm_block(hp, ax - bx1, ay - by1, b11, b12, b22);
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

cudaError_t build_block_layout(const double* X, int N, const double* X2, int M, const HelmParams& hp,
                               double diag_add, double* K, long ldk, cudaStream_t st) {
    if (N <= 0 || M <= 0) return cudaSuccess;
    const int symmetric = (X2 == nullptr);
    if (symmetric) X2 = X;
    dim3 grid((M + BB_J - 1) / BB_J, (N + BB_I - 1) / BB_I);
    const bool vec = (M % 2 == 0) && (ldk % 2 == 0) && ((reinterpret_cast<uintptr_t>(K) & 15) == 0);
    if (vec) build_block_kernel<true><<<grid, 256, 0, st>>>(X, N, X2, M, hp, diag_add, symmetric, K, ldk);
    else build_block_kernel<false><<<grid, 256, 0, st>>>(X, N, X2, M, hp, diag_add, symmetric, K, ldk);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// internal layout: pair-interleaved, padded, lower tiles.  One CTA per 128x128 tile
// (64x64 point pairs); a warp writes 512 contiguous bytes per matrix row.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_interleaved_kernel(const double* __restrict__ X, int N, HelmParams hp, double diag_add,
                         double* __restrict__ K, long ldk) {
    int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(I + 1) * (I + 2) / 2 <= t) ++I;
    while ((long)I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const int jj = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = J * 64 + jj;
    const bool jv = j < N;
    const double bx = jv ? X[2 * (long)j] : 0.0, by = jv ? X[2 * (long)j + 1] : 0.0;
#pragma unroll 4
    for (int ii = ty; ii < 64; ii += 4) {
        const int i = I * 64 + ii;
        double k11, k12, k22;
        if (jv && i < N) {
            helm_block(hp, X[2 * (long)i] - bx, X[2 * (long)i + 1] - by, k11, k12, k22);
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

cudaError_t build_interleaved_lower(const double* X, int N, const HelmParams& hp, double diag_add,
                                    double* K, long ldk, int npad, cudaStream_t st) {
    const int T = npad / TILE;
    build_interleaved_kernel<<<T * (T + 1) / 2, 256, 0, st>>>(X, N, hp, diag_add, K, ldk);
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
    __shared__ double sh[3 * 32];
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = blockIdx.x * GS_J + tx;
    double acc[3] = {0.0, 0.0, 0.0};
    if (j < M) {
        const double bx = X2[2 * (long)j], by = X2[2 * (long)j + 1];
        for (int r = 0; r < 8; ++r) {
            const int i = blockIdx.y * GS_I + ty * 8 + r;
            if (i >= N) break;
            double g[3][3];
            helm_block_grad(hp, compat, X[2 * (long)i] - bx, X[2 * (long)i + 1] - by, g);
            const double* w0 = W + (long)i * ld + j;
            const double* w1 = W + ((long)N + i) * ld + j;
            const double w11 = w0[0], w12 = w0[M] + w1[0], w22 = w1[M];
#pragma unroll
            for (int p = 0; p < 3; ++p) acc[p] += g[p][0] * w11 + g[p][1] * w12 + g[p][2] * w22;
        }
    }
    block_reduce<3>(acc, sh);
    if (threadIdx.x == 0) {
        double* o = partial + 3 * ((long)blockIdx.y * gridDim.x + blockIdx.x);
        o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
    }
}

int grad_sums_block_partials(int N, int M) {
    return ((M + GS_J - 1) / GS_J) * ((N + GS_I - 1) / GS_I);
}

cudaError_t kernel_grad_sums_block(const double* X, int N, const double* X2, int M, const HelmParams& hp,
                                   int compat, const double* dL_dK, long ld, double* partial,
                                   int partial_cap, double* out3, cudaStream_t st) {
    if (X2 == nullptr) X2 = X;
    dim3 grid((M + GS_J - 1) / GS_J, (N + GS_I - 1) / GS_I);
    int count = grid.x * grid.y;
    if (count > partial_cap) return cudaErrorInvalidValue;
    grad_sums_block_kernel<<<grid, 256, 0, st>>>(X, N, X2, M, hp, compat, dL_dK, ld, partial);
    final_reduce_kernel<3><<<1, 1024, 0, st>>>(partial, count, out3);
    return cudaGetLastError();
}

}  // namespace gp2d
