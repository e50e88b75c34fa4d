// Log-marginal-likelihood gradient: d LML / d(l_df, l_cf, ratio, noise)
//   = sum_ij W_ij dK_ij/dtheta,  W = 0.5 (alpha alpha^T - K^-1)     (noise: trace W)
// Replaces GPy's dL_dK + kern.update_gradients_full (myKernel.py:59-106) +
// likelihood.update_gradients, i.e. sklearn _gpr.py:629-654.  K^-1 = Z^T Z comes from the
// DMMA GEMM (lower tiles); dK/dtheta is regenerated on the fly from the point pairs, so
// the only HBM traffic is one read of the lower triangle of K^-1.
#include "common.cuh"
#include "dgemm.cuh"
#include "linalg.h"
#include "reduce.cuh"
#include "hsum.cuh"

namespace gp2d {

__device__ __forceinline__ void lml_grad_tile(const double* __restrict__ Kinv, long ld, const double* __restrict__ alpha,
                                              const double* __restrict__ X, int N, const HelmParams& hp, int compat,
                                              double* __restrict__ partial) {
    __shared__ double sh[(HELM_NP + 1) * 32];
    int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(I + 1) * (I + 2) / 2 <= t) ++I;
    while ((long)I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const int jj = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = J * 64 + jj;
    double acc[HELM_NP + 1] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (j < N) {
        const HelmPoint qb = helm_point(hp, X, j);
        const double b0 = alpha[2 * j], b1 = alpha[2 * j + 1];
        for (int ii = ty; ii < 64; ii += 4) {
            const int i = I * 64 + ii;
            if (i >= N || i < j) continue;
            const double* r0 = Kinv + (long)(2 * i) * ld + 2 * j;
            const double2 q0 = *reinterpret_cast<const double2*>(r0);
            const double2 q1 = *reinterpret_cast<const double2*>(r0 + ld);
            const double a0 = alpha[2 * i], a1 = alpha[2 * i + 1];
            const double W11 = 0.5 * (a0 * b0 - q0.x), W12 = 0.5 * (a0 * b1 - q0.y);
            const double W21 = 0.5 * (a1 * b0 - q1.x), W22 = 0.5 * (a1 * b1 - q1.y);
            double g[HELM_NP][3];
            const HelmPoint pa = helm_point(hp, X, i);
            helm_block_grad(hp, compat, pa.a - qb.a, pa.b - qb.b, pa.t - qb.t, g);
            const double wgt = (i == j) ? 1.0 : 2.0;
            const double Ws = W12 + W21;
#pragma unroll
            for (int p = 0; p < HELM_NP; ++p)
                acc[p] += wgt * (g[p][0] * W11 + g[p][1] * Ws + g[p][2] * W22);
            if (i == j) acc[HELM_NP] += W11 + W22;
        }
    }
    block_reduce<HELM_NP + 1>(acc, sh);
    if (threadIdx.x == 0) {
        double* o = partial + (HELM_NP + 1) * (long)blockIdx.x;
#pragma unroll
        for (int p = 0; p <= HELM_NP; ++p) o[p] = acc[p];
    }
}

__global__ void __launch_bounds__(256)
lml_grad_kernel(const double* __restrict__ Kinv, long ld, const double* __restrict__ alpha,
                const double* __restrict__ X, int N, HelmParams hp, int compat,
                double* __restrict__ partial) {
    lml_grad_tile(Kinv, ld, alpha, X, N, hp, compat, partial);
}

// batch (gridDim.y problems bstride doubles apart), parameters from the problem's BatchPar
__global__ void __launch_bounds__(256)
lml_grad_batched_kernel(const double* __restrict__ Kinv, long ld, const double* __restrict__ alpha,
                        const double* __restrict__ X, int N, const BatchPar* __restrict__ par, int compat,
                        double* __restrict__ partial, long bstride) {
    const long o = (long)blockIdx.y * bstride;
    const HelmParams hp = reinterpret_cast<const BatchPar*>(reinterpret_cast<const double*>(par) + o)->hp;
    lml_grad_tile(Kinv + o, ld, alpha + o, X + o, N, hp, compat, partial + o);
}

int lml_grad_partials(int npad) {
    int T = npad / TILE;
    return T * (T + 1) / 2;
}

// out6 = d LML / d(l_df, l_cf, ratio, tvar, lt, noise); partial holds (HELM_NP + 1) * lml_grad_partials doubles
cudaError_t lml_grad_reduce(const double* Kinv, long ld, int npad, const double* alpha_int,
                            const double* X, int N, const HelmParams& hp, int compat,
                            double* partial, double* out6, cudaStream_t st) {
    int count = lml_grad_partials(npad);
    lml_grad_kernel<<<count, 256, 0, st>>>(Kinv, ld, alpha_int, X, N, hp, compat, partial);
    final_reduce_kernel<HELM_NP + 1><<<1, 1024, 0, st>>>(partial, count, out6);
    return cudaGetLastError();
}

cudaError_t lml_grad_reduce_batched(const double* Kinv, long ld, int npad, const double* alpha_int, const double* X, int N,
                                    const BatchPar* par, int compat, double* partial, double* out6, int batch, long bstride,
                                    cudaStream_t st) {
    int count = lml_grad_partials(npad);
    lml_grad_batched_kernel<<<dim3(count, batch), 256, 0, st>>>(Kinv, ld, alpha_int, X, N, par, compat, partial, bstride);
    final_reduce_kernel<HELM_NP + 1><<<batch, 1024, 0, st>>>(partial, count, out6, bstride);
    return cudaGetLastError();
}

// ---- scalar ARD-RBF sum ----------------------------------------------------------------------
// out[Q (1 + D) + 1] = d LML / d(var_q, l_{q,d})_q, d LML / d noise, from Kinv (lower tiles, plain
// observation order, padded) and alpha.  One CTA per 128 x 128 lower tile.
__global__ void __launch_bounds__(256)
rbf_lml_grad_kernel(const double* __restrict__ Kinv, long ld, const double* __restrict__ alpha,
                    const double* __restrict__ X, int N, RbfParams rp, double* __restrict__ partial) {
    __shared__ double sh[(RBF_NG + 1) * 32];
    int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(I + 1) * (I + 2) / 2 <= t) ++I;
    while ((long)I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const int j = J * TILE + (threadIdx.x & 127), half = threadIdx.x >> 7;
    double acc[RBF_NG + 1];
#pragma unroll
    for (int q = 0; q <= RBF_NG; ++q) acc[q] = 0.0;
    if (j < N) {
        double b[RBF_MAXD];
        rbf_load_point(X, j, rp.D, b);
        const double aj = alpha[j];
        for (int ii = half; ii < TILE; ii += 2) {
            const int i = I * TILE + ii;
            if (i >= N || i < j) continue;
            double a[RBF_MAXD];
            rbf_load_point(X, i, rp.D, a);
            const double W = 0.5 * (alpha[i] * aj - Kinv[(long)i * ld + j]);
            double g[RBF_NG];
#pragma unroll
            for (int q = 0; q < RBF_NG; ++q) g[q] = 0.0;
            rbf_grad_terms(rp, a, b, (i == j) ? W : 2.0 * W, g);
#pragma unroll
            for (int q = 0; q < RBF_NG; ++q) acc[q] += g[q];
            if (i == j) acc[RBF_NG] += W;
        }
    }
    block_reduce<RBF_NG + 1>(acc, sh);
    if (threadIdx.x == 0) {
        double* o = partial + (RBF_NG + 1) * (long)blockIdx.x;
#pragma unroll
        for (int q = 0; q <= RBF_NG; ++q) o[q] = acc[q];
    }
}

__global__ void rbf_lml_grad_compact_kernel(const double* __restrict__ full, int Q, int D, double* __restrict__ out) {
    const int t = threadIdx.x, np = Q * (1 + D);
    if (t < np) {
        const int q = t / (1 + D), r = t % (1 + D);
        out[t] = full[q * (1 + RBF_MAXD) + r];
    } else if (t == np) {
        out[np] = full[RBF_NG];
    }
}

int rbf_lml_grad_partial_doubles(int npad) {
    const int T = npad / TILE;
    return (RBF_NG + 1) * (T * (T + 1) / 2 + 1);
}

cudaError_t rbf_lml_grad_reduce(const double* Kinv, long ld, int npad, const double* alpha, const double* X, int N,
                                const RbfParams& rp, double* partial, double* out, cudaStream_t st) {
    const int T = npad / TILE, count = T * (T + 1) / 2;
    double* full = partial + (size_t)(RBF_NG + 1) * count;
    rbf_lml_grad_kernel<<<count, 256, 0, st>>>(Kinv, ld, alpha, X, N, rp, partial);
    final_reduce_kernel<RBF_NG + 1><<<1, 1024, 0, st>>>(partial, count, full);
    rbf_lml_grad_compact_kernel<<<1, 32, 0, st>>>(full, rp.Q, rp.D, out);
    return cudaGetLastError();
}

// ---- sum of space-time Helmholtz terms (hsum.cuh) ---------------------------------------------
// out[4 Q + 1] = d LML / d(var, lt, la, lb)_q, then d LML / d noise.  grid (lower tiles, Q): one
// term per y-slice (the lower triangle of K^-1 is read once per term; it is L2-resident for the
// sizes where this matters and the reduction is a small part of an evaluation).
__global__ void __launch_bounds__(256)
hsum_lml_grad_kernel(const double* __restrict__ Kinv, long ld, const double* __restrict__ alpha,
                     const double* __restrict__ X, int N, const __grid_constant__ HsumParams hp,
                     double* __restrict__ partial) {
    __shared__ double sh[(HSUM_NP + 1) * 32];
    int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(I + 1) * (I + 2) / 2 <= t) ++I;
    while ((long)I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const int jj = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = J * 64 + jj;
    const HsumTerm term = hp.t[blockIdx.y];
    double acc4[HSUM_NP] = {0.0, 0.0, 0.0, 0.0};
    double tr = 0.0;
    if (j < N) {
        const HelmPoint qb = hsum_point(hp, X, j);
        const double b0 = alpha[2 * j], b1 = alpha[2 * j + 1];
        for (int ii = ty; ii < 64; ii += 4) {
            const int i = I * 64 + ii;
            if (i >= N || i < j) continue;
            const double* r0 = Kinv + (long)(2 * i) * ld + 2 * j;
            const double2 q0 = *reinterpret_cast<const double2*>(r0);
            const double2 q1 = *reinterpret_cast<const double2*>(r0 + ld);
            const double a0 = alpha[2 * i], a1 = alpha[2 * i + 1];
            const double wgt = (i == j) ? 0.5 : 1.0;          // 0.5 (alpha alpha^T - K^-1), off-diagonal pairs twice
            const double W11 = wgt * (a0 * b0 - q0.x), W12 = wgt * (a0 * b1 - q0.y);
            const double W21 = wgt * (a1 * b0 - q1.x), W22 = wgt * (a1 * b1 - q1.y);
            const HelmPoint pa = hsum_point(hp, X, i);
            const double d1 = pa.a - qb.a, d2 = pa.b - qb.b, dt = pa.t - qb.t;
            hsum_term_grad(term, dt * dt, d1 * d1, d2 * d2, d1 * d2, W11, W12 + W21, W22, acc4);
            if (i == j) tr += W11 + W22;
        }
    }
    double acc[HSUM_NP + 1] = {acc4[0], acc4[1], acc4[2], acc4[3], tr};
    block_reduce<HSUM_NP + 1>(acc, sh);
    if (threadIdx.x == 0) {
        double* o = partial + (HSUM_NP + 1) * ((long)blockIdx.y * gridDim.x + blockIdx.x);
#pragma unroll
        for (int p = 0; p <= HSUM_NP; ++p) o[p] = acc[p];
    }
}

__global__ void hsum_lml_grad_compact_kernel(const double* __restrict__ full, int Q, double* __restrict__ out) {
    const int t = threadIdx.x;
    if (t < HSUM_NP * Q) out[t] = full[(t / HSUM_NP) * (HSUM_NP + 1) + t % HSUM_NP];
    else if (t == HSUM_NP * Q) out[t] = full[HSUM_NP];      // trace of W from the first term's slice
}

size_t hsum_lml_grad_partial_doubles(int npad, int Q) {
    const int T = npad / TILE;
    return (size_t)(HSUM_NP + 1) * ((size_t)Q * (T * (T + 1) / 2) + HSUM_MAXQ);
}

cudaError_t hsum_lml_grad_reduce(const double* Kinv, long ld, int npad, const double* alpha_int, const double* X, int N,
                                 const HsumParams& hp, double* partial, double* out, cudaStream_t st) {
    const int T = npad / TILE, count = T * (T + 1) / 2;
    double* full = partial + (size_t)(HSUM_NP + 1) * count * hp.Q;
    hsum_lml_grad_kernel<<<dim3(count, hp.Q), 256, 0, st>>>(Kinv, ld, alpha_int, X, N, hp, partial);
    strided_final_reduce_kernel<HSUM_NP + 1><<<hp.Q, 1024, 0, st>>>(partial, count, full, HSUM_NP + 1);
    hsum_lml_grad_compact_kernel<<<1, 64, 0, st>>>(full, hp.Q, out);
    return cudaGetLastError();
}

}  // namespace gp2d
