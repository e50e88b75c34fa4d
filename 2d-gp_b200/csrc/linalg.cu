// Blocked fp64 Cholesky + triangular inverse on the DMMA pipe, and the vector kernels
// (alpha, log-likelihood) that follow it.
//
// Replaces the LAPACK calls behind the reference's fit: dpotrf/dpotri/dpotrs inside
// GPy's GPRegression (call sites GP_plots.py:763, krig.py:411), np.linalg.inv at
// GP_laser.py:118,180 and sklearn's cholesky/cho_solve (krig.py:182-185).
//
// Scheme ("potri recursion"), for a lower factorisation of the row-major SPD matrix A:
//   rec(A) :  rec(A11) -> L11 and Z11 = L11^-1
//             T   = A21 * Z11^T            (= L21; one GEMM, k clipped to the triangle)
//             A22 -= T * T^T               (SYRK, lower tiles only)
//             rec(A22) -> L22, Z22
//             Z21 = -Z22 * (T * Z11)       (two GEMMs, k clipped)
// so every flop outside the 128x128 leaves is a full-width DMMA GEMM, there is no
// latency-bound TRSM, and the by-product Z = L^-1 is exactly what the predictive pass
// (V = Z K*^T) and K^-1 = Z^T Z (gradient) need.  With need_inv=false the last step is
// skipped along the right spine (potrf-only: 8/7 of the minimal n^3/3 flops).
#include "dgemm.cuh"
#include "linalg.h"

#include <math.h>

namespace gp2d {

// ------------------------------------------------------------------------------------
// GEMM launcher
// ------------------------------------------------------------------------------------
static int g_cta_threads = 256;
void set_cta_threads(int nt) { g_cta_threads = (nt == 512) ? 512 : 256; }
int get_cta_threads() { return g_cta_threads; }

template <bool A_MN, bool B_MN, int NT>
static cudaError_t gemm_attr() {
    return cudaFuncSetAttribute(dgemm_kernel<A_MN, B_MN, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
}

cudaError_t dgemm_init() {
    static bool done = false;
    if (done) return cudaSuccess;
    cudaError_t e;
    if ((e = gemm_attr<false, false, 256>()) != cudaSuccess) return e;
    if ((e = gemm_attr<false, true, 256>()) != cudaSuccess) return e;
    if ((e = gemm_attr<true, true, 256>()) != cudaSuccess) return e;
    if ((e = gemm_attr<true, false, 256>()) != cudaSuccess) return e;
    if ((e = gemm_attr<false, false, 512>()) != cudaSuccess) return e;
    if ((e = gemm_attr<false, true, 512>()) != cudaSuccess) return e;
    if ((e = gemm_attr<true, true, 512>()) != cudaSuccess) return e;
    if ((e = gemm_attr<true, false, 512>()) != cudaSuccess) return e;
    done = true;
    return cudaSuccess;
}

template <int NT>
static void gemm_launch(bool a_mn, bool b_mn, unsigned grid, const GemmArgs& a, cudaStream_t st) {
    if (!a_mn && !b_mn) dgemm_kernel<false, false, NT><<<grid, NT, GEMM_SMEM_BYTES, st>>>(a);
    else if (!a_mn && b_mn) dgemm_kernel<false, true, NT><<<grid, NT, GEMM_SMEM_BYTES, st>>>(a);
    else if (a_mn && b_mn) dgemm_kernel<true, true, NT><<<grid, NT, GEMM_SMEM_BYTES, st>>>(a);
    else dgemm_kernel<true, false, NT><<<grid, NT, GEMM_SMEM_BYTES, st>>>(a);
}

cudaError_t launch_dgemm(bool a_mn, bool b_mn, const GemmArgs& a, cudaStream_t st) {
    if (a.M % TILE || a.N % TILE || a.K % BK || a.M <= 0 || a.N <= 0 || a.K <= 0) return cudaErrorInvalidValue;
    if ((a.lda & 1) || (a.ldb & 1) || (a.ldc & 1)) return cudaErrorInvalidValue;
    cudaError_t e = dgemm_init();
    if (e != cudaSuccess) return e;
    int tm = a.M / TILE, tn = a.N / TILE;
    if (a.lower_out && tm != tn) return cudaErrorInvalidValue;
    unsigned grid = a.lower_out ? (unsigned)((long)tm * (tm + 1) / 2) : (unsigned)(tm * tn);
    if (g_cta_threads == 512) gemm_launch<512>(a_mn, b_mn, grid, a, st);
    else gemm_launch<256>(a_mn, b_mn, grid, a, st);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// 128x128 leaf: L = chol(A_kk) written back to A (lower), Z_kk = L^-1 (full tile, zeros
// above the diagonal), log(diag L) and LAPACK-style info (1-based index of the first
// non-positive pivot, first failure wins).
// ------------------------------------------------------------------------------------
constexpr int LEAF_LD = TILE + 1;
constexpr int LEAF_SMEM_BYTES = TILE * LEAF_LD * (int)sizeof(double);

__global__ void __launch_bounds__(NTHREADS, 1)
potri_leaf_kernel(double* A, long lda, double* Z, long ldz, double* logdiag, int* info, int row0) {
    extern __shared__ double S[];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < TILE * TILE; idx += NTHREADS) {
        int r = idx >> 7, c = idx & 127;
        S[r * LEAF_LD + c] = (c <= r) ? A[(long)r * lda + c] : 0.0;
    }
    // right-looking Cholesky in shared memory
    const int ti = tid >> 4, tk = tid & 15;
    for (int j = 0; j < TILE; ++j) {
        __syncthreads();
        double d = S[j * LEAF_LD + j];
        if (!(d > 0.0)) {
            if (tid == 0) atomicCAS(info, 0, row0 + j + 1);
            d = nan("");
        }
        double sq = sqrt(d), rinv = 1.0 / sq;
        __syncthreads();
        if (tid == 0) { S[j * LEAF_LD + j] = sq; logdiag[j] = log(sq); }
        for (int i = j + 1 + tid; i < TILE; i += NTHREADS) S[i * LEAF_LD + j] *= rinv;
        __syncthreads();
        for (int i = j + 1 + ti; i < TILE; i += 16) {
            double lij = S[i * LEAF_LD + j];
            for (int k = j + 1 + tk; k <= i; k += 16)
                S[i * LEAF_LD + k] = fma(-lij, S[k * LEAF_LD + j], S[i * LEAF_LD + k]);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < TILE * TILE; idx += NTHREADS) {
        int r = idx >> 7, c = idx & 127;
        if (c <= r) A[(long)r * lda + c] = S[r * LEAF_LD + c];
    }
    // in-place inverse, columns right to left; two threads per row split the k-sum
    const int row = tid >> 1, half = tid & 1;
    for (int j = TILE - 1; j >= 0; --j) {
        __syncthreads();
        double t = 0.0;
        if (row > j)
            for (int k = j + 1 + half; k <= row; k += 2)
                t = fma(S[row * LEAF_LD + k], S[k * LEAF_LD + j], t);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        double dj = 1.0 / S[j * LEAF_LD + j];
        __syncthreads();
        if (half == 0) {
            if (row > j) S[row * LEAF_LD + j] = -t * dj;
            else if (row == j) S[j * LEAF_LD + j] = dj;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < TILE * TILE; idx += NTHREADS) {
        int r = idx >> 7, c = idx & 127;
        Z[(long)r * ldz + c] = (c <= r) ? S[r * LEAF_LD + c] : 0.0;
    }
}

__global__ void copy2d_kernel(const double* __restrict__ src, long lds, double* __restrict__ dst,
                              long ldd, int rows, int cols) {
    long total = (long)rows * (cols / 2);
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        int r = (int)(idx / (cols / 2)), c = (int)(idx % (cols / 2)) * 2;
        *reinterpret_cast<double2*>(dst + (long)r * ldd + c) =
            *reinterpret_cast<const double2*>(src + (long)r * lds + c);
    }
}

struct PotriCtx {
    double* A; long lda;
    double* Z; long ldz;
    double* logdiag; int* info;
    double* W;          // scratch (n/2 x n/2), only when keep_L
    bool keep_L;
    cudaStream_t st;
    cudaError_t err;
};

static void gemm_checked(PotriCtx& c, bool a_mn, bool b_mn, const GemmArgs& g) {
    if (c.err != cudaSuccess) return;
    c.err = launch_dgemm(a_mn, b_mn, g, c.st);
}

static void potri_rec(PotriCtx& c, int off, int n, bool need_inv) {
    if (c.err != cudaSuccess) return;
    if (n == TILE) {
        potri_leaf_kernel<<<1, NTHREADS, LEAF_SMEM_BYTES, c.st>>>(
            c.A + (long)off * (c.lda + 1), c.lda, c.Z + (long)off * (c.ldz + 1), c.ldz,
            c.logdiag + off, c.info, off);
        c.err = cudaGetLastError();
        return;
    }
    const int n1 = (n / TILE / 2) * TILE, n2 = n - n1;
    double* A21 = c.A + (long)(off + n1) * c.lda + off;
    double* A22 = c.A + (long)(off + n1) * (c.lda + 1);
    double* Z11 = c.Z + (long)off * (c.ldz + 1);
    double* Z21 = c.Z + (long)(off + n1) * c.ldz + off;
    double* Z22 = c.Z + (long)(off + n1) * (c.ldz + 1);

    potri_rec(c, off, n1, true);
    // T = A21 * Z11^T -> Z21 (scratch use of the block that will later hold Z21)
    gemm_checked(c, false, false, GemmArgs{A21, c.lda, Z11, c.ldz, Z21, c.ldz, n2, n1, n1, 1.0, 0.0, 0, KR_LE_N});
    // A22 -= T * T^T (lower tiles)
    gemm_checked(c, false, false, GemmArgs{Z21, c.ldz, Z21, c.ldz, A22, c.lda, n2, n2, n1, -1.0, 1.0, 1, KR_FULL});
    potri_rec(c, off + n1, n2, need_inv);
    if (c.err != cudaSuccess) return;
    double* U = A21; long ldu = c.lda;
    if (c.keep_L) {
        copy2d_kernel<<<296, 256, 0, c.st>>>(Z21, c.ldz, A21, c.lda, n2, n1);
        c.err = cudaGetLastError();
        U = c.W; ldu = n1;
    }
    if (need_inv) {
        // U = T * Z11 ; Z21 = -Z22 * U
        gemm_checked(c, false, true, GemmArgs{Z21, c.ldz, Z11, c.ldz, U, ldu, n2, n1, n1, 1.0, 0.0, 0, KR_GE_N});
        gemm_checked(c, false, true, GemmArgs{Z22, c.ldz, U, ldu, Z21, c.ldz, n2, n1, n2, -1.0, 0.0, 0, KR_LE_M});
    }
}

cudaError_t potri_lower(double* A, long lda, double* Z, long ldz, int n, double* logdiag, int* info,
                        bool need_inv, bool keep_L, double* W, cudaStream_t st) {
    if (n % TILE || n <= 0) return cudaErrorInvalidValue;
    static bool leaf_init = false;
    if (!leaf_init) {
        cudaError_t e = cudaFuncSetAttribute(potri_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        leaf_init = true;
    }
    cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    PotriCtx c{A, lda, Z, ldz, logdiag, info, W, keep_L, st, cudaSuccess};
    potri_rec(c, 0, n, need_inv);
    return c.err;
}

// ------------------------------------------------------------------------------------
// vectors: y -> interleaved, w = Z y, alpha = Z^T w, LML
// ------------------------------------------------------------------------------------
__global__ void interleave_kernel(const double* __restrict__ y, int N, double* __restrict__ yi, int npad) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npad) return;
    int i = p >> 1, c = p & 1;
    yi[p] = (i < N) ? y[(long)c * N + i] : 0.0;
}

__global__ void deinterleave_kernel(const double* __restrict__ xi, int N, double* __restrict__ x) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= 2 * N) return;
    int c = p / N, i = p - c * N;
    x[p] = xi[2 * i + c];
}

// w[i] = sum_{k<=i} Z[i][k] y[k]; one warp per row
__global__ void trmv_lower_kernel(const double* __restrict__ Z, long ldz, const double* __restrict__ y,
                                  double* __restrict__ w, int n) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    const double* zr = Z + (long)row * ldz;
    double s = 0.0;
    for (int k = lane; k <= row; k += 32) s = fma(zr[k], y[k], s);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) w[row] = s;
}

// partial[chunk][j] = sum_{i in chunk, i>=j} Z[i][j] w[i]; 128 columns per CTA, 256-row chunks
constexpr int TRMVT_ROWS = 256;
__global__ void trmvT_partial_kernel(const double* __restrict__ Z, long ldz, const double* __restrict__ w,
                                     double* __restrict__ partial, int n) {
    int j = blockIdx.x * 128 + threadIdx.x;
    int r0 = blockIdx.y * TRMVT_ROWS, r1 = min(n, r0 + TRMVT_ROWS);
    double s = 0.0;
    if (r1 > blockIdx.x * 128) {
        for (int i = max(r0, j); i < r1; ++i) s = fma(Z[(long)i * ldz + j], w[i], s);
    }
    partial[(long)blockIdx.y * n + j] = s;
}

__global__ void colsum_partials_kernel(const double* __restrict__ partial, int nchunks, int n,
                                       double* __restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[(long)c * n + j];
    out[j] = s;
}

// out[0] = -0.5 w'w - sum logdiag - N log(2 pi)      (n = 2N observations)
__global__ void lml_kernel(const double* __restrict__ w, const double* __restrict__ logdiag, int npad,
                           int N, double* __restrict__ out) {
    __shared__ double sh[2][32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) { a = fma(w[i], w[i], a); b += logdiag[i]; }
#pragma unroll
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][wp] = a; sh[1][wp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double sa = 0.0, sb = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { sa += sh[0][i]; sb += sh[1][i]; }
        out[0] = -0.5 * sa - sb - (double)N * 1.8378770664093454836;   // log(2 pi)
    }
}

cudaError_t solve_alpha_lml(const double* Z, long ldz, int npad, int N, const double* y_block,
                            double* y_int, double* w, double* alpha_int, double* partial,
                            const double* logdiag, double* lml_out, cudaStream_t st) {
    interleave_kernel<<<(npad + 255) / 256, 256, 0, st>>>(y_block, N, y_int, npad);
    trmv_lower_kernel<<<(npad + 7) / 8, 256, 0, st>>>(Z, ldz, y_int, w, npad);
    int nchunks = (npad + TRMVT_ROWS - 1) / TRMVT_ROWS;
    trmvT_partial_kernel<<<dim3(npad / 128, nchunks), 128, 0, st>>>(Z, ldz, w, partial, npad);
    colsum_partials_kernel<<<(npad + 255) / 256, 256, 0, st>>>(partial, nchunks, npad, alpha_int);
    lml_kernel<<<1, 1024, 0, st>>>(w, logdiag, npad, N, lml_out);
    return cudaGetLastError();
}

cudaError_t deinterleave(const double* xi, int N, double* x, cudaStream_t st) {
    deinterleave_kernel<<<(2 * N + 255) / 256, 256, 0, st>>>(xi, N, x);
    return cudaGetLastError();
}

}  // namespace gp2d
