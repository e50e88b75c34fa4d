// extern "C" surface of libgp2d (declared in include/gp2d.h).
#include "../../include/gp2d.h"

#include <math.h>
#include <stdio.h>

#include "common.cuh"
#include "dgemm.cuh"
#include "linalg.h"

using namespace gp2d;

namespace {

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : -(1000 + (int)e); }
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Private layout of the fit workspace: a pure function of N.
struct FitLayout {
    int npad;
    size_t off_A, off_Z, off_Zt, off_logdiag, off_yint, off_w, off_r, off_alpha, off_partial, off_X, off_scal, off_info, total;
    size_t off_perm;              // spatial order of the observations inside the fit (order.cu), order_n ints
    int order_n;                  // number of points when the fit sorts them, else 0
    size_t off_Zq, off_zunit;     // int8 digit slices of Z and their row units (Helmholtz layouts; 0 bytes otherwise)
    int i8;                       // the layout carries them
};
constexpr size_t GATE_OFF = 2 * sizeof(int);     // int after info in the off_info block: slice count of the int8 path, 0 = fp64 only

// n: scalar observations; x_doubles: size of the copy of the observation points;
// grad_doubles: partial-sum doubles the likelihood-gradient reduction needs
FitLayout fit_layout_general(size_t n_scalar, size_t x_doubles, size_t grad_doubles, bool i8 = false, int order_pts = 0) {
    FitLayout L;
    L.npad = round_up((int)n_scalar, TILE);
    const size_t n = (size_t)L.npad, d = sizeof(double);
    size_t o = 0;
    // scratch of the factorisation / likelihood
    L.off_A = o; o = align256(o + n * n * d);
    L.off_Z = o; o = align256(o + n * n * d);
    L.off_logdiag = o; o = align256(o + n * d);
    L.off_yint = o; o = align256(o + n * d);
    L.off_w = o; o = align256(o + n * d);
    L.off_r = o; o = align256(o + n * d);
    size_t nchunks = (n + 63) / 64;           // TRMVT_ROWS of linalg.cu
    size_t part = nchunks * n;
    if (grad_doubles > part) part = grad_doubles;
    L.off_partial = o; o = align256(o + part * d);
    L.order_n = 0;
    L.off_perm = 0;
    // what gp2d_predict reads, contiguous so that it can be shipped to another GPU in one piece
    L.off_Zt = o; o = align256(o + packed_tiles_doubles(L.npad) * d);
    L.i8 = i8 && L.npad <= i8_max_npad();
    // the sort serves the zero-slice skip of the int8 predictive kernel; below a few row blocks there is nothing to skip
    if (L.i8 && order_pts >= 256 && order_pts <= spatial_order_max_points()) L.order_n = order_pts;
    L.off_Zq = o; o = align256(o + (L.i8 ? i8_zq_bytes(L.npad) : 0));
    L.off_zunit = o; o = align256(o + (L.i8 ? n * d : 0));
    L.off_perm = o; o = align256(o + (size_t)L.order_n * sizeof(int));
    L.off_alpha = o; o = align256(o + n * d);
    L.off_X = o; o = align256(o + x_doubles * d);
    L.off_scal = o; o = align256(o + 32 * d);
    L.off_info = o; o = align256(o + 16);
    L.total = o;
    return L;
}

// ldx = 2: points (a, b); ldx = 3: space-time points (t, a, b)
FitLayout fit_layout(int N, int ldx = 2) {
    return fit_layout_general(2 * (size_t)N, (size_t)ldx * N,
                              (HELM_NP + 1) * (size_t)lml_grad_partials(round_up(2 * N, TILE)), /*i8=*/true, /*order_pts=*/N);
}

FitLayout rbf_layout(int N, int D) {
    return fit_layout_general((size_t)N, (size_t)N * D, (size_t)rbf_lml_grad_partial_doubles(round_up(N, TILE)));
}

template <class T>
inline T* at(void* base, size_t off) { return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + off); }
template <class T>
inline const T* at(const void* base, size_t off) { return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) + off); }

// the predictive kernel lists at most 1024 row blocks of 128 (and two n x n fp64 matrices of that
// size are already beyond one GPU's memory)
constexpr int MAX_NPAD = 131072;
bool size_ok(long n_scalar) { return n_scalar > 0 && n_scalar <= MAX_NPAD; }

bool theta_ok(double l_df, double l_cf, double ratio) {
    return l_df > 0.0 && l_cf > 0.0 && ratio >= 0.0 && ratio <= 1.0 && isfinite(l_df) && isfinite(l_cf);
}

// Robust mode of the factorisation.  The recursion forms every panel of L with the explicit inverse of
// the block above it, so its error grows with the condition number; cond(K + d I) <= n k** / d is
// known from the arguments alone.  Past 1e7 (parity at 1e-8 is no longer safe) the panels are refined
// against the factor (potri_lower, t_refine): +2 GEMMs per panel, no side-stream overlap.  The GP
// configurations of BASELINE.json sit at 1e4-1e5; the per-drifter track models at 1e12-1e13.
thread_local double g_robust_cond = 1e7;   // bring-up hook gp2d_dbg_set_robust_cond: per host thread
int refine_steps_for(double kss, long n_scalar, double diag_add) {
    if (!(diag_add > 0.0)) return 1;
    return (kss * (double)n_scalar / diag_add > g_robust_cond) ? 1 : 0;
}

// int8-sliced predictive kernel (predict_i8.cu).  Its error is that of dropping the digit products beyond
// the S-th.  Measured against an extended-precision product (tools/ozaki_emulate.py) the relative error of the
// variance follows the prior-to-noise ratio r = k** / (noise + jitter), not the matrix size: S = 6 gives
// 2e-10 at r = 4 (BASELINE configurations), 3e-9 at r = 20-40, 8e-8 at r = 400; S = 7 gives 1e-12, 1.5e-11 and
// 3e-10; from n = 2000 to n = 8000 both grow by 1.6.  With m = r sqrt(n / 4000): 6 slices up to m = 20, 7 up to
// m = 2000 (a factor 3 inside the 1e-8 parity bar), the fp64 kernel beyond, and always in the robust mode.
// Per host thread: 0 = automatic, 1 = off, 6 / 7 = that slice count whenever the layout carries the slices.
thread_local int g_i8_mode = 0;
constexpr double I8_M_S6 = 20.0, I8_M_S7 = 2000.0;
int i8_slices_for(double kss, long n_scalar, double diag_add, const FitLayout& L) {
    if (!L.i8 || g_i8_mode == 1) return 0;
    if (g_i8_mode == 6 || g_i8_mode == 7) return g_i8_mode;
    if (!(diag_add > 0.0) || refine_steps_for(kss, n_scalar, diag_add)) return 0;
    const double m = kss / diag_add * sqrt((double)n_scalar / 4000.0);
    return m <= I8_M_S6 ? 6 : m <= I8_M_S7 ? 7 : 0;
}

// factor + inverse of the padded covariance in ws (A destroyed or replaced by L), robust when asked
cudaError_t factor_ws(void* ws, const FitLayout& L, int refine, cudaStream_t st) {
    double* A = at<double>(ws, L.off_A);
    double* Z = at<double>(ws, L.off_Z);
    // robust mode keeps L in A and borrows the packed-tile region (written after the factorisation,
    // npad (npad + 128) / 2 doubles >= (npad / 2)^2) as panel scratch
    return potri_lower(A, L.npad, Z, L.npad, L.npad, at<double>(ws, L.off_logdiag), at<int>(ws, L.off_info),
                       /*need_inv=*/true, /*keep_L=*/refine > 0, refine > 0 ? at<double>(ws, L.off_Zt) : nullptr, st, refine);
}

// robust mode, after the covariance has been rebuilt into A: two refinement steps of alpha (linalg.h)
cudaError_t refine_alpha_ws(void* ws, const FitLayout& L, cudaStream_t st) {
    return refine_alpha(at<double>(ws, L.off_A), L.npad, at<double>(ws, L.off_Z), L.npad, L.npad, at<double>(ws, L.off_yint),
                        at<double>(ws, L.off_w), at<double>(ws, L.off_r), at<double>(ws, L.off_alpha),
                        at<double>(ws, L.off_partial), 2, st);
}

__global__ void fill_kernel(double* out, int n, double v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

// user matrix (lower) -> padded square with identity in the padding
__global__ void pad_lower_kernel(const double* __restrict__ A, long lda, int n, double* __restrict__ P, int npad) {
    long total = (long)npad * npad;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int r = (int)(idx / npad), c = (int)(idx % npad);
        double v = 0.0;
        if (r < n && c < n) v = (c <= r) ? A[(long)r * lda + c] : 0.0;
        else if (r == c) v = 1.0;
        P[idx] = v;
    }
}
__global__ void unpad_lower_kernel(const double* __restrict__ P, int npad, double* __restrict__ A, long lda, int n) {
    long total = (long)n * n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int r = (int)(idx / n), c = (int)(idx % n);
        if (c <= r) A[(long)r * lda + c] = P[(long)r * npad + c];
    }
}

// mode 0: DMMA only; 1: DFMA only; 2: 16 DMMA + 16 DFMA per iteration (shared FP64 pipe?)
template <int MODE>
__global__ void __launch_bounds__(NTHREADS) fp64_peak_kernel(double* out, int iters) {
    double c[16][2], d[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { c[j][0] = c[j][1] = 0.0; d[j] = 1e-3 * j; }
    const double a = 1e-3 * (threadIdx.x & 31), b = 1e-3;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dmma884(c[j][0], c[j][1], a, b);
        }
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { c[j][0] = fma(c[j][0], a, b); c[j][1] = fma(c[j][1], a, b); }
        }
        if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) d[j] = fma(d[j], a, b);
        }
        if (MODE == 3) {      // one dependent DFMA chain: latency = time / (16 * iters)
#pragma unroll
            for (int j = 0; j < 16; ++j) d[0] = fma(d[0], a, b);
        }
        if (MODE == 4) {      // one dependent DMMA chain
#pragma unroll
            for (int j = 0; j < 16; ++j) dmma884(c[0][0], c[0][1], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += c[j][0] + c[j][1] + d[j];
    if (s == 123.456) out[0] = s;      // keep the loops alive
}

// padded lower-tile K^-1 -> full symmetric user matrix
__global__ void unpad_symmetric_kernel(const double* __restrict__ P, int npad, double* __restrict__ A, long lda, int n) {
    long total = (long)n * n;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        int r = (int)(idx / n), c = (int)(idx % n);
        A[(long)r * lda + c] = (c <= r) ? P[(long)r * npad + c] : P[(long)c * npad + r];
    }
}

// the fit's permutation of the observations (null: caller's order)
inline const int* perm_of(const void* ws, const FitLayout& L) { return L.order_n ? at<int>(ws, L.off_perm) : nullptr; }

// observations into the fit state: in Z-order when the layout says so (order.cu), as given otherwise
cudaError_t stage_points(const double* X, int ldx, int xo, int N, void* ws, const FitLayout& L, cudaStream_t st) {
    if (!L.order_n) return cudaMemcpyAsync(at<double>(ws, L.off_X), X, (size_t)ldx * N * sizeof(double), cudaMemcpyDeviceToDevice, st);
    cudaError_t e = spatial_order(X, ldx, xo, N, at<int>(ws, L.off_perm), st);
    if (e != cudaSuccess) return e;
    return gather_points(X, ldx, N, at<int>(ws, L.off_perm), at<double>(ws, L.off_X), st);
}

// fit core shared by gp2d_fit and gp2d_lml_grad
cudaError_t fit_core(const double* X, int N, const double* y, const HelmParams& hp, double diag_add,
                     void* ws, const FitLayout& L, cudaStream_t st) {
    double* A = at<double>(ws, L.off_A);
    double* Z = at<double>(ws, L.off_Z);
    cudaError_t e;
    e = stage_points(X, hp.ldx, hp.xo, N, ws, L, st);
    if (e != cudaSuccess) return e;
    X = at<double>(ws, L.off_X);                  // from here on the observations in the fit's own order
    e = build_interleaved_lower(X, N, hp, diag_add, A, L.npad, L.npad, st);
    if (e != cudaSuccess) return e;
    const int refine = refine_steps_for(hp.tvar * (hp.w_df + hp.w_cf), 2L * N, diag_add);
    e = factor_ws(ws, L, refine, st);
    if (e != cudaSuccess) return e;
    e = pack_lower_tiles(Z, L.npad, L.npad, at<double>(ws, L.off_Zt), st);
    if (e != cudaSuccess) return e;
    {   // digit slices for the int8 predictive kernel (off_r is free until refine_alpha)
        const int s8 = i8_slices_for(hp.tvar * (hp.w_df + hp.w_cf), 2L * N, diag_add, L);
        if (s8) e = i8_quantize_lower(Z, L.npad, L.npad, at<double>(ws, L.off_zunit), at<double>(ws, L.off_r), at<int8_t>(ws, L.off_Zq), st);
        if (e != cudaSuccess) return e;
        e = i8_set_gate(at<int>(ws, L.off_info + GATE_OFF), s8, st);
        if (e != cudaSuccess) return e;
    }
    e = solve_alpha_lml(Z, L.npad, L.npad, N, 2, y, at<double>(ws, L.off_yint), at<double>(ws, L.off_w),
                        at<double>(ws, L.off_alpha), at<double>(ws, L.off_partial),
                        at<double>(ws, L.off_logdiag), at<double>(ws, L.off_scal), st, 1, 0, 0, perm_of(ws, L), 0);
    if (e != cudaSuccess || !refine) return e;
    e = build_interleaved_lower(X, N, hp, diag_add, A, L.npad, L.npad, st);      // A held the factor: the matrix again
    if (e != cudaSuccess) return e;
    return refine_alpha_ws(ws, L, st);
}

// Helmholtz-family prediction from a fit workspace: the int8 kernels run when the fit chose a slice count for
// them (a flag in the fit state, read on the device), the fp64 kernel when it did not.
cudaError_t predict_helm(const void* fit_ws, const FitLayout& L, int N, const HelmParams& hp, const double* Xs, int M,
                         long out_stride, double var_add, double* mean, double* var, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int* gate = nullptr;
    if (L.i8 && g_i8_mode != 1 && ws_bytes >= predict_i8_min_scratch_bytes(L.npad)) {
        gate = at<int>(fit_ws, L.off_info + GATE_OFF);
        cudaError_t e = predict_fused_i8(at<int8_t>(fit_ws, L.off_Zq), at<double>(fit_ws, L.off_zunit), gate, L.npad,
                                         at<double>(fit_ws, L.off_alpha), at<double>(fit_ws, L.off_X), N, hp, Xs, M, out_stride,
                                         var_add, mean, var, ws, ws_bytes, st);
        if (e != cudaSuccess) return e;
    }
    return predict_fused(at<double>(fit_ws, L.off_Zt), L.npad, at<double>(fit_ws, L.off_alpha), at<double>(fit_ws, L.off_X), N,
                         hp, Xs, M, out_stride, var_add, mean, var, (double*)ws, ws_bytes, st, gate);
}

}  // namespace

extern "C" {

int gp2d_version(void) { return GP2D_VERSION; }

const char* gp2d_error_string(int code) {
    static thread_local char buf[160];
    if (code == 0) return "ok";
    if (code > 0) { snprintf(buf, sizeof buf, "matrix not positive definite: pivot %d", code); return buf; }
    if (code > -1000) { snprintf(buf, sizeof buf, "invalid argument %d", -code); return buf; }
    snprintf(buf, sizeof buf, "CUDA error: %s", cudaGetErrorString((cudaError_t)(-code - 1000)));
    return buf;
}

int gp2d_kernel_build(const double* X, int N, const double* X2, int M, double l_df, double l_cf,
                      double ratio, double diag_add, double* K, int64_t ldk, void* stream) {
    if (!X) return -1;
    if (N < 0) return -2;
    if (M < 0) return -4;
    if (!theta_ok(l_df, l_cf, ratio)) return -5;
    if (!K && N > 0 && M > 0) return -9;
    if (ldk < 2 * (int64_t)M) return -10;
    if (X2 == nullptr && M != N) return -4;
    return cuda_rc(build_block_layout(X, N, X2, M, make_helm(l_df, l_cf, ratio), diag_add, K, (long)ldk,
                                      (cudaStream_t)stream));
}

int gp2d_kdiag(int M, double l_df, double l_cf, double ratio, double* out, void* stream) {
    if (M < 0) return -1;
    if (!theta_ok(l_df, l_cf, ratio)) return -2;
    if (M == 0) return 0;
    if (!out) return -5;
    HelmParams hp = make_helm(l_df, l_cf, ratio);
    fill_kernel<<<(2 * M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(out, 2 * M, hp.w_df + hp.w_cf);
    return cuda_rc(cudaGetLastError());
}

size_t gp2d_kernel_grad_workspace_bytes(int N, int M) {
    if (N <= 0 || M <= 0) return 256;
    return align256(HELM_NP * sizeof(double) * ((size_t)grad_sums_block_partials(N, M) + 1));
}

int gp2d_kernel_grad(const double* X, int N, const double* X2, int M, double l_df, double l_cf,
                     double ratio, int reference_compat, const double* dL_dK, int64_t ld, void* ws,
                     size_t ws_bytes, double* out3, void* stream) {
    if (!X) return -1;
    if (N <= 0) return -2;
    if (M <= 0 || (X2 == nullptr && M != N)) return -4;
    if (!theta_ok(l_df, l_cf, ratio)) return -5;
    if (!dL_dK) return -9;
    if (ld < 2 * (int64_t)M) return -10;
    if (!ws || ws_bytes < gp2d_kernel_grad_workspace_bytes(N, M)) return -12;
    if (!out3) return -13;
    return cuda_rc(kernel_grad_sums_block(X, N, X2, M, make_helm(l_df, l_cf, ratio), reference_compat != 0,
                                          dL_dK, (long)ld, (double*)ws, (int)(ws_bytes / (HELM_NP * sizeof(double))),
                                          out3, (cudaStream_t)stream));
}

size_t gp2d_potrf_workspace_bytes(int n) {
    if (n <= 0) return 256;
    size_t np = (size_t)round_up(n, TILE);
    return align256(np * np * 8) * 2 + align256((np / 2) * (np / 2) * 8 + 256) + align256(np * 8);
}

int gp2d_potrf(double* A, int n, int64_t lda, void* ws, size_t ws_bytes, int* info, void* stream) {
    if (!A) return -1;
    if (n <= 0) return -2;
    if (lda < n) return -3;
    if (!ws || ws_bytes < gp2d_potrf_workspace_bytes(n)) return -5;
    if (!info) return -6;
    cudaStream_t st = (cudaStream_t)stream;
    const int np = round_up(n, TILE);
    size_t o = 0;
    double* P = at<double>(ws, o); o += align256((size_t)np * np * 8);
    double* Z = at<double>(ws, o); o += align256((size_t)np * np * 8);
    double* W = at<double>(ws, o); o += align256((size_t)(np / 2) * (np / 2) * 8 + 256);
    double* logdiag = at<double>(ws, o);
    pad_lower_kernel<<<592, 256, 0, st>>>(A, (long)lda, n, P, np);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_rc(e);
    e = potri_lower(P, np, Z, np, np, logdiag, info, /*need_inv=*/false, /*keep_L=*/true, W, st);
    if (e != cudaSuccess) return cuda_rc(e);
    unpad_lower_kernel<<<592, 256, 0, st>>>(P, np, A, (long)lda, n);
    return cuda_rc(cudaGetLastError());
}

size_t gp2d_spd_inverse_workspace_bytes(int n) {
    if (n <= 0) return 256;
    size_t np = (size_t)round_up(n, TILE);
    return align256(np * np * 8) * 2 + align256((np / 2) * (np / 2) * 8 + 256) + align256(np * 8);
}

int gp2d_spd_inverse(double* A, int n, int64_t lda, void* ws, size_t ws_bytes, int* info, void* stream) {
    if (!A) return -1;
    if (n <= 0) return -2;
    if (lda < n) return -3;
    if (!ws || ws_bytes < gp2d_spd_inverse_workspace_bytes(n)) return -5;
    if (!info) return -6;
    cudaStream_t st = (cudaStream_t)stream;
    const int np = round_up(n, TILE);
    size_t o = 0;
    double* P = at<double>(ws, o); o += align256((size_t)np * np * 8);
    double* Z = at<double>(ws, o); o += align256((size_t)np * np * 8);
    double* W = at<double>(ws, o); o += align256((size_t)(np / 2) * (np / 2) * 8 + 256);
    double* logdiag = at<double>(ws, o);
    pad_lower_kernel<<<592, 256, 0, st>>>(A, (long)lda, n, P, np);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_rc(e);
    // a general-purpose inverse has to cope with ill-conditioned input: factor panels refined (linalg.h)
    e = potri_lower(P, np, Z, np, np, logdiag, info, /*need_inv=*/true, /*keep_L=*/true, W, st, /*t_refine=*/1);
    if (e != cudaSuccess) return cuda_rc(e);
    e = launch_dgemm(true, true, GemmArgs{Z, np, Z, np, P, np, np, np, np, 1.0, 0.0, 1, KR_GE_M}, st);
    if (e != cudaSuccess) return cuda_rc(e);
    unpad_symmetric_kernel<<<592, 256, 0, st>>>(P, np, A, (long)lda, n);
    return cuda_rc(cudaGetLastError());
}

int gp2d_dgemm(int transa, int transb, int M, int N, int K, double alpha, const double* A, int64_t lda,
               const double* B, int64_t ldb, double beta, double* C, int64_t ldc, void* stream) {
    if (M <= 0 || M % TILE) return -3;
    if (N <= 0 || N % TILE) return -4;
    if (K <= 0 || K % BK) return -5;
    if (!A || (reinterpret_cast<uintptr_t>(A) & 15)) return -7;
    if ((lda & 1) || lda < (transa ? M : K)) return -8;
    if (!B || (reinterpret_cast<uintptr_t>(B) & 15)) return -9;
    if ((ldb & 1) || ldb < (transb ? K : N)) return -10;
    if (!C || (reinterpret_cast<uintptr_t>(C) & 15)) return -12;
    if ((ldc & 1) || ldc < N) return -13;
    // kernel operand conventions: A "MN-major" == transposed storage; B "MN-major" == [k][n] storage
    return cuda_rc(launch_dgemm(transa != 0, transb == 0,
                                GemmArgs{A, (long)lda, B, (long)ldb, C, (long)ldc, M, N, K, alpha, beta, 0, KR_FULL},
                                (cudaStream_t)stream));
}

size_t gp2d_fit_workspace_bytes(int N) {
    if (!size_ok(2L * N)) return 0;
    return fit_layout(N).total;
}

int gp2d_fit_predict_state(int N, size_t* offset, size_t* bytes) {
    if (N <= 0) return -1;
    if (!offset) return -2;
    if (!bytes) return -3;
    FitLayout L = fit_layout(N);
    *offset = L.off_Zt;
    *bytes = L.total - L.off_Zt;
    return 0;
}

int gp2d_fit(const double* X, int N, const double* y, double l_df, double l_cf, double ratio,
             double noise, double jitter, void* ws, size_t ws_bytes, double* alpha_out,
             double* lml_out, int* info, void* stream) {
    if (!X) return -1;
    if (!size_ok(2L * N)) return -2;
    if (!y) return -3;
    if (!theta_ok(l_df, l_cf, ratio)) return -4;
    if (!(noise >= 0.0)) return -7;
    if (!(jitter >= 0.0)) return -8;
    FitLayout L = fit_layout(N);
    if (!ws) return -9;
    if (ws_bytes < L.total) return -10;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = fit_core(X, N, y, make_helm(l_df, l_cf, ratio), noise + jitter, ws, L, st);
    if (e != cudaSuccess) return cuda_rc(e);
    if (alpha_out) {
        e = deinterleave(at<double>(ws, L.off_alpha), N, alpha_out, st, 1, 0, perm_of(ws, L), 0);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    if (lml_out) {
        e = cudaMemcpyAsync(lml_out, at<double>(ws, L.off_scal), sizeof(double), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    if (info) {
        e = cudaMemcpyAsync(info, at<int>(ws, L.off_info), sizeof(int), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    return 0;
}

size_t gp2d_predict_workspace_bytes(int N, int M) {
    if (N <= 0 || M <= 0) return 256;
    const int npad = round_up(2 * N, TILE);
    size_t b = predict_scratch_bytes(npad, M, 64);
    if (npad <= i8_max_npad()) {
        const size_t b8 = predict_i8_scratch_bytes(npad);
        if (b8 > b) b = b8;
    }
    return align256(b);
}

int gp2d_set_option(int key, double value) {
    if (key == GP2D_OPT_PREDICT_I8) {
        const int v = (int)value;
        if (v != 0 && v != 1 && v != 6 && v != 7) return -2;
        g_i8_mode = v;
        return 0;
    }
    return -1;
}

int gp2d_predict(const void* fit_ws, int N, double l_df, double l_cf, double ratio, const double* Xs,
                 int M, int64_t out_stride, double var_add, double* mean, double* var, void* ws,
                 size_t ws_bytes, void* stream) {
    if (!fit_ws) return -1;
    if (!size_ok(2L * N)) return -2;
    if (!theta_ok(l_df, l_cf, ratio)) return -3;
    if (M < 0) return -7;
    if (M == 0) return 0;
    if (!Xs) return -6;
    if (out_stride < M) return -8;
    if (!mean) return -10;
    if (!var) return -11;
    FitLayout L = fit_layout(N);
    if (!ws) return -12;
    if (ws_bytes < predict_panel_bytes(L.npad)) return -13;
    return cuda_rc(predict_helm(fit_ws, L, N, make_helm(l_df, l_cf, ratio), Xs, M, (long)out_stride, var_add, mean, var, ws,
                                ws_bytes, (cudaStream_t)stream));
}

int gp2d_lml_grad(const double* X, int N, const double* y, double l_df, double l_cf, double ratio,
                  double noise, double jitter, int reference_compat, void* ws, size_t ws_bytes,
                  double* out5, int* info, void* stream) {
    if (!X) return -1;
    if (!size_ok(2L * N)) return -2;
    if (!y) return -3;
    if (!theta_ok(l_df, l_cf, ratio)) return -4;
    if (!(noise >= 0.0)) return -7;
    if (!(jitter >= 0.0)) return -8;
    FitLayout L = fit_layout(N);
    if (!ws) return -10;
    if (ws_bytes < L.total) return -11;
    if (!out5) return -12;
    cudaStream_t st = (cudaStream_t)stream;
    HelmParams hp = make_helm(l_df, l_cf, ratio);
    cudaError_t e = fit_core(X, N, y, hp, noise + jitter, ws, L, st);
    if (e != cudaSuccess) return cuda_rc(e);
    // K^-1 = Z^T Z (lower tiles) into the now-free A buffer
    double* Z = at<double>(ws, L.off_Z);
    double* Kinv = at<double>(ws, L.off_A);
    e = launch_dgemm(true, true, GemmArgs{Z, L.npad, Z, L.npad, Kinv, L.npad, L.npad, L.npad, L.npad, 1.0, 0.0, 1, KR_GE_M}, st);
    if (e != cudaSuccess) return cuda_rc(e);
    double* scal = at<double>(ws, L.off_scal);
    // scal = (LML, d/dl_df, d/dl_cf, d/dratio, d/dtvar, d/dlt, d/dnoise)
    e = lml_grad_reduce(Kinv, L.npad, L.npad, at<double>(ws, L.off_alpha), at<double>(ws, L.off_X), N, hp,
                        reference_compat != 0, at<double>(ws, L.off_partial), scal + 1, st);
    if (e != cudaSuccess) return cuda_rc(e);
    e = cudaMemcpyAsync(out5, scal, 4 * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_rc(e);
    e = cudaMemcpyAsync(out5 + 4, scal + 6, sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_rc(e);
    if (info) {
        e = cudaMemcpyAsync(info, at<int>(ws, L.off_info), sizeof(int), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    return 0;
}

/* ---- batches of independent problems of one size (survey kernel k6) ------------------------------- */

namespace {

// scal[16..29) of every problem's workspace holds its BatchPar (kernel parameters + diagonal term)
constexpr size_t BATCH_PAR_OFF = 16 * sizeof(double);
static_assert(BATCH_PAR_OFF + sizeof(BatchPar) <= 32 * sizeof(double), "BatchPar must fit the scalar block");

// out[b][0..4] = (LML, d/dl_df, d/dl_cf, d/dratio, d/dnoise) (grad) or out[b] = LML; info[b]
__global__ void gather_batch_out_kernel(const double* __restrict__ scal, const int* __restrict__ info_ws, long bstride,
                                        int with_grad, double* __restrict__ out, int* __restrict__ info, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const double* s = scal + (long)b * bstride;
    if (out) {
        if (with_grad) {
            out[5 * b + 0] = s[0]; out[5 * b + 1] = s[1]; out[5 * b + 2] = s[2]; out[5 * b + 3] = s[3]; out[5 * b + 4] = s[6];
        } else {
            out[b] = s[0];
        }
    }
    if (info) info[b] = info_ws[(long)b * 2 * bstride];
}

struct BatchProblem { HelmParams hp; double noise, diag_add; int refine; };

// validates theta4 = [B][4] rows (l_df, l_cf, ratio, noise) and fills the per-problem records
int batch_problems(const double* theta4, int B, int N, double jitter, BatchProblem* out) {
    for (int b = 0; b < B; ++b) {
        const double* t = theta4 + 4 * b;
        if (!theta_ok(t[0], t[1], t[2]) || !(t[3] >= 0.0)) return -1;
        out[b].hp = make_helm(t[0], t[1], t[2]);
        out[b].noise = t[3];
        out[b].diag_add = t[3] + jitter;
        out[b].refine = refine_steps_for(out[b].hp.w_df + out[b].hp.w_cf, 2L * N, out[b].diag_add);
    }
    return 0;
}

// plain-mode problems [b0, b0 + nb) in one chain of launches (every launch covers the whole run)
cudaError_t fit_core_batched(const double* X, long x_stride, int N, const double* y, long y_stride, const BatchProblem* pr,
                             int nb, void* ws0, const FitLayout& L, cudaStream_t st) {
    const long bs = (long)(L.total / sizeof(double));
    BatchPar host[64];
    BatchPar* par = at<BatchPar>(ws0, L.off_scal + BATCH_PAR_OFF);
    cudaError_t e = cudaSuccess;
    for (int c0 = 0; c0 < nb && e == cudaSuccess; c0 += 64) {
        const int nc = nb - c0 < 64 ? nb - c0 : 64;
        for (int i = 0; i < nc; ++i) { host[i].hp = pr[c0 + i].hp; host[i].diag_add = pr[c0 + i].diag_add; }
        e = scatter_batch_params(host, nc, reinterpret_cast<BatchPar*>(reinterpret_cast<double*>(par) + (long)c0 * bs), bs,
                                 X + (long)c0 * x_stride, x_stride, 2 * N, at<double>(ws0, L.off_X) + (long)c0 * bs, st);
    }
    if (e != cudaSuccess) return e;
    if (L.order_n) {      // every problem sorts its own observations (the same order for shared data), as fit_core does
        e = spatial_order(X, 2, 0, N, at<int>(ws0, L.off_perm), st, nb, x_stride, 2 * bs);
        if (e != cudaSuccess) return e;
        e = gather_points(X, 2, N, at<int>(ws0, L.off_perm), at<double>(ws0, L.off_X), st, nb, x_stride, 2 * bs, bs);
        if (e != cudaSuccess) return e;
    }
    double* A = at<double>(ws0, L.off_A);
    double* Z = at<double>(ws0, L.off_Z);
    e = build_interleaved_lower_batched(at<double>(ws0, L.off_X), bs, N, par, A, L.npad, L.npad, nb, bs, st);
    if (e != cudaSuccess) return e;
    e = potri_lower(A, L.npad, Z, L.npad, L.npad, at<double>(ws0, L.off_logdiag), at<int>(ws0, L.off_info),
                    /*need_inv=*/true, /*keep_L=*/false, nullptr, st, 0, nb, bs);
    if (e != cudaSuccess) return e;
    e = pack_lower_tiles(Z, L.npad, L.npad, at<double>(ws0, L.off_Zt), st, nb, bs);
    if (e != cudaSuccess) return e;
    {   // digit slices for the int8 predictive kernel, exactly as fit_core prepares them
        bool any = false;
        for (int b = 0; b < nb; ++b) any = any || i8_slices_for(pr[b].hp.w_df + pr[b].hp.w_cf, 2L * N, pr[b].diag_add, L) != 0;
        if (any) e = i8_quantize_lower(Z, L.npad, L.npad, at<double>(ws0, L.off_zunit), at<double>(ws0, L.off_r),
                                       at<int8_t>(ws0, L.off_Zq), st, nb, bs);
        for (int b = 0; b < nb && e == cudaSuccess; ++b)
            e = i8_set_gate(at<int>(ws0, L.off_info + GATE_OFF) + 2 * (long)b * bs,
                            i8_slices_for(pr[b].hp.w_df + pr[b].hp.w_cf, 2L * N, pr[b].diag_add, L), st);
        if (e != cudaSuccess) return e;
    }
    return solve_alpha_lml(Z, L.npad, L.npad, N, 2, y, at<double>(ws0, L.off_yint), at<double>(ws0, L.off_w),
                           at<double>(ws0, L.off_alpha), at<double>(ws0, L.off_partial), at<double>(ws0, L.off_logdiag),
                           at<double>(ws0, L.off_scal), st, nb, y_stride, bs, perm_of(ws0, L), 2 * bs);
}

// shared driver of gp2d_fit_batched / gp2d_lml_grad_batched: runs of consecutive plain-mode problems go
// through the batched chain, robust-mode problems (ill-conditioned, rare) one by one through the
// single-problem entry points, so every problem gets exactly the arithmetic of its unbatched call
int run_batch(bool grad, const double* X, int64_t x_stride, int N, const double* y, int64_t y_stride, int B,
              const double* theta4, double jitter, int compat, void* ws, double* alpha_out, double* out, int* info,
              cudaStream_t st) {
    const FitLayout L = fit_layout(N);
    const long bs = (long)(L.total / sizeof(double));
    BatchProblem* pr = new BatchProblem[B];
    if (batch_problems(theta4, B, N, jitter, pr)) { delete[] pr; return -7; }
    int rc = 0;
    cudaError_t e = cudaSuccess;
    for (int b0 = 0; b0 < B && rc == 0 && e == cudaSuccess;) {
        char* ws0 = reinterpret_cast<char*>(ws) + (size_t)b0 * L.total;
        if (pr[b0].refine) {
            const double* t = theta4 + 4 * b0;
            if (grad) rc = gp2d_lml_grad(X + b0 * x_stride, N, y + b0 * y_stride, t[0], t[1], t[2], t[3], jitter, compat, ws0,
                                         L.total, out + 5 * b0, info ? info + b0 : nullptr, st);
            else rc = gp2d_fit(X + b0 * x_stride, N, y + b0 * y_stride, t[0], t[1], t[2], t[3], jitter, ws0, L.total,
                               alpha_out ? alpha_out + (size_t)2 * N * b0 : nullptr, out ? out + b0 : nullptr,
                               info ? info + b0 : nullptr, st);
            ++b0;
            continue;
        }
        int nb = 1;
        while (b0 + nb < B && !pr[b0 + nb].refine) ++nb;
        e = fit_core_batched(X + b0 * x_stride, (long)x_stride, N, y + b0 * y_stride, (long)y_stride, pr + b0, nb, ws0, L, st);
        if (e != cudaSuccess) break;
        if (grad) {
            double* Z = at<double>(ws0, L.off_Z);
            double* Kinv = at<double>(ws0, L.off_A);
            GemmArgs g{Z, L.npad, Z, L.npad, Kinv, L.npad, L.npad, L.npad, L.npad, 1.0, 0.0, 1, KR_GE_M};
            g.batch = nb; g.bsA = g.bsB = g.bsC = nb > 1 ? bs : 0;
            e = launch_dgemm(true, true, g, st);
            if (e != cudaSuccess) break;
            e = lml_grad_reduce_batched(Kinv, L.npad, L.npad, at<double>(ws0, L.off_alpha), at<double>(ws0, L.off_X), N,
                                        at<BatchPar>(ws0, L.off_scal + BATCH_PAR_OFF), compat, at<double>(ws0, L.off_partial),
                                        at<double>(ws0, L.off_scal) + 1, nb, bs, st);
            if (e != cudaSuccess) break;
        } else if (alpha_out) {
            e = deinterleave(at<double>(ws0, L.off_alpha), N, alpha_out + (size_t)2 * N * b0, st, nb, bs, perm_of(ws0, L), 2 * bs);
            if (e != cudaSuccess) break;
        }
        gather_batch_out_kernel<<<(nb + 127) / 128, 128, 0, st>>>(at<double>(ws0, L.off_scal), at<int>(ws0, L.off_info), bs,
                                                                  grad ? 1 : 0, out ? out + (grad ? 5 : 1) * b0 : nullptr,
                                                                  info ? info + b0 : nullptr, nb);
        e = cudaGetLastError();
        b0 += nb;
    }
    delete[] pr;
    if (rc) return rc;
    return cuda_rc(e);
}

}  // namespace

size_t gp2d_fit_batched_workspace_bytes(int N, int B) {
    if (!size_ok(2L * N) || B < 1) return 0;
    return fit_layout(N).total * (size_t)B;
}

int gp2d_fit_batched(const double* X, int64_t x_stride, int N, const double* y, int64_t y_stride, int B,
                     const double* theta4, double jitter, void* ws, size_t ws_bytes, double* alpha_out,
                     double* lml_out, int* info, void* stream) {
    if (!X) return -1;
    if (x_stride != 0 && x_stride < 2 * (int64_t)N) return -2;
    if (!size_ok(2L * N)) return -3;
    if (!y) return -4;
    if (y_stride != 0 && y_stride < 2 * (int64_t)N) return -5;
    if (B < 1) return -6;
    if (!theta4) return -7;
    if (!(jitter >= 0.0)) return -8;
    if (!ws) return -9;
    if (ws_bytes < gp2d_fit_batched_workspace_bytes(N, B)) return -10;
    return run_batch(false, X, x_stride, N, y, y_stride, B, theta4, jitter, 0, ws, alpha_out, lml_out, info, (cudaStream_t)stream);
}

int gp2d_lml_grad_batched(const double* X, int64_t x_stride, int N, const double* y, int64_t y_stride, int B,
                          const double* theta4, double jitter, int reference_compat, void* ws, size_t ws_bytes,
                          double* out5, int* info, void* stream) {
    if (!X) return -1;
    if (x_stride != 0 && x_stride < 2 * (int64_t)N) return -2;
    if (!size_ok(2L * N)) return -3;
    if (!y) return -4;
    if (y_stride != 0 && y_stride < 2 * (int64_t)N) return -5;
    if (B < 1) return -6;
    if (!theta4) return -7;
    if (!(jitter >= 0.0)) return -8;
    if (!ws) return -10;
    if (ws_bytes < gp2d_fit_batched_workspace_bytes(N, B)) return -11;
    if (!out5) return -12;
    return run_batch(true, X, x_stride, N, y, y_stride, B, theta4, jitter, reference_compat != 0, ws, nullptr, out5, info,
                     (cudaStream_t)stream);
}

/* ---- space-time product kernel: tvar exp(-dt^2 / 2 lt^2) * Helmholtz(a, b) ------------------------ */

namespace {
bool st_ok(double tvar, double lt) { return tvar > 0.0 && lt > 0.0 && isfinite(tvar) && isfinite(lt); }
}

int gp2d_st_kernel_build(const double* X3, int N, const double* X3b, int M, double l_df, double l_cf, double ratio,
                         double tvar, double lt, double diag_add, double* K, int64_t ldk, void* stream) {
    if (!X3) return -1;
    if (N < 0) return -2;
    if (M < 0 || (X3b == nullptr && M != N)) return -4;
    if (!theta_ok(l_df, l_cf, ratio)) return -5;
    if (!st_ok(tvar, lt)) return -8;
    if (!K && N > 0 && M > 0) return -11;
    if (ldk < 2 * (int64_t)M) return -12;
    return cuda_rc(build_block_layout(X3, N, X3b, M, make_helm_st(l_df, l_cf, ratio, tvar, lt), diag_add, K, (long)ldk,
                                      (cudaStream_t)stream));
}

int gp2d_st_kernel_grad(const double* X3, int N, const double* X3b, int M, double l_df, double l_cf, double ratio,
                        double tvar, double lt, const double* dL_dK, int64_t ld, void* ws, size_t ws_bytes,
                        double* out5, void* stream) {
    if (!X3) return -1;
    if (N <= 0) return -2;
    if (M <= 0 || (X3b == nullptr && M != N)) return -4;
    if (!theta_ok(l_df, l_cf, ratio)) return -5;
    if (!st_ok(tvar, lt)) return -8;
    if (!dL_dK) return -10;
    if (ld < 2 * (int64_t)M) return -11;
    if (!ws || ws_bytes < gp2d_kernel_grad_workspace_bytes(N, M)) return -13;
    if (!out5) return -14;
    return cuda_rc(kernel_grad_sums_block(X3, N, X3b, M, make_helm_st(l_df, l_cf, ratio, tvar, lt), false, dL_dK,
                                          (long)ld, (double*)ws, (int)(ws_bytes / (HELM_NP * sizeof(double))), out5,
                                          (cudaStream_t)stream));
}

size_t gp2d_st_fit_workspace_bytes(int N) {
    if (!size_ok(2L * N)) return 0;
    return fit_layout(N, 3).total;
}

int gp2d_st_fit_predict_state(int N, size_t* offset, size_t* bytes) {
    if (N <= 0) return -1;
    if (!offset) return -2;
    if (!bytes) return -3;
    FitLayout L = fit_layout(N, 3);
    *offset = L.off_Zt;
    *bytes = L.total - L.off_Zt;
    return 0;
}

int gp2d_st_fit(const double* X3, int N, const double* y, double l_df, double l_cf, double ratio, double tvar,
                double lt, double noise, double jitter, void* ws, size_t ws_bytes, double* alpha_out,
                double* lml_out, int* info, void* stream) {
    if (!X3) return -1;
    if (!size_ok(2L * N)) return -2;
    if (!y) return -3;
    if (!theta_ok(l_df, l_cf, ratio)) return -4;
    if (!st_ok(tvar, lt)) return -7;
    if (!(noise >= 0.0)) return -9;
    if (!(jitter >= 0.0)) return -10;
    FitLayout L = fit_layout(N, 3);
    if (!ws) return -11;
    if (ws_bytes < L.total) return -12;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = fit_core(X3, N, y, make_helm_st(l_df, l_cf, ratio, tvar, lt), noise + jitter, ws, L, st);
    if (e != cudaSuccess) return cuda_rc(e);
    if (alpha_out) {
        e = deinterleave(at<double>(ws, L.off_alpha), N, alpha_out, st, 1, 0, perm_of(ws, L), 0);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    if (lml_out) {
        e = cudaMemcpyAsync(lml_out, at<double>(ws, L.off_scal), sizeof(double), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    if (info) {
        e = cudaMemcpyAsync(info, at<int>(ws, L.off_info), sizeof(int), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    return 0;
}

int gp2d_st_predict(const void* fit_ws, int N, double l_df, double l_cf, double ratio, double tvar, double lt,
                    const double* Xs3, int M, int64_t out_stride, double var_add, double* mean, double* var,
                    void* ws, size_t ws_bytes, void* stream) {
    if (!fit_ws) return -1;
    if (!size_ok(2L * N)) return -2;
    if (!theta_ok(l_df, l_cf, ratio)) return -3;
    if (!st_ok(tvar, lt)) return -6;
    if (M < 0) return -9;
    if (M == 0) return 0;
    if (!Xs3) return -8;
    if (out_stride < M) return -10;
    if (!mean) return -12;
    if (!var) return -13;
    FitLayout L = fit_layout(N, 3);
    if (!ws) return -14;
    if (ws_bytes < predict_panel_bytes(L.npad)) return -15;
    return cuda_rc(predict_helm(fit_ws, L, N, make_helm_st(l_df, l_cf, ratio, tvar, lt), Xs3, M, (long)out_stride, var_add,
                                mean, var, ws, ws_bytes, (cudaStream_t)stream));
}

int gp2d_st_lml_grad(const double* X3, int N, const double* y, double l_df, double l_cf, double ratio, double tvar,
                     double lt, double noise, double jitter, void* ws, size_t ws_bytes, double* out7, int* info,
                     void* stream) {
    if (!X3) return -1;
    if (!size_ok(2L * N)) return -2;
    if (!y) return -3;
    if (!theta_ok(l_df, l_cf, ratio)) return -4;
    if (!st_ok(tvar, lt)) return -7;
    if (!(noise >= 0.0)) return -9;
    if (!(jitter >= 0.0)) return -10;
    FitLayout L = fit_layout(N, 3);
    if (!ws) return -11;
    if (ws_bytes < L.total) return -12;
    if (!out7) return -13;
    cudaStream_t st = (cudaStream_t)stream;
    HelmParams hp = make_helm_st(l_df, l_cf, ratio, tvar, lt);
    cudaError_t e = fit_core(X3, N, y, hp, noise + jitter, ws, L, st);
    if (e != cudaSuccess) return cuda_rc(e);
    double* Z = at<double>(ws, L.off_Z);
    double* Kinv = at<double>(ws, L.off_A);
    e = launch_dgemm(true, true, GemmArgs{Z, L.npad, Z, L.npad, Kinv, L.npad, L.npad, L.npad, L.npad, 1.0, 0.0, 1, KR_GE_M}, st);
    if (e != cudaSuccess) return cuda_rc(e);
    double* scal = at<double>(ws, L.off_scal);
    e = lml_grad_reduce(Kinv, L.npad, L.npad, at<double>(ws, L.off_alpha), at<double>(ws, L.off_X), N, hp, false,
                        at<double>(ws, L.off_partial), scal + 1, st);
    if (e != cudaSuccess) return cuda_rc(e);
    e = cudaMemcpyAsync(out7, scal, 7 * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_rc(e);
    if (info) {
        e = cudaMemcpyAsync(info, at<int>(ws, L.off_info), sizeof(int), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    return 0;
}

/* ---- scalar ARD-RBF sum family --------------------------------------------------------------- */

namespace {
cudaError_t rbf_fit_core(const double* X, int N, const double* y, const RbfParams& rp, double diag_add, void* ws,
                         const FitLayout& L, cudaStream_t st) {
    double* A = at<double>(ws, L.off_A);
    double* Z = at<double>(ws, L.off_Z);
    cudaError_t e;
    e = cudaMemcpyAsync(at<double>(ws, L.off_X), X, (size_t)N * rp.D * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
    e = rbf_build_padded_lower(X, N, rp, diag_add, A, L.npad, L.npad, st);
    if (e != cudaSuccess) return e;
    const int refine = refine_steps_for(rp.kss, (long)N, diag_add);
    e = factor_ws(ws, L, refine, st);
    if (e != cudaSuccess) return e;
    e = pack_lower_tiles(Z, L.npad, L.npad, at<double>(ws, L.off_Zt), st);
    if (e != cudaSuccess) return e;
    e = solve_alpha_lml(Z, L.npad, L.npad, N, 1, y, at<double>(ws, L.off_yint), at<double>(ws, L.off_w),
                        at<double>(ws, L.off_alpha), at<double>(ws, L.off_partial),
                        at<double>(ws, L.off_logdiag), at<double>(ws, L.off_scal), st);
    if (e != cudaSuccess || !refine) return e;
    e = rbf_build_padded_lower(X, N, rp, diag_add, A, L.npad, L.npad, st);
    if (e != cudaSuccess) return e;
    return refine_alpha_ws(ws, L, st);
}
}  // namespace

int gp2d_rbf_kernel_build(const double* X, int N, const double* X2, int M, int D, int Q, const double* var,
                          const double* ls, double diag_add, double* K, int64_t ldk, void* stream) {
    if (!X) return -1;
    if (N < 0) return -2;
    if (M < 0 || (X2 == nullptr && M != N)) return -4;
    RbfParams rp;
    if (!make_rbf(D, Q, var, ls, &rp)) return -5;
    if (!K && N > 0 && M > 0) return -10;
    if (ldk < M) return -11;
    return cuda_rc(rbf_build(X, N, X2, M, rp, diag_add, K, (long)ldk, (cudaStream_t)stream));
}

size_t gp2d_rbf_kernel_grad_workspace_bytes(int N, int M) {
    if (N <= 0 || M <= 0) return 256;
    return align256(sizeof(double) * RBF_NG * ((size_t)rbf_grad_partials(N, M) + 1));
}

int gp2d_rbf_kernel_grad(const double* X, int N, const double* X2, int M, int D, int Q, const double* var,
                         const double* ls, const double* dL_dK, int64_t ld, void* ws, size_t ws_bytes,
                         double* out, void* stream) {
    if (!X) return -1;
    if (N <= 0) return -2;
    if (M <= 0 || (X2 == nullptr && M != N)) return -4;
    RbfParams rp;
    if (!make_rbf(D, Q, var, ls, &rp)) return -5;
    if (!dL_dK) return -9;
    if (ld < M) return -10;
    if (!ws || ws_bytes < gp2d_rbf_kernel_grad_workspace_bytes(N, M)) return -12;
    if (!out) return -13;
    return cuda_rc(rbf_grad_sums(X, N, X2, M, rp, dL_dK, (long)ld, (double*)ws,
                                 (int)(ws_bytes / (RBF_NG * sizeof(double))), out, (cudaStream_t)stream));
}

size_t gp2d_rbf_fit_workspace_bytes(int N, int D) {
    if (!size_ok(N) || D < 1 || D > RBF_MAXD) return 0;
    return rbf_layout(N, D).total;
}

int gp2d_rbf_fit_predict_state(int N, int D, size_t* offset, size_t* bytes) {
    if (N <= 0) return -1;
    if (D < 1 || D > RBF_MAXD) return -2;
    if (!offset) return -3;
    if (!bytes) return -4;
    FitLayout L = rbf_layout(N, D);
    *offset = L.off_Zt;
    *bytes = L.total - L.off_Zt;
    return 0;
}

int gp2d_rbf_fit(const double* X, int N, int D, const double* y, int Q, const double* var, const double* ls,
                 double noise, double jitter, void* ws, size_t ws_bytes, double* alpha_out, double* lml_out,
                 int* info, void* stream) {
    if (!X) return -1;
    if (!size_ok(N)) return -2;
    if (!y) return -4;
    RbfParams rp;
    if (!make_rbf(D, Q, var, ls, &rp)) return -5;
    if (!(noise >= 0.0)) return -8;
    if (!(jitter >= 0.0)) return -9;
    FitLayout L = rbf_layout(N, D);
    if (!ws) return -10;
    if (ws_bytes < L.total) return -11;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = rbf_fit_core(X, N, y, rp, noise + jitter, ws, L, st);
    if (e != cudaSuccess) return cuda_rc(e);
    if (alpha_out) {
        e = cudaMemcpyAsync(alpha_out, at<double>(ws, L.off_alpha), (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    if (lml_out) {
        e = cudaMemcpyAsync(lml_out, at<double>(ws, L.off_scal), sizeof(double), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    if (info) {
        e = cudaMemcpyAsync(info, at<int>(ws, L.off_info), sizeof(int), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    return 0;
}

size_t gp2d_rbf_predict_workspace_bytes(int N, int M) {
    if (N <= 0 || M <= 0) return 256;
    return align256(predict_scratch_bytes(round_up(N, TILE), M, 128));
}

int gp2d_rbf_predict(const void* fit_ws, int N, int D, int Q, const double* var, const double* ls,
                     const double* Xs, int M, double var_add, double* mean, double* variance, void* ws,
                     size_t ws_bytes, void* stream) {
    if (!fit_ws) return -1;
    if (!size_ok(N)) return -2;
    RbfParams rp;
    if (!make_rbf(D, Q, var, ls, &rp)) return -3;
    if (M < 0) return -8;
    if (M == 0) return 0;
    if (!Xs) return -7;
    if (!mean) return -10;
    if (!variance) return -11;
    FitLayout L = rbf_layout(N, D);
    if (!ws) return -12;
    if (ws_bytes < predict_panel_bytes(L.npad)) return -13;
    return cuda_rc(predict_fused_rbf(at<double>(fit_ws, L.off_Zt), L.npad, at<double>(fit_ws, L.off_alpha),
                                     at<double>(fit_ws, L.off_X), N, rp, Xs, M, var_add, mean, variance,
                                     (double*)ws, ws_bytes, (cudaStream_t)stream));
}

int gp2d_rbf_lml_grad(const double* X, int N, int D, const double* y, int Q, const double* var, const double* ls,
                      double noise, double jitter, void* ws, size_t ws_bytes, double* out, int* info, void* stream) {
    if (!X) return -1;
    if (!size_ok(N)) return -2;
    if (!y) return -4;
    RbfParams rp;
    if (!make_rbf(D, Q, var, ls, &rp)) return -5;
    if (!(noise >= 0.0)) return -8;
    if (!(jitter >= 0.0)) return -9;
    FitLayout L = rbf_layout(N, D);
    if (!ws) return -10;
    if (ws_bytes < L.total) return -11;
    if (!out) return -12;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = rbf_fit_core(X, N, y, rp, noise + jitter, ws, L, st);
    if (e != cudaSuccess) return cuda_rc(e);
    double* Z = at<double>(ws, L.off_Z);
    double* Kinv = at<double>(ws, L.off_A);
    e = launch_dgemm(true, true, GemmArgs{Z, L.npad, Z, L.npad, Kinv, L.npad, L.npad, L.npad, L.npad, 1.0, 0.0, 1, KR_GE_M}, st);
    if (e != cudaSuccess) return cuda_rc(e);
    double* scal = at<double>(ws, L.off_scal);
    e = rbf_lml_grad_reduce(Kinv, L.npad, L.npad, at<double>(ws, L.off_alpha), at<double>(ws, L.off_X), N, rp,
                            at<double>(ws, L.off_partial), scal + 1, st);
    if (e != cudaSuccess) return cuda_rc(e);
    e = cudaMemcpyAsync(out, scal, (size_t)(2 + Q * (1 + D)) * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_rc(e);
    if (info) {
        e = cudaMemcpyAsync(info, at<int>(ws, L.off_info), sizeof(int), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    return 0;
}

/* ---- sum of space-time Helmholtz terms (hsum.cuh; krig.py:396-407) ------------------------------- */

namespace {
FitLayout hsum_layout(int N, int ldx, int Q) {
    return fit_layout_general(2 * (size_t)N, (size_t)ldx * N, hsum_lml_grad_partial_doubles(round_up(2 * N, TILE), Q));
}

cudaError_t hsum_fit_core(const double* X, int N, const double* y, const HsumParams& sp, double diag_add, void* ws,
                          const FitLayout& L, cudaStream_t st) {
    double* A = at<double>(ws, L.off_A);
    double* Z = at<double>(ws, L.off_Z);
    cudaError_t e;
    e = cudaMemcpyAsync(at<double>(ws, L.off_X), X, (size_t)sp.ldx * N * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
    e = hsum_build_interleaved_lower(X, N, sp, diag_add, A, L.npad, L.npad, st);
    if (e != cudaSuccess) return e;
    const int refine = refine_steps_for(sp.kss0 > sp.kss1 ? sp.kss0 : sp.kss1, 2L * N, diag_add);
    e = factor_ws(ws, L, refine, st);
    if (e != cudaSuccess) return e;
    e = pack_lower_tiles(Z, L.npad, L.npad, at<double>(ws, L.off_Zt), st);
    if (e != cudaSuccess) return e;
    e = solve_alpha_lml(Z, L.npad, L.npad, N, 2, y, at<double>(ws, L.off_yint), at<double>(ws, L.off_w),
                        at<double>(ws, L.off_alpha), at<double>(ws, L.off_partial),
                        at<double>(ws, L.off_logdiag), at<double>(ws, L.off_scal), st);
    if (e != cudaSuccess || !refine) return e;
    e = hsum_build_interleaved_lower(X, N, sp, diag_add, A, L.npad, L.npad, st);
    if (e != cudaSuccess) return e;
    return refine_alpha_ws(ws, L, st);
}
}  // namespace

int gp2d_hsum_kernel_build(const double* X, int N, const double* X2, int M, int ldx, int Q, const int* type,
                           const double* params, double diag_add, double* K, int64_t ldk, void* stream) {
    if (!X) return -1;
    if (N < 0) return -2;
    if (M < 0 || (X2 == nullptr && M != N)) return -4;
    HsumParams sp;
    if (!make_hsum(ldx, Q, type, params, &sp)) return -5;
    if (!K && N > 0 && M > 0) return -10;
    if (ldk < 2 * (int64_t)M) return -11;
    return cuda_rc(hsum_build_block_layout(X, N, X2, M, sp, diag_add, K, (long)ldk, (cudaStream_t)stream));
}

int gp2d_hsum_kdiag(int M, int ldx, int Q, const int* type, const double* params, double* out, void* stream) {
    if (M < 0) return -1;
    HsumParams sp;
    if (!make_hsum(ldx, Q, type, params, &sp)) return -2;
    if (M == 0) return 0;
    if (!out) return -6;
    cudaStream_t st = (cudaStream_t)stream;
    fill_kernel<<<(M + 255) / 256, 256, 0, st>>>(out, M, sp.kss0);
    fill_kernel<<<(M + 255) / 256, 256, 0, st>>>(out + M, M, sp.kss1);
    return cuda_rc(cudaGetLastError());
}

size_t gp2d_hsum_kernel_grad_workspace_bytes(int N, int M, int Q) {
    if (N <= 0 || M <= 0 || Q < 1 || Q > HSUM_MAXQ) return 256;
    return align256(sizeof(double) * HSUM_NP * (size_t)Q * (size_t)grad_sums_block_partials(N, M));
}

int gp2d_hsum_kernel_grad(const double* X, int N, const double* X2, int M, int ldx, int Q, const int* type,
                          const double* params, const double* dL_dK, int64_t ld, void* ws, size_t ws_bytes,
                          double* out, void* stream) {
    if (!X) return -1;
    if (N <= 0) return -2;
    if (M <= 0 || (X2 == nullptr && M != N)) return -4;
    HsumParams sp;
    if (!make_hsum(ldx, Q, type, params, &sp)) return -5;
    if (!dL_dK) return -9;
    if (ld < 2 * (int64_t)M) return -10;
    if (!ws || ws_bytes < gp2d_hsum_kernel_grad_workspace_bytes(N, M, Q)) return -12;
    if (!out) return -13;
    return cuda_rc(hsum_grad_sums(X, N, X2, M, sp, dL_dK, (long)ld, (double*)ws, ws_bytes / sizeof(double), out,
                                  (cudaStream_t)stream));
}

size_t gp2d_hsum_fit_workspace_bytes(int N, int ldx, int Q) {
    if (!size_ok(2L * N) || (ldx != 2 && ldx != 3) || Q < 1 || Q > HSUM_MAXQ) return 0;
    return hsum_layout(N, ldx, Q).total;
}

int gp2d_hsum_fit_predict_state(int N, int ldx, int Q, size_t* offset, size_t* bytes) {
    if (N <= 0) return -1;
    if (ldx != 2 && ldx != 3) return -2;
    if (Q < 1 || Q > HSUM_MAXQ) return -3;
    if (!offset) return -4;
    if (!bytes) return -5;
    FitLayout L = hsum_layout(N, ldx, Q);
    *offset = L.off_Zt;
    *bytes = L.total - L.off_Zt;
    return 0;
}

int gp2d_hsum_fit(const double* X, int N, int ldx, const double* y, int Q, const int* type, const double* params,
                  double noise, double jitter, void* ws, size_t ws_bytes, double* alpha_out, double* lml_out,
                  int* info, void* stream) {
    if (!X) return -1;
    if (!size_ok(2L * N)) return -2;
    if (!y) return -4;
    HsumParams sp;
    if (!make_hsum(ldx, Q, type, params, &sp)) return -5;
    if (!(noise >= 0.0)) return -8;
    if (!(jitter >= 0.0)) return -9;
    FitLayout L = hsum_layout(N, ldx, Q);
    if (!ws) return -10;
    if (ws_bytes < L.total) return -11;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = hsum_fit_core(X, N, y, sp, noise + jitter, ws, L, st);
    if (e != cudaSuccess) return cuda_rc(e);
    if (alpha_out) {
        e = deinterleave(at<double>(ws, L.off_alpha), N, alpha_out, st, 1, 0, perm_of(ws, L), 0);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    if (lml_out) {
        e = cudaMemcpyAsync(lml_out, at<double>(ws, L.off_scal), sizeof(double), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    if (info) {
        e = cudaMemcpyAsync(info, at<int>(ws, L.off_info), sizeof(int), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    return 0;
}

int gp2d_hsum_predict(const void* fit_ws, int N, int ldx, int Q, const int* type, const double* params,
                      const double* Xs, int M, int64_t out_stride, double var_add, double* mean, double* var,
                      void* ws, size_t ws_bytes, void* stream) {
    if (!fit_ws) return -1;
    if (!size_ok(2L * N)) return -2;
    HsumParams sp;
    if (!make_hsum(ldx, Q, type, params, &sp)) return -3;
    if (M < 0) return -8;
    if (M == 0) return 0;
    if (!Xs) return -7;
    if (out_stride < M) return -9;
    if (!mean) return -11;
    if (!var) return -12;
    FitLayout L = hsum_layout(N, ldx, Q);
    if (!ws) return -13;
    if (ws_bytes < predict_panel_bytes(L.npad)) return -14;
    return cuda_rc(predict_fused_hsum(at<double>(fit_ws, L.off_Zt), L.npad, at<double>(fit_ws, L.off_alpha),
                                      at<double>(fit_ws, L.off_X), N, sp, Xs, M, (long)out_stride, var_add, mean, var,
                                      (double*)ws, ws_bytes, (cudaStream_t)stream));
}

int gp2d_hsum_lml_grad(const double* X, int N, int ldx, const double* y, int Q, const int* type, const double* params,
                       double noise, double jitter, void* ws, size_t ws_bytes, double* out, int* info, void* stream) {
    if (!X) return -1;
    if (!size_ok(2L * N)) return -2;
    if (!y) return -4;
    HsumParams sp;
    if (!make_hsum(ldx, Q, type, params, &sp)) return -5;
    if (!(noise >= 0.0)) return -8;
    if (!(jitter >= 0.0)) return -9;
    FitLayout L = hsum_layout(N, ldx, Q);
    if (!ws) return -10;
    if (ws_bytes < L.total) return -11;
    if (!out) return -12;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = hsum_fit_core(X, N, y, sp, noise + jitter, ws, L, st);
    if (e != cudaSuccess) return cuda_rc(e);
    double* Z = at<double>(ws, L.off_Z);
    double* Kinv = at<double>(ws, L.off_A);
    e = launch_dgemm(true, true, GemmArgs{Z, L.npad, Z, L.npad, Kinv, L.npad, L.npad, L.npad, L.npad, 1.0, 0.0, 1, KR_GE_M}, st);
    if (e != cudaSuccess) return cuda_rc(e);
    double* scal = at<double>(ws, L.off_scal);
    // the gradient goes straight to the caller's buffer: out = (LML, d/d(var, lt, la, lb)_q ..., d/dnoise)
    e = hsum_lml_grad_reduce(Kinv, L.npad, L.npad, at<double>(ws, L.off_alpha), at<double>(ws, L.off_X), N, sp,
                             at<double>(ws, L.off_partial), out + 1, st);
    if (e != cudaSuccess) return cuda_rc(e);
    e = cudaMemcpyAsync(out, scal, sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return cuda_rc(e);
    if (info) {
        e = cudaMemcpyAsync(info, at<int>(ws, L.off_info), sizeof(int), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    return 0;
}

namespace {
// Device buffer of the host-pointer entry points, cached per host thread: cudaMalloc / cudaFree of the
// fit workspace cost as much as the whole configs[0] pipeline (2.0 ms against 1.04 ms).  Grown on
// demand, released at thread exit or by gp2d_host_release(); nothing else is kept between calls.
struct HostCache {
    char* dev = nullptr;
    size_t cap = 0;
    int device = -1;
    cudaError_t reserve(size_t bytes) {
        int cur = 0;
        cudaError_t e = cudaGetDevice(&cur);
        if (e != cudaSuccess) return e;
        if (dev && cur == device && cap >= bytes) return cudaSuccess;
        release();
        e = cudaMalloc(&dev, bytes);
        if (e != cudaSuccess) { dev = nullptr; return e; }
        cap = bytes;
        device = cur;
        return cudaSuccess;
    }
    void release() {
        if (dev) {
            int cur = 0;
            const bool sw = cudaGetDevice(&cur) == cudaSuccess && cur != device;
            if (sw) cudaSetDevice(device);
            cudaFree(dev);
            if (sw) cudaSetDevice(cur);
        }
        dev = nullptr; cap = 0; device = -1;
    }
    ~HostCache() { release(); }       // at process exit the context may be gone: errors ignored
};
thread_local HostCache g_host_cache;
}  // namespace

void gp2d_host_release(void) { g_host_cache.release(); }

int gp2d_fit_predict_host(const double* X, int N, const double* y, double l_df, double l_cf,
                          double ratio, double noise, double jitter, const double* Xs, int M,
                          int include_noise, double* mean, double* var, double* lml) {
    if (!X) return -1;
    if (!size_ok(2L * N)) return -2;
    if (!y) return -3;
    if (!theta_ok(l_df, l_cf, ratio)) return -4;
    if (M < 0) return -10;
    if (M > 0 && (!Xs || !mean || !var)) return -9;
    const size_t wsb = gp2d_fit_workspace_bytes(N);
    const size_t pwsb = gp2d_predict_workspace_bytes(N, M);
    const size_t in_b = align256(2 * (size_t)N * 8) * 2 + align256(2 * (size_t)(M > 0 ? M : 1) * 8);
    const size_t out_b = align256(2 * (size_t)(M > 0 ? M : 1) * 8) * 2 + 256;
    cudaError_t e = g_host_cache.reserve(wsb + in_b + out_b + pwsb);
    if (e != cudaSuccess) return cuda_rc(e);
    char* dev = g_host_cache.dev;
    char* q = dev + wsb;
    double* dX = (double*)q; q += align256(2 * (size_t)N * 8);
    double* dy = (double*)q; q += align256(2 * (size_t)N * 8);
    double* dXs = (double*)q; q += align256(2 * (size_t)(M > 0 ? M : 1) * 8);
    double* dmean = (double*)q; q += align256(2 * (size_t)(M > 0 ? M : 1) * 8);
    double* dvar = (double*)q; q += align256(2 * (size_t)(M > 0 ? M : 1) * 8);
    double* dlml = (double*)q; int* dinfo = (int*)(q + 64); q += 256;
    void* dpws = (void*)q;
    int rc = 0, info = 0;
    cudaStream_t st = 0;
    do {
        if ((e = cudaMemcpyAsync(dX, X, 2 * (size_t)N * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(dy, y, 2 * (size_t)N * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        if (M > 0 && (e = cudaMemcpyAsync(dXs, Xs, 2 * (size_t)M * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        rc = gp2d_fit(dX, N, dy, l_df, l_cf, ratio, noise, jitter, dev, wsb, nullptr, dlml, dinfo, st);
        if (rc) break;
        rc = gp2d_predict(dev, N, l_df, l_cf, ratio, dXs, M, M, include_noise ? noise : 0.0, dmean, dvar, dpws, pwsb, st);
        if (rc) break;
        if (M > 0) {
            if ((e = cudaMemcpyAsync(mean, dmean, 2 * (size_t)M * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
            if ((e = cudaMemcpyAsync(var, dvar, 2 * (size_t)M * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        }
        if (lml && (e = cudaMemcpyAsync(lml, dlml, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        if ((e = cudaMemcpyAsync(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        e = cudaStreamSynchronize(st);
    } while (0);
    if (rc) return rc;
    if (e != cudaSuccess) return cuda_rc(e);
    return info;
}

/* ---- bring-up / test hooks (not part of include/gp2d.h) ---------------------------------- */
// Register-resident FP64 loops: the tensor-pipe ceiling bench.py reports against
// (MEASURED_PEAKS.json carries no fp64 figure).  mode 0: DMMA.8x8x4, flops = ctas * 8 warps *
// iters * 16 * 512; mode 1: DFMA, flops = ctas * 256 threads * iters * 32 * 2; mode 2: both.
int gp2d_dbg_fp64_peak(int iters, int ctas, double* out, void* stream) {
    fp64_peak_kernel<0><<<ctas, NTHREADS, 0, (cudaStream_t)stream>>>(out, iters);
    return cuda_rc(cudaGetLastError());
}
int gp2d_dbg_fp64_mode(int mode, int iters, int ctas, double* out, void* stream) {
    if (mode == 0) fp64_peak_kernel<0><<<ctas, NTHREADS, 0, (cudaStream_t)stream>>>(out, iters);
    else if (mode == 1) fp64_peak_kernel<1><<<ctas, NTHREADS, 0, (cudaStream_t)stream>>>(out, iters);
    else if (mode == 2) fp64_peak_kernel<2><<<ctas, NTHREADS, 0, (cudaStream_t)stream>>>(out, iters);
    else if (mode == 3) fp64_peak_kernel<3><<<ctas, 32, 0, (cudaStream_t)stream>>>(out, iters);
    else fp64_peak_kernel<4><<<ctas, 32, 0, (cudaStream_t)stream>>>(out, iters);
    return cuda_rc(cudaGetLastError());
}

int gp2d_dbg_set_small_tile_threshold(int t) { set_small_tile_threshold(t); return t; }

int gp2d_dbg_set_robust_cond(double c) { g_robust_cond = c; return 0; }

int gp2d_dbg_set_potri_overlap(int on) { set_potri_overlap(on != 0); return on; }

int gp2d_dbg_set_predict_split(int s) { set_predict_split(s); return s; }
int gp2d_dbg_set_i8(int v) { set_i8_debug(v); return v; }

// the permutation the last fit in this workspace (Helmholtz / space-time families) applied to its observations:
// perm_out[i] = caller's index of internal observation i (device pointer, N ints); returns 0 when the fit keeps the
// caller's order (perm_out untouched), N otherwise
int gp2d_dbg_fit_order(const void* fit_ws, int N, int ldx, int* perm_out, void* stream) {
    if (!fit_ws || !perm_out || N <= 0 || (ldx != 2 && ldx != 3)) return -1;
    const FitLayout L = fit_layout(N, ldx);
    if (!L.order_n) return 0;
    cudaError_t e = cudaMemcpyAsync(perm_out, at<int>(fit_ws, L.off_perm), (size_t)N * sizeof(int), cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream);
    return e == cudaSuccess ? N : cuda_rc(e);
}

int gp2d_dbg_gemm(int a_mn, int b_mn, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                  int64_t ldc, int M, int N, int K, double alpha, double beta, int lower_out, int krule,
                  void* stream) {
    return cuda_rc(launch_dgemm(a_mn != 0, b_mn != 0,
                                GemmArgs{A, (long)lda, B, (long)ldb, C, (long)ldc, M, N, K, alpha, beta, lower_out, krule},
                                (cudaStream_t)stream));
}

int gp2d_dbg_potri(double* A, int n, double* Z, double* logdiag, int* info, int need_inv, int keep_L,
                   double* W, void* stream) {
    return cuda_rc(potri_lower(A, n, Z, n, n, logdiag, info, need_inv != 0, keep_L != 0, W, (cudaStream_t)stream));
}

}  // extern "C"
