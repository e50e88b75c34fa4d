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
template <bool A_MN, bool B_MN>
static cudaError_t gemm_attr64() {
    return cudaFuncSetAttribute(dgemm_kernel<A_MN, B_MN, 128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                gemm_smem_bytes(64));
}
template <bool A_MN, bool B_MN>
static cudaError_t gemm_attr_ws() {
    return cudaFuncSetAttribute(dgemm_ws_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GEMM_WS_SMEM_BYTES);
}

cudaError_t dgemm_init() {
    static PerDeviceOnce once;
    const int slot = once.pending();
    if (slot < 0) return cudaSuccess;
    cudaError_t e;
    if ((e = gemm_attr64<false, false>()) != cudaSuccess) return e;
    if ((e = gemm_attr64<false, true>()) != cudaSuccess) return e;
    if ((e = gemm_attr64<true, true>()) != cudaSuccess) return e;
    if ((e = gemm_attr64<true, false>()) != cudaSuccess) return e;
    if ((e = gemm_attr_ws<false, false>()) != cudaSuccess) return e;
    if ((e = gemm_attr_ws<false, true>()) != cudaSuccess) return e;
    if ((e = gemm_attr_ws<true, true>()) != cudaSuccess) return e;
    if ((e = gemm_attr_ws<true, false>()) != cudaSuccess) return e;
    once.done[slot] = true;
    return cudaSuccess;
}

static void gemm_launch64(bool a_mn, bool b_mn, unsigned tiles, const GemmArgs& a, cudaStream_t st) {
    constexpr int SM = gemm_smem_bytes(64);
    const dim3 grid(tiles, (unsigned)a.batch);
    if (!a_mn && !b_mn) dgemm_kernel<false, false, 128, 64><<<grid, 128, SM, st>>>(a);
    else if (!a_mn && b_mn) dgemm_kernel<false, true, 128, 64><<<grid, 128, SM, st>>>(a);
    else if (a_mn && b_mn) dgemm_kernel<true, true, 128, 64><<<grid, 128, SM, st>>>(a);
    else dgemm_kernel<true, false, 128, 64><<<grid, 128, SM, st>>>(a);
}
static void gemm_launch_ws(bool a_mn, bool b_mn, unsigned tiles, const GemmArgs& a, cudaStream_t st) {
    constexpr int SM = GEMM_WS_SMEM_BYTES;
    const dim3 grid(tiles, (unsigned)a.batch);
    if (!a_mn && !b_mn) dgemm_ws_kernel<false, false><<<grid, WS_THREADS, SM, st>>>(a);
    else if (!a_mn && b_mn) dgemm_ws_kernel<false, true><<<grid, WS_THREADS, SM, st>>>(a);
    else if (a_mn && b_mn) dgemm_ws_kernel<true, true><<<grid, WS_THREADS, SM, st>>>(a);
    else dgemm_ws_kernel<true, false><<<grid, WS_THREADS, SM, st>>>(a);
}

// Tile shape per launch.  Measured per 16-deep k-step on B200 (tools/explore.py gemm_small):
//   128-tile warp-specialised kernel: 2.23 us per wave of 148 CTAs (one CTA per SM);
//   64-tile kernel: 0.82 / 1.20 / 1.65 us with 1 / 2 / 3 CTAs resident per SM, 1.65 us per wave of
//   3 x 148 beyond that.
// The cheaper estimate wins: small problems and awkward wave counts (e.g. 300 tiles = 2.03 waves)
// go to the 64-tile kernel, everything else to the 128-tile one.  A non-negative override forces
// 128-tiles at or above that many 128-tiles (0: always, huge: never) -- tests use it.
static thread_local int g_small_tile_threshold = -1;      // bring-up override, per host thread
void set_small_tile_threshold(int t) { g_small_tile_threshold = t; }

static int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

static bool use_tile64(long t128, long t64) {
    if (g_small_tile_threshold >= 0) return t128 < g_small_tile_threshold;
    const long sms = num_sms();
    const double est128 = 2.23 * (double)((t128 + sms - 1) / sms);
    const long c = (t64 + sms - 1) / sms;
    const double est64 = c <= 1 ? 0.82 : c == 2 ? 1.20 : c == 3 ? 1.65 : 1.65 * (double)((t64 + 3 * sms - 1) / (3 * sms));
    return est64 < est128;
}

cudaError_t launch_dgemm(bool a_mn, bool b_mn, const GemmArgs& a, cudaStream_t st) {
    if (a.M % TILE || a.N % TILE || a.K % BK || a.M <= 0 || a.N <= 0 || a.K <= 0) return cudaErrorInvalidValue;
    if ((a.lda & 1) || (a.ldb & 1) || (a.ldc & 1)) return cudaErrorInvalidValue;
    cudaError_t e = dgemm_init();
    if (e != cudaSuccess) return e;
    const long tm = a.M / TILE, tn = a.N / TILE;
    if (a.lower_out && tm != tn) return cudaErrorInvalidValue;
    if (a.batch < 1 || a.batch > 65535) return cudaErrorInvalidValue;
    const long t128 = a.lower_out ? tm * (tm + 1) / 2 : tm * tn;
    const long t64 = a.lower_out ? (2 * tm) * (2 * tm + 1) / 2 : 4 * tm * tn;
    // the tile shape is chosen for the whole batch; both shapes accumulate every output element over k in
    // the same order with the same DMMA shape, so the choice never changes a result bit
    if (use_tile64(t128 * a.batch, t64 * a.batch)) gemm_launch64(a_mn, b_mn, (unsigned)t64, a, st);
    else gemm_launch_ws(a_mn, b_mn, (unsigned)t128, a, st);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// 128x128 leaf: L = chol(A_kk) written back to A (lower), Z_kk = L^-1 (full tile, zeros
// above the diagonal), log(diag L) and LAPACK-style info (1-based index of the first
// non-positive pivot, first failure wins).
// ------------------------------------------------------------------------------------
// One CTA, 256 threads as a 16 x 16 grid; thread (tr, tc) keeps the 8 x 8 elements
// (i, c) = (tr + 16 a, tc + 16 b) in registers (2-D cyclic, so the shrinking active window
// stays balanced).  A single in-place sweep does both jobs, in UNSCALED form (LDL^T-like):
// with d_j the pivot and q = column j of the Schur complement, at step j every row i > j gets
//     M[i,c] -= (q_i / d_j) * m_c,   m_c = q_c       (c > j : Cholesky trailing update)
//                                    m_c = Y[j,c]    (c <= j: inverse accumulator, Y[j,j] = 1)
// and the slots of column j are recycled for the inverse (Y[:,j] starts at 0).  At the end
//     L[i,j] = q_i / sqrt(d_j),   Z[i,c] = Y[i,c] / sqrt(d_i)          (Z = L^-1).
// Per step the critical path is one barrier, one reciprocal and one rank-1 update from
// registers: the column is broadcast through shared memory, the row comes from the owning
// lane of the same half-warp by shuffle (tc = tid>>4, tr = tid&15), every register index is
// static (outer loop over j>>4 unrolled), and all sqrt/log work is deferred.
constexpr int LEAF_LD = TILE + 1;
constexpr int LEAF_SMEM_BYTES = TILE * LEAF_LD * (int)sizeof(double);

// Reciprocal of a positive pivot: hardware seed (rcp.approx.ftz.f64, ~20 bits) and two Newton steps (error 2^-40, then
// below double rounding).  The reciprocal heads the dependent chain of every column step of the leaf; the IEEE division
// the compiler emits for 1.0 / d is twice as long.  NaN stays NaN (a non-positive pivot has been turned into one).
__device__ __forceinline__ double pivot_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
}

__global__ void __launch_bounds__(NTHREADS, 1)
potri_leaf_kernel(double* A, long lda, double* Z, long ldz, double* logdiag, int* info, int row0, int keep_L,
                  long bstride) {
    // batch: problem blockIdx.x lives bstride doubles (2 bstride ints) further on in every array
    A += (long)blockIdx.x * bstride;
    Z += (long)blockIdx.x * bstride;
    logdiag += (long)blockIdx.x * bstride;
    info += (long)blockIdx.x * 2 * bstride;
    extern __shared__ double S[];                 // [128][129] staging: A in, q/L out, Z out
    __shared__ double colbuf[2][TILE];
    __shared__ double dsave[TILE];                // pivots d_j, then 1/sqrt(d_j)
    __shared__ int first_bad;
    const int tid = threadIdx.x, lane = tid & 31;
    const int tc = tid >> 4, tr = tid & 15;
    if (tid == 0) first_bad = TILE;
    // stage the tile through shared memory, 8 independent loads in flight per thread (one load per
    // iteration serialised the prologue on the global-memory latency: 11 % of the kernel in the profile)
    for (int idx0 = tid; idx0 < TILE * TILE; idx0 += 8 * NTHREADS) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = idx0 + u * NTHREADS, r = idx >> 7, c = idx & 127;
            v[u] = (c <= r) ? A[(long)r * lda + c] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = idx0 + u * NTHREADS, r = idx >> 7, c = idx & 127;
            S[r * LEAF_LD + c] = v[u];
        }
    }
    __syncthreads();
    double M[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) M[a][b] = S[(tr + 16 * a) * LEAF_LD + tc + 16 * b];
    __syncthreads();
#pragma unroll
    for (int ja = 0; ja < 8; ++ja) {
        for (int jr = 0; jr < 16; ++jr) {
            const int j = 16 * ja + jr;
            double* cb = colbuf[j & 1];
            if (tc == jr) {
                // publish column j (raw), keep it for L, recycle the slots for Y[:,j]
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    const int i = tr + 16 * a;
                    cb[i] = M[a][ja];
                    S[i * LEAF_LD + j] = M[a][ja];
                    M[a][ja] = (i == j) ? 1.0 : 0.0;
                }
            }
            __syncthreads();
            double d = cb[j];
            if (tid == 0) dsave[j] = d;
            if (!(d > 0.0)) d = nan("");
            const double rd = pivot_rcp(d);
            // row multipliers from the lane that owns row j in this half-warp (unscaled Y[j,c])
            const int src = (lane & 16) | jr;
            double m[8], l[8];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const double y = __shfl_sync(0xffffffffu, M[ja][b], src);
                if (b < ja) m[b] = y;
                else if (b > ja) m[b] = cb[tc + 16 * b];
                else m[b] = (tc <= jr) ? y : cb[tc + 16 * b];
            }
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const bool active = (a > ja) || (a == ja && tr > jr);
                l[a] = active ? cb[tr + 16 * a] * rd : 0.0;
            }
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                if (a < ja) continue;
#pragma unroll
                for (int b = 0; b < 8; ++b) M[a][b] = fma(-l[a], m[b], M[a][b]);
            }
        }
    }
    __syncthreads();
    if (tid < TILE) {
        const double d = dsave[tid];
        logdiag[tid] = 0.5 * log(d);
        dsave[tid] = (d > 0.0) ? rsqrt(d) : nan("");
        if (!(d > 0.0)) atomicMin(&first_bad, tid);
    }
    __syncthreads();
    // LAPACK info: first non-positive pivot, unless an earlier block already failed
    if (tid == 0 && first_bad < TILE) atomicCAS(info, 0, row0 + first_bad + 1);
    // L = q * diag(1/sqrt(d)) back to A (lower), Z = diag(1/sqrt(d)) * Y
    for (int idx = tid; idx < TILE * TILE; idx += NTHREADS) {
        const int r = idx >> 7, c = idx & 127;
        // with keep_L the strict upper triangle is cleared too: the refinement GEMMs read whole diagonal tiles of L
        if (c <= r) A[(long)r * lda + c] = S[r * LEAF_LD + c] * dsave[c];
        else if (keep_L) A[(long)r * lda + c] = 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const double sc = dsave[tr + 16 * a];
#pragma unroll
        for (int b = 0; b < 8; ++b) S[(tr + 16 * a) * LEAF_LD + tc + 16 * b] = M[a][b] * sc;
    }
    __syncthreads();
    for (int idx = tid; idx < TILE * TILE; idx += NTHREADS) {
        const int r = idx >> 7, c = idx & 127;
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

// Side streams: at every node the GEMM U = T * Z11 (first half of the inverse update) depends only
// on T and Z11, so it runs beside the SYRK and the whole recursion into A22 -- a chain of small,
// latency-bound kernels that leaves most SMs idle.  One side stream and two events per recursion
// depth (at most one side GEMM is outstanding per depth), per host thread; events are re-recorded
// freely because a stream wait captures the record that precedes it.
constexpr int POTRI_MAX_DEPTH = 12;
struct PotriSide {
    cudaStream_t stream[POTRI_MAX_DEPTH];
    cudaEvent_t forked[POTRI_MAX_DEPTH], joined[POTRI_MAX_DEPTH];
    int device = -1;
    bool ok = false;
    bool init() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        if (ok && dev == device) return true;
        if (ok) return false;            // a thread that switches devices falls back to one stream
        for (int i = 0; i < POTRI_MAX_DEPTH; ++i) {
            if (cudaStreamCreateWithFlags(&stream[i], cudaStreamNonBlocking) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&forked[i], cudaEventDisableTiming) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&joined[i], cudaEventDisableTiming) != cudaSuccess) return false;
        }
        device = dev;
        ok = true;
        return true;
    }
    // worker threads (concurrent restarts) come and go: give the streams back when the thread ends;
    // at process exit the context may already be gone, errors are ignored
    ~PotriSide() {
        if (!ok) return;
        for (int i = 0; i < POTRI_MAX_DEPTH; ++i) {
            cudaStreamDestroy(stream[i]);
            cudaEventDestroy(forked[i]);
            cudaEventDestroy(joined[i]);
        }
    }
};
static thread_local PotriSide g_side;
static thread_local bool g_potri_overlap = true;           // bring-up switch, per host thread
void set_potri_overlap(bool on) { g_potri_overlap = on; }

struct PotriCtx {
    double* A; long lda;
    double* Z; long ldz;
    double* logdiag; int* info;
    double* W;          // scratch (n/2 x n/2), only when keep_L
    bool keep_L;
    cudaStream_t st;
    cudaError_t err;
    PotriSide* side;    // nullptr: everything on st
    int t_refine;       // refinement steps of every panel T (needs keep_L), see potri_lower
    int batch;          // independent problems of the same size, bstride doubles apart in A, Z, logdiag (and info)
    long bstride;
};

static void gemm_checked(PotriCtx& c, bool a_mn, bool b_mn, GemmArgs g, cudaStream_t st) {
    if (c.err != cudaSuccess) return;
    g.batch = c.batch;
    g.bsA = g.bsB = g.bsC = c.bstride;      // every operand of the recursion lives in the problem's own workspace
    c.err = launch_dgemm(a_mn, b_mn, g, st);
}
static void cuda_checked(PotriCtx& c, cudaError_t e) {
    if (c.err == cudaSuccess) c.err = e;
}

static void potri_rec(PotriCtx& c, int off, int n, bool need_inv, int depth) {
    if (c.err != cudaSuccess) return;
    if (n == TILE) {
        potri_leaf_kernel<<<c.batch, NTHREADS, LEAF_SMEM_BYTES, c.st>>>(
            c.A + (long)off * (c.lda + 1), c.lda, c.Z + (long)off * (c.ldz + 1), c.ldz,
            c.logdiag + off, c.info, off, c.keep_L ? 1 : 0, c.bstride);
        c.err = cudaGetLastError();
        return;
    }
    const int n1 = (n / TILE / 2) * TILE, n2 = n - n1;
    double* A21 = c.A + (long)(off + n1) * c.lda + off;
    double* A22 = c.A + (long)(off + n1) * (c.lda + 1);
    double* Z11 = c.Z + (long)off * (c.ldz + 1);
    double* Z21 = c.Z + (long)(off + n1) * c.ldz + off;
    double* Z22 = c.Z + (long)(off + n1) * (c.ldz + 1);

    potri_rec(c, off, n1, true, depth + 1);
    // T = A21 * Z11^T -> Z21 (scratch use of the block that will later hold Z21)
    gemm_checked(c, false, false, GemmArgs{A21, c.lda, Z11, c.ldz, Z21, c.ldz, n2, n1, n1, 1.0, 0.0, 0, KR_LE_N}, c.st);
    // T was formed with the EXPLICIT inverse of L11, so its error grows with cond(L11); the Schur
    // complement A22 - T T^T of an ill-conditioned covariance (entries ~ noise level under entries
    // ~ prior variance) does not survive that.  One step of iterative refinement against the factor
    // itself, T += (A21 - T L11^T) Z11^T, brings T L11^T = A21 down to rounding level -- the
    // backward-stable result a triangular solve would give -- with two more GEMMs on the DMMA pipe.
    for (int it = 0; it < c.t_refine && c.keep_L; ++it) {
        const double* L11 = c.A + (long)off * (c.lda + 1);
        if (c.err == cudaSuccess) {
            copy2d_kernel<<<296, 256, 0, c.st>>>(A21, c.lda, c.W, n1, n2, n1);
            c.err = cudaGetLastError();
        }
        gemm_checked(c, false, false, GemmArgs{Z21, c.ldz, L11, c.lda, c.W, n1, n2, n1, n1, -1.0, 1.0, 0, KR_LE_N}, c.st);
        gemm_checked(c, false, false, GemmArgs{c.W, n1, Z11, c.ldz, Z21, c.ldz, n2, n1, n1, 1.0, 1.0, 0, KR_LE_N}, c.st);
    }
    // U = T * Z11 into A21 (free once T is formed) on the side stream of this depth
    const bool fork = need_inv && !c.keep_L && c.side && depth < POTRI_MAX_DEPTH && n >= 4 * TILE;
    if (fork && c.err == cudaSuccess) {
        cuda_checked(c, cudaEventRecord(c.side->forked[depth], c.st));
        cuda_checked(c, cudaStreamWaitEvent(c.side->stream[depth], c.side->forked[depth], 0));
        gemm_checked(c, false, true, GemmArgs{Z21, c.ldz, Z11, c.ldz, A21, c.lda, n2, n1, n1, 1.0, 0.0, 0, KR_GE_N},
                     c.side->stream[depth]);
        cuda_checked(c, cudaEventRecord(c.side->joined[depth], c.side->stream[depth]));
    }
    // A22 -= T * T^T (lower tiles)
    gemm_checked(c, false, false, GemmArgs{Z21, c.ldz, Z21, c.ldz, A22, c.lda, n2, n2, n1, -1.0, 1.0, 1, KR_FULL}, c.st);
    potri_rec(c, off + n1, n2, need_inv, depth + 1);
    if (c.err != cudaSuccess) return;
    double* U = A21; long ldu = c.lda;
    if (c.keep_L) {
        copy2d_kernel<<<296, 256, 0, c.st>>>(Z21, c.ldz, A21, c.lda, n2, n1);
        c.err = cudaGetLastError();
        U = c.W; ldu = n1;
    }
    if (need_inv) {
        // U = T * Z11 ; Z21 = -Z22 * U
        if (fork) cuda_checked(c, cudaStreamWaitEvent(c.st, c.side->joined[depth], 0));
        else gemm_checked(c, false, true, GemmArgs{Z21, c.ldz, Z11, c.ldz, U, ldu, n2, n1, n1, 1.0, 0.0, 0, KR_GE_N}, c.st);
        gemm_checked(c, false, true, GemmArgs{Z22, c.ldz, U, ldu, Z21, c.ldz, n2, n1, n2, -1.0, 0.0, 0, KR_LE_M}, c.st);
    }
}

cudaError_t potri_lower(double* A, long lda, double* Z, long ldz, int n, double* logdiag, int* info,
                        bool need_inv, bool keep_L, double* W, cudaStream_t st, int t_refine, int batch, long bstride) {
    if (n % TILE || n <= 0 || batch < 1) return cudaErrorInvalidValue;
    if (batch > 1 && (keep_L || t_refine > 0)) return cudaErrorInvalidValue;      // the robust mode shares one scratch W
    static PerDeviceOnce leaf_once;
    const int slot = leaf_once.pending();
    if (slot >= 0) {
        cudaError_t e0 = cudaFuncSetAttribute(potri_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF_SMEM_BYTES);
        if (e0 != cudaSuccess) return e0;
        leaf_once.done[slot] = true;
    }
    cudaError_t e = batch == 1 ? cudaMemsetAsync(info, 0, sizeof(int), st)
                               : cudaMemset2DAsync(info, (size_t)bstride * sizeof(double), 0, sizeof(int), (size_t)batch, st);
    if (e != cudaSuccess) return e;
    PotriSide* side = (g_potri_overlap && !keep_L && need_inv && n >= 4 * TILE && g_side.init()) ? &g_side : nullptr;
    if (t_refine > 0 && (!keep_L || !W)) return cudaErrorInvalidValue;
    PotriCtx c{A, lda, Z, ldz, logdiag, info, W, keep_L, st, cudaSuccess, side, t_refine, batch, batch > 1 ? bstride : 0};
    potri_rec(c, 0, n, need_inv, 0);
    return c.err;
}

// ------------------------------------------------------------------------------------
// Z = L^-1 (row-major, lower) -> tile-major copy for the predictive pass: the 128 x 16 tiles of
// row block li, k-tile kt <= 8 li + 7 are stored contiguously in the order the predict kernel
// consumes them (tile number 8 li (li+1)/2 + kt), each already in the XOR-swizzled
// shared-memory image (kmaj_off), so one 16 KB bulk copy brings a whole operand stage.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_lower_tiles_kernel(const double* __restrict__ Z, long ldz, double* __restrict__ Zt, long bstride) {
    Z += (long)blockIdx.y * bstride;
    Zt += (long)blockIdx.y * bstride;
    int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(I + 1) * (I + 2) / 2 <= t) ++I;
    while ((long)I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const double* src = Z + (long)I * TILE * ldz + (long)J * TILE;
    double* dst = Zt + (8 * ((size_t)I * (I + 1) / 2) + 8 * (size_t)J) * TILE_DOUBLES;
#pragma unroll 4
    for (int q = threadIdx.x; q < TILE * 64; q += 256) {
        const int r = q >> 6, c = q & 63, kt = c >> 3, ch = c & 7;
        const double2 v = *reinterpret_cast<const double2*>(src + (long)r * ldz + 2 * c);
        *reinterpret_cast<double2*>(dst + (size_t)kt * TILE_DOUBLES + r * BK + ((ch ^ ((r & 3) << 1)) << 1)) = v;
    }
}

size_t packed_tiles_doubles(int npad) {
    const size_t nb = (size_t)npad / TILE;
    return 8 * (nb * (nb + 1) / 2) * (size_t)TILE_DOUBLES;
}

cudaError_t pack_lower_tiles(const double* Z, long ldz, int npad, double* Zt, cudaStream_t st, int batch, long bstride) {
    const int nb = npad / TILE;
    pack_lower_tiles_kernel<<<dim3(nb * (nb + 1) / 2, batch), 256, 0, st>>>(Z, ldz, Zt, batch > 1 ? bstride : 0);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// vectors: y -> interleaved, w = Z y, alpha = Z^T w, LML
// ------------------------------------------------------------------------------------
// ncomp = 2: stacked [u; v] -> pair-interleaved; ncomp = 1: scalar observations, zero padded
// perm (optional): internal observation i is the caller's observation perm[i] (order.cu)
__global__ void interleave_kernel(const double* __restrict__ y, int N, int ncomp, double* __restrict__ yi, int npad,
                                  long y_bstride, long bstride, const int* __restrict__ perm, long p_bstride) {
    y += (long)blockIdx.y * y_bstride;
    yi += (long)blockIdx.y * bstride;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npad) return;
    if (ncomp == 2) {
        int i = p >> 1, c = p & 1;
        if (i < N && perm) i = perm[(long)blockIdx.y * p_bstride + i];
        yi[p] = ((p >> 1) < N) ? y[(long)c * N + i] : 0.0;
    } else {
        yi[p] = (p < N) ? y[p] : 0.0;
    }
}

__global__ void deinterleave_kernel(const double* __restrict__ xi, int N, double* __restrict__ x, long bstride,
                                    const int* __restrict__ perm, long p_bstride) {
    xi += (long)blockIdx.y * bstride;
    x += (long)blockIdx.y * 2 * N;
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= 2 * N) return;
    int c = p / N, i = p - c * N;
    const int dst = perm ? perm[(long)blockIdx.y * p_bstride + i] : i;
    x[(long)c * N + dst] = xi[2 * i + c];
}

// w[i] = sum_{k<=i} Z[i][k] y[k]; one warp per row
__global__ void trmv_lower_kernel(const double* __restrict__ Z, long ldz, const double* __restrict__ y,
                                  double* __restrict__ w, int n, long bstride = 0) {
    Z += (long)blockIdx.y * bstride;
    y += (long)blockIdx.y * bstride;
    w += (long)blockIdx.y * bstride;
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

// partial[chunk][j] = sum_{i in chunk, i>=j} Z[i][j] w[i]; 128 columns per CTA, 64-row chunks, four
// independent partial sums per thread (a single dependent chain over 256 rows kept one load in flight:
// 0.6 TB/s in the launch list)
constexpr int TRMVT_ROWS = 64;
__global__ void trmvT_partial_kernel(const double* __restrict__ Z, long ldz, const double* __restrict__ w,
                                     double* __restrict__ partial, int n, long bstride = 0) {
    Z += (long)blockIdx.z * bstride;
    w += (long)blockIdx.z * bstride;
    partial += (long)blockIdx.z * bstride;
    int j = blockIdx.x * 128 + threadIdx.x;
    int r0 = blockIdx.y * TRMVT_ROWS, r1 = min(n, r0 + TRMVT_ROWS);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (r1 > blockIdx.x * 128) {
        int i = max(r0, j);
        const double* z = Z + (long)i * ldz + j;
        for (; i + 3 < r1; i += 4, z += 4 * ldz) {
            s0 = fma(z[0], w[i], s0);
            s1 = fma(z[ldz], w[i + 1], s1);
            s2 = fma(z[2 * ldz], w[i + 2], s2);
            s3 = fma(z[3 * ldz], w[i + 3], s3);
        }
        for (; i < r1; ++i, z += ldz) s0 = fma(z[0], w[i], s0);
    }
    partial[(long)blockIdx.y * n + j] = (s0 + s1) + (s2 + s3);
}

__global__ void colsum_partials_kernel(const double* __restrict__ partial, int nchunks, int n,
                                       double* __restrict__ out, long bstride = 0) {
    partial += (long)blockIdx.y * bstride;
    out += (long)blockIdx.y * bstride;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[(long)c * n + j];
    out[j] = s;
}

// out[0] = -0.5 w'w - sum logdiag - (n/2) log(2 pi)      (n scalar observations)
__global__ void lml_kernel(const double* __restrict__ w, const double* __restrict__ logdiag, int npad,
                           double half_n, double* __restrict__ out, long bstride) {
    w += (long)blockIdx.x * bstride;
    logdiag += (long)blockIdx.x * bstride;
    out += (long)blockIdx.x * bstride;
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
        out[0] = -0.5 * sa - sb - half_n * 1.8378770664093454836;   // log(2 pi)
    }
}

// r = y - (t1 + t2 - diag(K) alpha): residual of the symmetric system from its two triangular products
__global__ void sym_residual_kernel(const double* __restrict__ y, const double* __restrict__ t1, const double* t2,
                                    const double* __restrict__ K, long ldk, const double* __restrict__ alpha,
                                    double* r, int n) {      // r may alias t2
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = y[i] - (t1[i] + t2[i] - K[(long)i * (ldk + 1)] * alpha[i]);
}
__global__ void axpy1_kernel(const double* __restrict__ d, double* __restrict__ x, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += d[i];
}

// Iterative refinement of alpha against the matrix itself (robust mode): `steps` times
//   alpha += Z^T Z (y - K alpha),   K = lower tiles of the padded covariance (identity padding).
// With refined factor panels Z^T Z is a good preconditioner and two steps bring K alpha = y down to
// the rounding level of the residual -- what a backward-stable triangular solve delivers -- while
// the explicit-inverse product alone is off by ~cond(K) eps.  t1, t2: scratch vectors of npad.
cudaError_t refine_alpha(const double* K, long ldk, const double* Z, long ldz, int npad, const double* y_int,
                         double* t1, double* t2, double* alpha_int, double* partial, int steps, cudaStream_t st) {
    const int nchunks = (npad + TRMVT_ROWS - 1) / TRMVT_ROWS;
    const dim3 gT(npad / 128, nchunks);
    const int gv = (npad + 255) / 256, gr = (npad + 7) / 8;
    for (int it = 0; it < steps; ++it) {
        trmv_lower_kernel<<<gr, 256, 0, st>>>(K, ldk, alpha_int, t1, npad);                 // tril(K) alpha
        trmvT_partial_kernel<<<gT, 128, 0, st>>>(K, ldk, alpha_int, partial, npad);         // tril(K)^T alpha
        colsum_partials_kernel<<<gv, 256, 0, st>>>(partial, nchunks, npad, t2);
        sym_residual_kernel<<<gv, 256, 0, st>>>(y_int, t1, t2, K, ldk, alpha_int, t2, npad);   // t2 = r
        trmv_lower_kernel<<<gr, 256, 0, st>>>(Z, ldz, t2, t1, npad);                        // Z r
        trmvT_partial_kernel<<<gT, 128, 0, st>>>(Z, ldz, t1, partial, npad);                // Z^T (Z r)
        colsum_partials_kernel<<<gv, 256, 0, st>>>(partial, nchunks, npad, t2);
        axpy1_kernel<<<gv, 256, 0, st>>>(t2, alpha_int, npad);
    }
    return cudaGetLastError();
}

cudaError_t solve_alpha_lml(const double* Z, long ldz, int npad, int N, int ncomp, const double* y_block,
                            double* y_int, double* w, double* alpha_int, double* partial,
                            const double* logdiag, double* lml_out, cudaStream_t st, int batch, long y_bstride,
                            long bstride, const int* perm, long p_bstride) {
    if (batch == 1) bstride = y_bstride = p_bstride = 0;
    interleave_kernel<<<dim3((npad + 255) / 256, batch), 256, 0, st>>>(y_block, N, ncomp, y_int, npad, y_bstride, bstride, perm, p_bstride);
    trmv_lower_kernel<<<dim3((npad + 7) / 8, batch), 256, 0, st>>>(Z, ldz, y_int, w, npad, bstride);
    int nchunks = (npad + TRMVT_ROWS - 1) / TRMVT_ROWS;
    trmvT_partial_kernel<<<dim3(npad / 128, nchunks, batch), 128, 0, st>>>(Z, ldz, w, partial, npad, bstride);
    colsum_partials_kernel<<<dim3((npad + 255) / 256, batch), 256, 0, st>>>(partial, nchunks, npad, alpha_int, bstride);
    lml_kernel<<<batch, 1024, 0, st>>>(w, logdiag, npad, 0.5 * ncomp * (double)N, lml_out, bstride);
    return cudaGetLastError();
}

cudaError_t deinterleave(const double* xi, int N, double* x, cudaStream_t st, int batch, long bstride, const int* perm, long p_bstride) {
    deinterleave_kernel<<<dim3((2 * N + 255) / 256, batch), 256, 0, st>>>(xi, N, x, batch > 1 ? bstride : 0, perm, batch > 1 ? p_bstride : 0);
    return cudaGetLastError();
}

}  // namespace gp2d
