// Spatial order of the observations inside a fit (Helmholtz families with the int8 predictive state).
//
// The int8 predictive kernel (predict_i8.cu) neither copies nor multiplies digit slices that are identically zero:
// tiles of K* between a group of 16 consecutive observations and 40 consecutive grid points that lie far apart, tiles
// of L^-1 far from the diagonal.  Both only exist when consecutive observations are close in space.  The reference
// hands over drifters in whatever order the tracks file has them (GP_laser.py:62-110, krig.py:300-369), for which
// tools/sparsity_emulate.py finds no zero slice at all at configs[1] (0.998 of the products) against 0.61 after a
// Morton sort.  A GP does not care about the order of its observations, so the fit sorts them along a Z-order curve:
// X and y are permuted on the way into the fit state, alpha on the way out; the caller never sees the internal order.
#include "common.cuh"
#include "linalg.h"

namespace gp2d {

constexpr int ORDER_THREADS = 1024;

__device__ __forceinline__ unsigned spread8(unsigned v) {      // abcdefgh -> 0a0b0c0d0e0f0g0h
    v = (v | (v << 4)) & 0x0F0Fu;
    v = (v | (v << 2)) & 0x3333u;
    v = (v | (v << 1)) & 0x5555u;
    return v;
}

// one CTA per problem: 16-bit Morton key of (a, b) on a 256 x 256 raster of the bounding box, index in the low 16
// bits (ties keep the caller's order: the result is deterministic), bitonic sort of P = 2^k >= N words in shared memory
__global__ void __launch_bounds__(ORDER_THREADS) spatial_order_kernel(const double* __restrict__ X, int ldx, int xo, int N, int P,
                                                                      int* __restrict__ perm, long x_bstride, long p_bstride) {
    extern __shared__ unsigned keys[];
    __shared__ double red[4][32];
    X += (long)blockIdx.x * x_bstride;
    perm += (long)blockIdx.x * p_bstride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double lo0 = 1e300, hi0 = -1e300, lo1 = 1e300, hi1 = -1e300;
    for (int i = tid; i < N; i += ORDER_THREADS) {
        const double a = X[(long)i * ldx + xo], b = X[(long)i * ldx + xo + 1];
        lo0 = fmin(lo0, a); hi0 = fmax(hi0, a); lo1 = fmin(lo1, b); hi1 = fmax(hi1, b);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo0 = fmin(lo0, __shfl_xor_sync(0xffffffffu, lo0, o)); hi0 = fmax(hi0, __shfl_xor_sync(0xffffffffu, hi0, o));
        lo1 = fmin(lo1, __shfl_xor_sync(0xffffffffu, lo1, o)); hi1 = fmax(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
    }
    if (lane == 0) { red[0][warp] = lo0; red[1][warp] = hi0; red[2][warp] = lo1; red[3][warp] = hi1; }
    __syncthreads();
    lo0 = red[0][0]; hi0 = red[1][0]; lo1 = red[2][0]; hi1 = red[3][0];
    for (int w = 1; w < ORDER_THREADS / 32; ++w) {
        lo0 = fmin(lo0, red[0][w]); hi0 = fmax(hi0, red[1][w]); lo1 = fmin(lo1, red[2][w]); hi1 = fmax(hi1, red[3][w]);
    }
    // one raster for both axes (square cells): the curve then follows distance, not the aspect ratio of the box
    const double ext = fmax(hi0 - lo0, hi1 - lo1);
    const double s = (ext > 0.0 && isfinite(ext)) ? 255.999 / ext : 0.0;
    for (int i = tid; i < P; i += ORDER_THREADS) {
        unsigned k = 0xFFFF0000u | (unsigned)(i & 0xFFFF);
        if (i < N) {
            const double a = X[(long)i * ldx + xo], b = X[(long)i * ldx + xo + 1];
            const unsigned qa = (unsigned)fmin(fmax((a - lo0) * s, 0.0), 255.0), qb = (unsigned)fmin(fmax((b - lo1) * s, 0.0), 255.0);
            k = ((spread8(qa) | (spread8(qb) << 1)) << 16) | (unsigned)i;
        }
        keys[i] = k;                                   // padding sorts last (key 0xFFFF, and its index is >= N)
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += ORDER_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const unsigned x = keys[i], y = keys[l];
                const bool up = (i & k) == 0;
                if ((x > y) == up) { keys[i] = y; keys[l] = x; }
            }
            __syncthreads();
        }
    }
    // a real point may share the key 0xFFFF with the padding, but its index is below N and the padding's is not: the first
    // N words are the real points
    for (int i = tid; i < N; i += ORDER_THREADS) perm[i] = (int)(keys[i] & 0xFFFFu);
}

__global__ void gather_points_kernel(const double* __restrict__ X, int ldx, int N, const int* __restrict__ perm, double* __restrict__ out,
                                     long x_bstride, long p_bstride, long o_bstride) {
    X += (long)blockIdx.y * x_bstride;
    perm += (long)blockIdx.y * p_bstride;
    out += (long)blockIdx.y * o_bstride;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int s = perm[i];
    for (int c = 0; c < ldx; ++c) out[(long)i * ldx + c] = X[(long)s * ldx + c];
}

int spatial_order_max_points() { return 32768; }

cudaError_t spatial_order(const double* X, int ldx, int xo, int N, int* perm, cudaStream_t st, int batch, long x_bstride, long p_bstride) {
    if (N <= 0 || N > spatial_order_max_points()) return cudaErrorInvalidValue;
    int P = 2;
    while (P < N) P <<= 1;
    static PerDeviceOnce once;
    const int slot = once.pending();
    if (slot >= 0) {
        cudaError_t e = cudaFuncSetAttribute(spatial_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4);
        if (e != cudaSuccess) return e;
        once.done[slot] = true;
    }
    spatial_order_kernel<<<batch, ORDER_THREADS, (size_t)P * sizeof(unsigned), st>>>(X, ldx, xo, N, P, perm, x_bstride, p_bstride);
    return cudaGetLastError();
}

cudaError_t gather_points(const double* X, int ldx, int N, const int* perm, double* out, cudaStream_t st, int batch, long x_bstride,
                          long p_bstride, long o_bstride) {
    gather_points_kernel<<<dim3((N + 255) / 256, batch), 256, 0, st>>>(X, ldx, N, perm, out, x_bstride, p_bstride, o_bstride);
    return cudaGetLastError();
}

}  // namespace gp2d
