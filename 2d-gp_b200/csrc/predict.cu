// Fused predictive pass.  A persistent CTA owns a tile of 64 grid points (128 stacked
// columns) at a time and runs two phases on it:
//   1. generate the K* panel [npad x 128] ONCE (1-2 exps per point pair give the four
//      entries of a 2x2 block thanks to the pair-interleaved internal ordering), store it
//      into the CTA's private scratch panel (written once, re-read from L2 by every row
//      block) and accumulate mean = K* alpha on the way;
//   2. V = Z K*^T on the DMMA pipe (Z = L^-1), row block after row block, with both
//      operands streamed by cp.async; the column sums of V^2 stay in registers, so
//      var = k** - colsumsq(V) needs no atomics, no partial buffers and is bit-reproducible.
// Replaces GPy model.predict (krig.py:543-544; GP_plots.py:768), GP_scripts.getMean +
// the diagonal of GP_laser.py:129-131, and sklearn predict(return_std=True)
// (krig.py:194; algebra of _gpr.py:446-491).
//
// Why a scratch panel: DMMA and DFMA share one FP64 pipe on sm_100a (measured: 37 TF/s
// each, 36 TF/s mixed).  Regenerating K* tiles inside the k-loop of every row block cost
// 2 x (n/256) exps per pair and 20% of the kernel; generating once costs < 1% and the
// re-reads (n/256 x 1 KB per panel row, from L2) ride on otherwise idle bandwidth.
// The full K* (n x 2M) still never exists: only one [npad x 128] panel per resident CTA.
#include "common.cuh"
#include "dgemm.cuh"
#include "linalg.h"

namespace gp2d {

struct PredictArgs {
    const double* Z; long ldz; int npad;
    const double* alpha;          // interleaved, zero padded to npad
    const double* X; int N;
    HelmParams hp;
    const double* Xs; int M;
    long out_stride;              // component stride of mean/var
    double kss, var_add;
    double* mean; double* var;
    double* scratch;              // gridDim.x panels of npad x 128 doubles
    int ntiles;
};

constexpr int P_STAGES = 4;
constexpr int PRED_SMEM_BYTES = P_STAGES * 2 * TILE_DOUBLES * (int)sizeof(double);   // 128 KB

template <int NT>
__global__ void __launch_bounds__(NT, 1) predict_kernel(PredictArgs p) {
    constexpr int MB = mblocks(NT), WM = warps_m(NT), OS = NT / 64;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + P_STAGES * TILE_DOUBLES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int gjl = tid & 63, os = tid >> 6;
    const int nb = p.npad / TILE;
    const int total = 8 * (nb * (nb + 1) / 2);
    double* panel = p.scratch + (size_t)blockIdx.x * p.npad * TILE;

    for (int ct = blockIdx.x; ct < p.ntiles; ct += gridDim.x) {
        const int gp0 = ct * 64;
        // ---------------- phase 1: K* panel + mean -----------------------------------------
        double mu0 = 0.0, mu1 = 0.0;
        {
            const int gj = gp0 + gjl;
            const bool gvalid = gj < p.M;
            const double gx = gvalid ? p.Xs[2 * (long)gj] : 0.0;
            const double gy = gvalid ? p.Xs[2 * (long)gj + 1] : 0.0;
            const int nobs_pad = p.npad >> 1;
#pragma unroll 2
            for (int o = os; o < nobs_pad; o += OS) {
                double k11 = 0.0, k12 = 0.0, k22 = 0.0;
                if (gvalid && o < p.N) {
                    helm_block(p.hp, p.X[2 * (long)o] - gx, p.X[2 * (long)o + 1] - gy, k11, k12, k22);
                    const double a0 = p.alpha[2 * o], a1 = p.alpha[2 * o + 1];
                    mu0 = fma(k11, a0, fma(k12, a1, mu0));
                    mu1 = fma(k12, a0, fma(k22, a1, mu1));
                }
                double* r0 = panel + (size_t)(2 * o) * TILE + 2 * gjl;
                *reinterpret_cast<double2*>(r0) = make_double2(k11, k12);
                *reinterpret_cast<double2*>(r0 + TILE) = make_double2(k12, k22);
            }
        }
        __syncthreads();      // panel (global) visible to every thread of the CTA

        // ---------------- phase 2: column sums of (Z K*^T)^2 -------------------------------
        double acc[MB][4][2];
        double css[4][2];
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) css[j][0] = css[j][1] = 0.0;

        int ci = 0, ckt = 0, li = 0, lkt = 0;      // compute / load cursors (row block, k tile)
        auto load_AB = [&](int stage) {
            if (li < nb) {
                load_tile_async<false, NT>(As + stage * TILE_DOUBLES,
                                           p.Z + (long)li * TILE * p.ldz + lkt * BK, p.ldz, tid);
                load_tile_async<true, NT>(Bs + stage * TILE_DOUBLES, panel + (size_t)lkt * BK * TILE, TILE, tid);
                if (++lkt == 8 * (li + 1)) { lkt = 0; ++li; }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int s = 0; s < P_STAGES - 1; ++s) load_AB(s);

        for (int it = 0; it < total; ++it) {
            cp_async_wait<P_STAGES - 2>();
            __syncthreads();
            load_AB((it + P_STAGES - 1) % P_STAGES);
            mma_stage<false, true, MB>(As + (it % P_STAGES) * TILE_DOUBLES, Bs + (it % P_STAGES) * TILE_DOUBLES,
                                       wm, wn, lane, acc);
            if (++ckt == 8 * (ci + 1)) {
                ckt = 0; ++ci;
#pragma unroll
                for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        css[j][0] = fma(acc[mb][j][0], acc[mb][j][0], css[j][0]);
                        css[j][1] = fma(acc[mb][j][1], acc[mb][j][1], css[j][1]);
                        acc[mb][j][0] = acc[mb][j][1] = 0.0;
                    }
            }
        }
        cp_async_wait<0>();
        __syncthreads();      // all panel / smem reads done: both may be reused below

        // reduce css over the 8 row groups of a warp (lane>>2) and the M-warps; mu over the
        // obs slots.  Fixed order: results do not depend on the grid partition.
        double* sh_css = smem;               // [WM][128]
        double* sh_mu = smem + WM * TILE;    // [OS][64][2]
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double v = css[j][e];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if ((lane >> 2) == 0) sh_css[wm * TILE + wn * 32 + j * 8 + 2 * (lane & 3) + e] = v;
            }
        sh_mu[(os * 64 + gjl) * 2 + 0] = mu0;
        sh_mu[(os * 64 + gjl) * 2 + 1] = mu1;
        __syncthreads();
        if (tid < TILE) {
            const int pj = tid >> 1, c = tid & 1;
            const int j = gp0 + pj;
            if (j < p.M) {
                double ss = 0.0;
#pragma unroll
                for (int w = 0; w < WM; ++w) ss += sh_css[w * TILE + tid];
                double v = p.kss - ss;
                v = v < 0.0 ? 0.0 : v;
                p.var[(long)c * p.out_stride + j] = v + p.var_add;
                double m = 0.0;
#pragma unroll
                for (int o = 0; o < OS; ++o) m += sh_mu[(o * 64 + pj) * 2 + c];
                p.mean[(long)c * p.out_stride + j] = m;
            }
        }
        __syncthreads();      // sh_* consumed before the next tile's cp.async overwrites smem
    }
}

size_t predict_panel_bytes(int npad) { return (size_t)npad * TILE * sizeof(double); }

int predict_max_ctas() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

cudaError_t predict_fused(const double* Z, long ldz, int npad, const double* alpha_int,
                          const double* X, int N, const HelmParams& hp, const double* Xs, int M,
                          long out_stride, double var_add, double* mean, double* var,
                          double* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    static bool init = false;
    if (!init) {
        cudaError_t e = cudaFuncSetAttribute(predict_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(predict_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        init = true;
    }
    PredictArgs a;
    a.Z = Z; a.ldz = ldz; a.npad = npad; a.alpha = alpha_int; a.X = X; a.N = N; a.hp = hp;
    a.Xs = Xs; a.M = M; a.out_stride = out_stride;
    a.kss = hp.w_df + hp.w_cf;           // ratio/l_df^2 + (1-ratio)/l_cf^2   (myKernel.py:55-57)
    a.var_add = var_add; a.mean = mean; a.var = var;
    a.scratch = scratch;
    a.ntiles = (M + 63) / 64;
    long panels = (long)(scratch_bytes / predict_panel_bytes(npad));
    long grid = a.ntiles;
    if (grid > predict_max_ctas()) grid = predict_max_ctas();
    if (grid > panels) grid = panels;
    if (grid <= 0) return cudaErrorInvalidValue;
    // balance the tail: every CTA gets ceil(ntiles/grid) or one fewer tiles
    long per = (a.ntiles + grid - 1) / grid;
    grid = (a.ntiles + per - 1) / per;
    if (get_cta_threads() == 512) predict_kernel<512><<<(unsigned)grid, 512, PRED_SMEM_BYTES, st>>>(a);
    else predict_kernel<256><<<(unsigned)grid, 256, PRED_SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

}  // namespace gp2d
