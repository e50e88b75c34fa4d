// Fused predictive pass: for a tile of 64 grid points (128 stacked columns)
//     K*   generated on the fly into the shared-memory B operand (never touches HBM)
//     mean = K* alpha                      (accumulated while the last row block is generated)
//     V    = Z K*^T   (Z = L^-1, DMMA)     -> var = k** - colsumsq(V), clamped at 0
// Replaces GPy model.predict (krig.py:543-544; GP_plots.py:768), GP_scripts.getMean +
// the diagonal of GP_laser.py:129-131, and sklearn predict(return_std=True)
// (krig.py:194; algebra of _gpr.py:446-491).
//
// One CTA owns a column tile and walks every row block of Z, so the column sums of V^2
// stay in registers: no atomics, no partial buffers, bit-reproducible.  Every CTA does the
// same amount of work (the whole lower triangle of Z), so the grid is load-balanced by
// construction.  A tiles: 4-stage cp.async ring; B tiles: generated one iteration ahead
// into a 2-deep ring by all 256 threads (1-2 exps per point pair give 4 matrix entries
// thanks to the pair-interleaved internal ordering).
#include "common.cuh"
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
};

constexpr int P_ASTAGES = 4;
constexpr int P_BSTAGES = 2;
constexpr int PRED_SMEM_BYTES = (P_ASTAGES + P_BSTAGES) * TILE_DOUBLES * (int)sizeof(double);   // 96 KB

__global__ void __launch_bounds__(NTHREADS, 1) predict_kernel(PredictArgs p) {
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + P_ASTAGES * TILE_DOUBLES;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int gp0 = blockIdx.x * 64;
    const int gjl = tid & 63, os = tid >> 6;
    const int gj = gp0 + gjl;
    const bool gvalid = gj < p.M;
    const double gx = gvalid ? p.Xs[2 * (long)gj] : 0.0;
    const double gy = gvalid ? p.Xs[2 * (long)gj + 1] : 0.0;
    const int nb = p.npad / TILE;
    const int total = 8 * (nb * (nb + 1) / 2);

    double acc[8][4][2];
    double css[4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) css[j][0] = css[j][1] = 0.0;
    double mu0 = 0.0, mu1 = 0.0;

    // cursors over (row block, k tile): compute, A-load (3 ahead), B-generate (1 ahead)
    int ci = 0, ckt = 0, li = 0, lkt = 0, gi = 0, gkt = 0;

    auto load_A = [&](int stage) {
        if (li < nb) {
            load_tile_async<false>(As + stage * TILE_DOUBLES,
                                   p.Z + (long)li * TILE * p.ldz + lkt * BK, p.ldz, tid);
            if (++lkt == 8 * (li + 1)) { lkt = 0; ++li; }
        }
        cp_async_commit();
    };
    auto gen_B = [&](int stage) {
        if (gi >= nb) return;
        double* B = Bs + stage * TILE_DOUBLES;
        const bool last = (gi == nb - 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ol = os + 4 * h;
            const int o = gkt * 8 + ol;
            double k11 = 0.0, k12 = 0.0, k22 = 0.0;
            if (gvalid && o < p.N) {
                helm_block(p.hp, p.X[2 * (long)o] - gx, p.X[2 * (long)o + 1] - gy, k11, k12, k22);
                if (last) {
                    const double a0 = p.alpha[2 * o], a1 = p.alpha[2 * o + 1];
                    mu0 = fma(k11, a0, fma(k12, a1, mu0));
                    mu1 = fma(k12, a0, fma(k22, a1, mu1));
                }
            }
            const int k = 2 * ol;
            *reinterpret_cast<double2*>(B + mnmaj_off(k, 2 * gjl)) = make_double2(k11, k12);
            *reinterpret_cast<double2*>(B + mnmaj_off(k + 1, 2 * gjl)) = make_double2(k12, k22);
        }
        if (++gkt == 8 * (gi + 1)) { gkt = 0; ++gi; }
    };

#pragma unroll
    for (int s = 0; s < P_ASTAGES - 1; ++s) load_A(s);
    gen_B(0);

    for (int it = 0; it < total; ++it) {
        cp_async_wait<P_ASTAGES - 2>();
        __syncthreads();
        load_A((it + P_ASTAGES - 1) % P_ASTAGES);
        gen_B((it + 1) & 1);
        mma_stage<false, true>(As + (it % P_ASTAGES) * TILE_DOUBLES, Bs + (it & 1) * TILE_DOUBLES,
                               wm, wn, lane, acc);
        if (++ckt == 8 * (ci + 1)) {
            ckt = 0; ++ci;
#pragma unroll
            for (int mb = 0; mb < 8; ++mb)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    css[j][0] = fma(acc[mb][j][0], acc[mb][j][0], css[j][0]);
                    css[j][1] = fma(acc[mb][j][1], acc[mb][j][1], css[j][1]);
                    acc[mb][j][0] = acc[mb][j][1] = 0.0;
                }
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // reduce: css over the 8 row groups of the warp (lane>>2), then over the two M-warps;
    // mu over the four obs slots.  smem is reused.
    double* sh_css = smem;               // [2][128]
    double* sh_mu = smem + 256;          // [4][64][2]
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
            double ss = sh_css[tid] + sh_css[TILE + tid];
            double v = p.kss - ss;
            v = v < 0.0 ? 0.0 : v;
            p.var[(long)c * p.out_stride + j] = v + p.var_add;
            double m = (sh_mu[(0 * 64 + pj) * 2 + c] + sh_mu[(1 * 64 + pj) * 2 + c]) +
                       (sh_mu[(2 * 64 + pj) * 2 + c] + sh_mu[(3 * 64 + pj) * 2 + c]);
            p.mean[(long)c * p.out_stride + j] = m;
        }
    }
}

cudaError_t predict_fused(const double* Z, long ldz, int npad, const double* alpha_int,
                          const double* X, int N, const HelmParams& hp, const double* Xs, int M,
                          long out_stride, double var_add, double* mean, double* var, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    static bool init = false;
    if (!init) {
        cudaError_t e = cudaFuncSetAttribute(predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        init = true;
    }
    PredictArgs a;
    a.Z = Z; a.ldz = ldz; a.npad = npad; a.alpha = alpha_int; a.X = X; a.N = N; a.hp = hp;
    a.Xs = Xs; a.M = M; a.out_stride = out_stride;
    a.kss = hp.w_df + hp.w_cf;           // ratio/l_df^2 + (1-ratio)/l_cf^2   (myKernel.py:55-57)
    a.var_add = var_add; a.mean = mean; a.var = var;
    predict_kernel<<<(M + 63) / 64, NTHREADS, PRED_SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

}  // namespace gp2d
