// FP64 GEMM on the DMMA tensor pipe for the blocked Cholesky / triangular inverse / K^-1.
//
//   C[M,N] = alpha * opA(A) * opB(B) + beta * C        (all row-major storage, ld in doubles)
//
// A operand: element (m,k) at A[m*lda + k] (K-major, "N") or A[k*lda + m] (MN-major, "T").
// B operand: element (n,k) at B[n*ldb + k] (K-major, i.e. C = A * B^T) or
//            B[k*ldb + n] (MN-major, i.e. C = A * B).
// M, N are multiples of 128 and K of 16; triangular structure is exploited by clipping the
// k-range per output tile (KR_* flags) and by skipping tiles above the diagonal
// (lower_out).  Two kernels: dgemm_ws_kernel, 128 x 128 x 16 tiles, warp-specialised
// (pipeline.cuh), for everything that fills the GPU; dgemm_kernel<.., 128, 64>, 64 x 64 x 16 tiles
// with 4 warps of 32x32, 4-stage cp.async pipeline and up to 3 CTAs per SM, for the small GEMMs of
// the Cholesky recursion.  XOR-swizzled shared memory, DMMA.8x8x4 register tiles.
#pragma once
#include "common.cuh"
#include "pipeline.cuh"

namespace gp2d {

enum : int {
    KR_FULL = 0,
    KR_LE_M = 1,   // k <  (tm+1)*TS : A operand lower-triangular in (m,k)
    KR_LE_N = 2,   // k <  (tn+1)*TS : B operand lower-triangular in (n,k)  (Z^T as K-major B)
    KR_GE_N = 4,   // k >= tn*TS     : B operand (k,n) lower-triangular, zero for k < n
    KR_GE_M = 8,   // k >= tm*TS     : A operand (k,m) lower-triangular, zero for k < m
};

struct GemmArgs {
    const double* A; long lda;
    const double* B; long ldb;
    double* C; long ldc;
    int M, N, K;
    double alpha, beta;
    int lower_out;
    int krule;
    // batch of independent products of the same shape (gridDim.y): problem b works on A + b bsA,
    // B + b bsB, C + b bsC (strides in doubles)
    int batch = 1;
    long bsA = 0, bsB = 0, bsC = 0;
};

// Tile (tm, tn), tn <= tm, number t of the banded walk over the lower triangle of a tiles_m x tiles_m
// grid.  Bands of GRL tile rows in row order (band starting at row r0 begins at t = r0 (r0+1)/2, as in
// plain row-major numbering); inside a band the columns 0..r0, which every row of the band owns, are
// walked column by column, then the small triangle on the diagonal row by row: the CTAs in flight share
// operand panels in L2, like the grouped raster of the rectangular case.
template <int GRL>
__device__ __forceinline__ void lower_tile(int t, int tiles_m, int& tm, int& tn) {
    int row = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(row + 1) * (row + 2) / 2 <= t) ++row;
    while ((long)row * (row + 1) / 2 > t) --row;
    const int r0 = (row / GRL) * GRL;
    const int h = min(GRL, tiles_m - r0);
    int q = t - (int)((long)r0 * (r0 + 1) / 2);
    const int rect = (r0 + 1) * h;
    if (q < rect) {
        tn = q / h;
        tm = r0 + q % h;
        return;
    }
    q -= rect;                                   // rows r0+1 .. r0+h-1, columns r0+1 .. row
    int i = (int)((sqrt(8.0 * (double)q + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= q) ++i;
    while (i * (i + 1) / 2 > q) --i;
    tm = r0 + 1 + i;
    tn = r0 + 1 + (q - i * (i + 1) / 2);
}

constexpr int GEMM_STAGES = 4;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * 2 * TILE_DOUBLES * (int)sizeof(double);   // 128 KB (TS = 128)
constexpr int gemm_smem_bytes(int ts) { return GEMM_STAGES * 2 * ts * BK * (int)sizeof(double); }

template <bool A_MN, bool B_MN, int NT, int TS>
__global__ void __launch_bounds__(NT, (TS == 64 ? 3 : 1)) dgemm_kernel(GemmArgs p) {
    p.A += (long)blockIdx.y * p.bsA;
    p.B += (long)blockIdx.y * p.bsB;
    p.C += (long)blockIdx.y * p.bsC;
    using Cfg = TileCfg<TS, NT>;
    constexpr int MB = Cfg::MB, WM = Cfg::WM, SD = Cfg::STAGE_DOUBLES;
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + GEMM_STAGES * SD;

    int tm, tn;
    if (p.lower_out) {
        // lower triangle of the tile grid in bands of tile rows, heaviest rows last (lower_tile)
        lower_tile<TS == 64 ? 16 : 8>((int)blockIdx.x, p.M / TS, tm, tn);
    } else {
        // grouped raster: bands of GR tile rows, walked column by column, so that the CTAs in flight
        // form a near-square patch and share operand panels in L2 (plain row-major order streamed
        // 61 GB from DRAM for an 8192^3 product whose operands are 1 GB)
        constexpr int GR = TS == 64 ? 16 : 8;
        const int tiles_m = p.M / TS, tiles_n = p.N / TS;
        const int per = GR * tiles_n, band = blockIdx.x / per, first = band * GR;
        const int h = min(GR, tiles_m - first), r = blockIdx.x - band * per;
        tm = first + r % h;
        tn = r / h;
    }
    int k0 = 0, k1 = p.K;
    if (p.krule & KR_LE_M) k1 = min(k1, (tm + 1) * TS);
    if (p.krule & KR_LE_N) k1 = min(k1, (tn + 1) * TS);
    if (p.krule & KR_GE_N) k0 = max(k0, tn * TS);
    if (p.krule & KR_GE_M) k0 = max(k0, tm * TS);
    const int nk = (k1 - k0) / BK;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;

    const double* Ag = A_MN ? p.A + (long)k0 * p.lda + (long)tm * TS
                            : p.A + (long)tm * TS * p.lda + k0;
    const double* Bg = B_MN ? p.B + (long)k0 * p.ldb + (long)tn * TS
                            : p.B + (long)tn * TS * p.ldb + k0;
    const long a_step = A_MN ? (long)BK * p.lda : BK;
    const long b_step = B_MN ? (long)BK * p.ldb : BK;

    double acc[MB][4][2];
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; ++s) {
        if (s < nk) {
            load_tile_async<A_MN, NT, TS>(As + s * SD, Ag + s * a_step, p.lda, tid);
            load_tile_async<B_MN, NT, TS>(Bs + s * SD, Bg + s * b_step, p.ldb, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<GEMM_STAGES - 2>();
        __syncthreads();
        int nxt = kt + GEMM_STAGES - 1;
        if (nxt < nk) {
            int s = nxt % GEMM_STAGES;
            load_tile_async<A_MN, NT, TS>(As + s * SD, Ag + nxt * a_step, p.lda, tid);
            load_tile_async<B_MN, NT, TS>(Bs + s * SD, Bg + nxt * b_step, p.ldb, tid);
        }
        cp_async_commit();
        int s = kt % GEMM_STAGES;
        mma_stage<A_MN, B_MN, MB, TS>(As + s * SD, Bs + s * SD, wm, wn, lane, acc);
    }
    cp_async_wait<0>();

    // epilogue: lane holds C[g][2*tig + {0,1}] of each 8x8 block -> 16-byte stores
    const int g = lane >> 2, tig = lane & 3;
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
        long row = (long)tm * TS + wm * (MB * 8) + mb * 8 + g;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            long col = (long)tn * TS + wn * 32 + nb * 8 + 2 * tig;
            double2* dst = reinterpret_cast<double2*>(p.C + row * p.ldc + col);
            double2 v;
            v.x = alpha * acc[mb][nb][0];
            v.y = alpha * acc[mb][nb][1];
            if (beta != 0.0) {
                double2 old = *dst;
                v.x = fma(beta, old.x, v.x);
                v.y = fma(beta, old.y, v.y);
            }
            *dst = v;
        }
    }
}


// ---- warp-specialised 128 x 128 tile kernel (pipeline.cuh) ---------------------------------
// Same contract as dgemm_kernel<A_MN, B_MN, 256, 128>; 8 consumer warps + 1 producer warp,
// 6-stage mbarrier ring, no CTA-wide barrier in the k-loop.
constexpr int GEMM_WS_SMEM_BYTES = WS_RING_BYTES + WS_BAR_BYTES;

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(WS_THREADS, 1) dgemm_ws_kernel(GemmArgs p) {
    p.A += (long)blockIdx.y * p.bsA;
    p.B += (long)blockIdx.y * p.bsB;
    p.C += (long)blockIdx.y * p.bsC;
    constexpr int TS = TILE;
    extern __shared__ __align__(16) double smem[];
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + WS_STAGES * WS_STAGE_DOUBLES);
    WsBarriers wb{bars, bars + WS_STAGES};

    int tm, tn;
    if (p.lower_out) {
        // lower triangle of the tile grid in bands of tile rows, heaviest rows last (lower_tile)
        lower_tile<TS == 64 ? 16 : 8>((int)blockIdx.x, p.M / TS, tm, tn);
    } else {
        // grouped raster: bands of GR tile rows, walked column by column, so that the CTAs in flight
        // form a near-square patch and share operand panels in L2 (plain row-major order streamed
        // 61 GB from DRAM for an 8192^3 product whose operands are 1 GB)
        constexpr int GR = TS == 64 ? 16 : 8;
        const int tiles_m = p.M / TS, tiles_n = p.N / TS;
        const int per = GR * tiles_n, band = blockIdx.x / per, first = band * GR;
        const int h = min(GR, tiles_m - first), r = blockIdx.x - band * per;
        tm = first + r % h;
        tn = r / h;
    }
    int k0 = 0, k1 = p.K;
    if (p.krule & KR_LE_M) k1 = min(k1, (tm + 1) * TS);
    if (p.krule & KR_LE_N) k1 = min(k1, (tn + 1) * TS);
    if (p.krule & KR_GE_N) k0 = max(k0, tn * TS);
    if (p.krule & KR_GE_M) k0 = max(k0, tm * TS);
    const int nk = (k1 - k0) / BK;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    wb.init(tid);
    __syncthreads();

    if (warp == WS_CONSUMERS / 32) {
        // ---------------- producer warp ----------------
        const double* Ag = A_MN ? p.A + (long)k0 * p.lda + (long)tm * TS : p.A + (long)tm * TS * p.lda + k0;
        const double* Bg = B_MN ? p.B + (long)k0 * p.ldb + (long)tn * TS : p.B + (long)tn * TS * p.ldb + k0;
        const long a_step = A_MN ? (long)BK * p.lda : BK;
        const long b_step = B_MN ? (long)BK * p.ldb : BK;
        int s = 0;
        unsigned ph = 0;
        for (int kt = 0; kt < nk; ++kt) {
            if (kt >= WS_STAGES) mbar_wait(wb.empty + s, ph ^ 1u);
            double* st = smem + s * WS_STAGE_DOUBLES;
            ws_load_tile<A_MN>(st, Ag, p.lda, lane);
            ws_load_tile<B_MN>(st + TILE_DOUBLES, Bg, p.ldb, lane);
            cp_async_arrive_noinc(wb.full + s);
            Ag += a_step;
            Bg += b_step;
            if (++s == WS_STAGES) { s = 0; ph ^= 1u; }
        }
        cp_async_wait_all();
        return;
    }

    // ---------------- consumer warps ----------------
    const int wm = warp & 1, wn = warp >> 1;
    FragLane<A_MN, 64> fa;
    FragLane<B_MN, 32> fb;
    fa.init(wm, lane);
    fb.init(wn, lane);
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    {
        int s = 0;
        unsigned ph = 0;
        for (int kt = 0; kt < nk; ++kt) {
            mbar_wait(wb.full + s, ph);
            const double* st = smem + s * WS_STAGE_DOUBLES;
            ws_mma_stage<A_MN, B_MN>(st, st + TILE_DOUBLES, fa, fb, acc);
            __syncwarp();
            if (lane == 0) mbar_arrive(wb.empty + s);
            if (++s == WS_STAGES) { s = 0; ph ^= 1u; }
        }
    }
    const int g = lane >> 2, tig = lane & 3;
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int mb = 0; mb < 8; ++mb) {
        long row = (long)tm * TS + wm * 64 + mb * 8 + g;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            long col = (long)tn * TS + wn * 32 + nb * 8 + 2 * tig;
            double2* dst = reinterpret_cast<double2*>(p.C + row * p.ldc + col);
            double2 v;
            v.x = alpha * acc[mb][nb][0];
            v.y = alpha * acc[mb][nb][1];
            if (beta != 0.0) {
                double2 old = *dst;
                v.x = fma(beta, old.x, v.x);
                v.y = fma(beta, old.y, v.y);
            }
            *dst = v;
        }
    }
}

// host launcher (linalg.cu)
cudaError_t launch_dgemm(bool a_mn, bool b_mn, const GemmArgs& args, cudaStream_t stream);
cudaError_t dgemm_init();
void set_small_tile_threshold(int tiles128);

}  // namespace gp2d
