// Shared device helpers for the gp2d sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gp2d {

constexpr int TILE = 128;      // every internal matrix dimension is padded to a multiple of this
constexpr int BK = 16;         // k-depth of one shared-memory stage (doubles)
constexpr int NTHREADS = 256;  // CTA of the leaf / probe kernels (the GEMM-shaped kernels define their own shapes)
// CTA shapes: NT threads = (NT/128) warps along M x 4 warps along N over a 128 x 128 tile;
// a warp owns MB(NT) x 4 DMMA accumulator blocks (8 x 8 each): 64x32 at NT=256, 32x32 at NT=512.
__host__ __device__ constexpr int warps_m(int nt) { return nt / 128; }
__host__ __device__ constexpr int mblocks(int nt) { return 16 / warps_m(nt); }
// General form: a TS x TS CTA tile (TS = 128, or 64 for small problems) with NT threads:
// TS/32 warps along N (32 columns each), the rest along M.
template <int TS_, int NT_>
struct TileCfg {
    static constexpr int TS = TS_, NT = NT_;
    static constexpr int WN = TS / 32;
    static constexpr int WM = (NT / 32) / WN;
    static constexpr int MB = TS / (8 * WM);
    static constexpr int STAGE_DOUBLES = TS * BK;
};
constexpr int TILE_DOUBLES = TILE * BK;   // 2048 doubles = 16 KB per operand stage

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Function attributes (opt-in shared memory) are per device: remember which devices were set up.
// Racing first calls repeat an idempotent cudaFuncSetAttribute, nothing worse.
struct PerDeviceOnce {
    bool done[64] = {};
    // returns the device slot to initialise, or -1 when this device is already set up
    int pending() const {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
        return done[dev] ? -1 : dev;
    }
};

// ---- cp.async (LDGSTS) -------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- FP64 tensor pipe: DMMA.8x8x4 (the only native f64 MMA shape on sm_100a) --------
// A 8x4 row-major fragment: lane holds A[lane>>2][lane&3]
// B 4x8 col-major fragment: lane holds B[lane&3][lane>>2]
// C 8x8: lane holds C[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- shared-memory operand tiles -----------------------------------------------------
// K-major tile  : T[128 rows(m or n)][16 k]   (element (r,k)); 16-byte chunk index k>>1 is
//                 XOR-swizzled with ((r&3)<<1) so the 4 rows x 4 k of a half-warp fragment
//                 read hit 16 distinct 8-byte bank pairs.
// MN-major tile : T[16 k][128 (m or n)]       (element (k,r)); chunk index r>>1 is swizzled
//                 with ((k&3)<<1).
__device__ __forceinline__ int kmaj_off(int r, int k) {
    return r * BK + ((((k >> 1) ^ ((r & 3) << 1)) << 1) | (k & 1));
}
template <int TS = TILE>
__device__ __forceinline__ int mnmaj_off(int k, int r) {
    return k * TS + ((((r >> 1) ^ ((k & 3) << 1)) << 1) | (r & 1));
}

// Load one 128x16 operand tile with 16-byte cp.async, 4 chunks per thread.
// KMAJ: global element (r,k) at g[r*ld + k];  MN-major: global element (k,r) at g[k*ld + r].
template <bool MNMAJ, int NT = NTHREADS, int TS = TILE>
__device__ __forceinline__ void load_tile_async(double* smem, const double* g, long ld, int tid) {
#pragma unroll
    for (int it = 0; it < TS * 8 / NT; ++it) {
        int q = tid + it * NT;
        if (!MNMAJ) {
            int r = q >> 3, ch = q & 7;
            cp_async16(smem + r * BK + ((ch ^ ((r & 3) << 1)) << 1), g + (long)r * ld + (ch << 1));
        } else {
            int k = q / (TS / 2), ch = q % (TS / 2);
            cp_async16(smem + k * TS + ((ch ^ ((k & 3) << 1)) << 1), g + (long)k * ld + (ch << 1));
        }
    }
}

// One BK=16 stage of DMMAs for a (8*MB)x32 warp tile.  acc[mb][nb][2].
template <bool A_MN, bool B_MN, int MB, int TS = TILE>
__device__ __forceinline__ void mma_stage(const double* As, const double* Bs, int wm, int wn,
                                          int lane, double (&acc)[MB][4][2]) {
    const int g = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
        double a[MB], b[4];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
            int m = wm * (MB * 8) + mb * 8 + g;
            a[mb] = A_MN ? As[mnmaj_off<TS>(kk + tig, m)] : As[kmaj_off(m, kk + tig)];
        }
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            int n = wn * 32 + nb * 8 + g;
            b[nb] = B_MN ? Bs[mnmaj_off<TS>(kk + tig, n)] : Bs[kmaj_off(n, kk + tig)];
        }
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
    }
}

}  // namespace gp2d
