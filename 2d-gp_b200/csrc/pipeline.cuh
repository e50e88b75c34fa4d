// Warp-specialised operand pipeline for the DMMA kernels (GEMM, fused predict).
//
// A CTA is 8 consumer warps (2 along M x 4 along N over a 128 x 128 tile, 64 x 32 each) plus
// ONE producer warp.  The producer streams 128 x 16 operand tiles into a ring of shared-memory
// stages with cp.async and signals each stage through an mbarrier
// (cp.async.mbarrier.arrive.noinc: the barrier flips when the copies have landed); consumers
// wait on that barrier, run 128 DMMAs per warp from the stage and release it through a second
// mbarrier.  No __syncthreads in the k-loop: the consumers never execute address arithmetic
// for the copies, and their own fragment addresses are per-lane constants computed once
// (round-1 profile: ~150 integer instructions + a CTA barrier per 128 DMMAs cost 14 % of the
// FP64 tensor pipe in the single-role kernel).
//
// Shared-memory layouts are those of common.cuh (XOR-swizzled, conflict-free LDS.64):
//   K-major  tile T[128][16]: element (r,k) at r*16 + 4*((k>>2) ^ (r&3)) + (k&3)
//   MN-major tile T[16][128]: element (k,r) at k*128 + ((((r>>1) ^ ((k&3)<<1)) << 1) | (r&1))
// (the first is algebraically the kmaj_off of common.cuh).
#pragma once
#include "common.cuh"

namespace gp2d {

constexpr int WS_STAGES = 6;                                  // 6 x 32 KB ring
constexpr int WS_CONSUMERS = 256;                             // 8 warps
constexpr int WS_THREADS = WS_CONSUMERS + 32;                 // + 1 producer warp
constexpr int WS_STAGE_DOUBLES = 2 * TILE_DOUBLES;            // A tile then B tile
constexpr int WS_RING_BYTES = WS_STAGES * WS_STAGE_DOUBLES * (int)sizeof(double);   // 196 608
constexpr int WS_BAR_BYTES = 2 * WS_STAGES * 8;               // full[], empty[]

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one non-blocking look: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
// the barrier receives one arrival when all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// bulk (TMA, 1-D) copy of a contiguous, 16-byte aligned block; completion is counted in bytes on
// the mbarrier, which the issuing thread first arms with mbar_arrive_expect_tx
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// orders this thread's generic-proxy writes before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

struct WsBarriers {
    unsigned long long* full;    // [WS_STAGES]; arrivals: 32 producer lanes (cp.async completion) or 1 (bulk)
    unsigned long long* empty;   // [WS_STAGES], 8 arrivals (one lane per consumer warp)
    __device__ __forceinline__ void init(int tid, int full_count = 32) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < WS_STAGES; ++s) { mbar_init(full + s, full_count); mbar_init(empty + s, WS_CONSUMERS / 32); }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
    }
};

// ---- producer: one warp copies a 128 x 16 operand tile ------------------------------------
// K-major source: element (r,k) at g[r*ld + k];  MN-major source: element (k,r) at g[k*ld + r].
template <bool MNMAJ>
__device__ __forceinline__ void ws_load_tile(double* smem, const double* g, long ld, int lane) {
    if (!MNMAJ) {
        // chunk q = lane + 32 i: row (lane>>3) + 4 i, 16-byte chunk lane&7
        const int r0 = lane >> 3, ch = lane & 7;
        double* d = smem + r0 * BK + ((ch ^ (r0 << 1)) << 1);
        const double* s = g + (long)r0 * ld + (ch << 1);
        const long step = 4 * ld;
#pragma unroll
        for (int i = 0; i < 32; ++i) { cp_async16(d + 64 * i, s); s += step; }
    } else {
        // chunk q = lane + 32 i: k = i>>1, 16-byte chunk lane + 32 (i&1) of 64
        const double* s = g + (lane << 1);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int k = i >> 1;
            cp_async16(smem + k * TILE + 64 * (i & 1) + ((lane ^ ((k & 3) << 1)) << 1), s + 64 * (i & 1));
            if (i & 1) s += ld;
        }
    }
}

// ---- consumer: per-lane fragment addresses -------------------------------------------------
// W = rows of the tile owned by one warp along this operand (64 for A, 32 for B).
template <bool MNMAJ, int W>
struct FragLane {
    int base, x1, x2, x3;   // K-major: base + {0,x1,x2,x3}[k>>2] + 128*blk ; MN-major: even/odd bases
    __device__ __forceinline__ void init(int w, int lane) {
        const int g = lane >> 2, tig = lane & 3;
        if (!MNMAJ) {
            const int z = g & 3;
            base = (w * W + g) * BK + tig + 4 * z;        // k-group 0: 4*(0^z)
            x1 = 4 * ((1 ^ z) - z);
            x2 = 4 * ((2 ^ z) - z);
            x3 = 4 * ((3 ^ z) - z);
        } else {
            const int t1 = tig >> 1;
            const int b = tig * TILE + w * W + 2 * ((g >> 1) ^ ((tig & 1) << 1)) + (g & 1);
            base = b + 8 * t1;                             // even blocks: (blk ^ t1) * 8 = blk*8 + 8 t1
            x1 = b - 8 * t1;                               // odd blocks:  (blk ^ t1) * 8 = blk*8 - 8 t1
            x2 = x3 = 0;
        }
    }
    // element of block `blk` (8 rows) for k-group q (k = 4q + tig)
    template <int Q, int BLK>
    __device__ __forceinline__ double ld(const double* tile) const {
        if (!MNMAJ) {
            const int x = Q == 0 ? 0 : Q == 1 ? x1 : Q == 2 ? x2 : x3;
            return tile[base + x + BLK * 8 * BK];
        } else {
            return tile[((BLK & 1) ? x1 : base) + Q * 4 * TILE + BLK * 8];
        }
    }
};

template <bool A_MN, bool B_MN, int Q>
__device__ __forceinline__ void ws_mma_kgroup(const double* As, const double* Bs, const FragLane<A_MN, 64>& fa,
                                              const FragLane<B_MN, 32>& fb, double (&acc)[8][4][2]) {
    double a[8], b[4];
    a[0] = fa.template ld<Q, 0>(As); a[1] = fa.template ld<Q, 1>(As);
    a[2] = fa.template ld<Q, 2>(As); a[3] = fa.template ld<Q, 3>(As);
    a[4] = fa.template ld<Q, 4>(As); a[5] = fa.template ld<Q, 5>(As);
    a[6] = fa.template ld<Q, 6>(As); a[7] = fa.template ld<Q, 7>(As);
    b[0] = fb.template ld<Q, 0>(Bs); b[1] = fb.template ld<Q, 1>(Bs);
    b[2] = fb.template ld<Q, 2>(Bs); b[3] = fb.template ld<Q, 3>(Bs);
#pragma unroll
    for (int mb = 0; mb < 8; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
}

// same, but only the first 4 row blocks (32 rows) of the warp tile: the ragged last row block
template <bool A_MN, bool B_MN, int Q>
__device__ __forceinline__ void ws_mma_kgroup_half(const double* As, const double* Bs, const FragLane<A_MN, 64>& fa,
                                                   const FragLane<B_MN, 32>& fb, double (&acc)[8][4][2]) {
    double a[4], b[4];
    a[0] = fa.template ld<Q, 0>(As); a[1] = fa.template ld<Q, 1>(As);
    a[2] = fa.template ld<Q, 2>(As); a[3] = fa.template ld<Q, 3>(As);
    b[0] = fb.template ld<Q, 0>(Bs); b[1] = fb.template ld<Q, 1>(Bs);
    b[2] = fb.template ld<Q, 2>(Bs); b[3] = fb.template ld<Q, 3>(Bs);
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
}
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void ws_mma_stage_half(const double* As, const double* Bs, const FragLane<A_MN, 64>& fa,
                                                  const FragLane<B_MN, 32>& fb, double (&acc)[8][4][2]) {
    ws_mma_kgroup_half<A_MN, B_MN, 0>(As, Bs, fa, fb, acc);
    ws_mma_kgroup_half<A_MN, B_MN, 1>(As, Bs, fa, fb, acc);
    ws_mma_kgroup_half<A_MN, B_MN, 2>(As, Bs, fa, fb, acc);
    ws_mma_kgroup_half<A_MN, B_MN, 3>(As, Bs, fa, fb, acc);
}

// one 128 x 128 x 16 stage: 128 DMMAs per warp
template <bool A_MN, bool B_MN>
__device__ __forceinline__ void ws_mma_stage(const double* As, const double* Bs, const FragLane<A_MN, 64>& fa,
                                             const FragLane<B_MN, 32>& fb, double (&acc)[8][4][2]) {
    ws_mma_kgroup<A_MN, B_MN, 0>(As, Bs, fa, fb, acc);
    ws_mma_kgroup<A_MN, B_MN, 1>(As, Bs, fa, fb, acc);
    ws_mma_kgroup<A_MN, B_MN, 2>(As, Bs, fa, fb, acc);
    ws_mma_kgroup<A_MN, B_MN, 3>(As, Bs, fa, fb, acc);
}

}  // namespace gp2d
