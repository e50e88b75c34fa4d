// Second bring-up probe for the int8 tcgen05 path: the true issue rate of tcgen05.mma kind::i8 (M = 128) with the
// descriptors precomputed and the slice-pair loop fully unrolled (the first probe was bound by its own descriptor
// arithmetic: 132 clk per MMA for N = 64 and N = 128 alike), alone and fed by bulk copies from an L2-resident
// source, for the (N, S) shapes the predictive kernel can use (S accumulators of N columns must fit 512 TMEM columns).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe2 umma_probe2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(unsigned* smem_slot, unsigned ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void umma_i8_acc(unsigned taddr, uint64_t adesc, uint64_t bdesc, unsigned idesc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.eq.b32 p, 1, 1;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(unsigned saddr, int mode) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
    const uint64_t lbo = mode == 0 ? (128 >> 4) : 1, sbo = mode == 2 ? (1024 >> 4) : (256 >> 4);
    const uint64_t layout = mode == 0 ? 0 : mode == 1 ? 6 : 2;
    d |= lbo << 16;
    d |= sbo << 32;
    d |= 1ull << 46;
    d |= layout << 61;
    return d;
}
__host__ __device__ inline unsigned make_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

// One CTA per SM; thread 0 = copy producer, thread 32 = MMA issuer.  A stage is KS k-steps of 32: S A-tiles and
// S B-tiles per k-step.  MODE 0: SWIZZLE_NONE tile images [rows x 32 B] (4096 / N*32 bytes); MODE 2: SWIZZLE_128B
// tile images [rows x 128 B] holding 4 k-steps (KS must be 4).  what & 1: MMAs, what & 2: bulk copies.
template <int N, int S, int MODE, int KS>
__global__ void __launch_bounds__(128, 1) probe_rate(const uint8_t* __restrict__ src, size_t span, int iters, int what, int nstages,
                                                     long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long full[8], empty[8], done;
    __shared__ unsigned tslot;
    constexpr int A_TILE = MODE == 2 ? 128 * 128 : 4096, B_TILE = MODE == 2 ? N * 128 : N * 32;
    constexpr int TPS = MODE == 2 ? 1 : KS;                     // tile images per slice and stage
    constexpr int A_BYTES = S * TPS * A_TILE, B_BYTES = S * TPS * B_TILE, STAGE = A_BYTES + B_BYTES;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < 8; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(&done, 1);
        fence_barrier_init();
    }
    for (int i = tid; i < nstages * STAGE; i += 128) smem[i] = (uint8_t)(i * 7 + 3);
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = tslot;
    const bool do_mma = what & 1, do_copy = what & 2;
    long long t0 = clock64();
    if (tid == 0 && do_copy) {
        size_t off = ((size_t)blockIdx.x * 977 * STAGE) % (span - STAGE);
        off &= ~(size_t)1023;
        for (int it = 0; it < iters; ++it) {
            const int s = it % nstages;
            if (it >= nstages) mbar_wait(empty + s, ((it / nstages) - 1) & 1);
            mbar_arrive_expect_tx(full + s, STAGE);
            bulk_g2s(smem + (size_t)s * STAGE, src + off, A_BYTES, full + s);
            bulk_g2s(smem + (size_t)s * STAGE + A_BYTES, src + off + A_BYTES, B_BYTES, full + s);
            off += STAGE;
            if (off + STAGE > span) off = 0;
        }
    }
    if (tid == 32) {
        const unsigned idesc = make_idesc(128, N);
        const uint64_t ad0 = make_desc(smem_u32(smem), MODE), bd0 = make_desc(smem_u32(smem) + A_BYTES, MODE);
        int s = 0;
        unsigned ph = 0;
        for (int it = 0; it < iters; ++it) {
            if (do_copy) { mbar_wait(full + s, ph); tc_fence_after(); }
            if (do_mma) {
                const uint64_t ad = ad0 + (uint64_t)((s * STAGE) >> 4), bd = bd0 + (uint64_t)((s * STAGE) >> 4);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    // k-step offset inside a stage: next tile image (mode 0) or 32 bytes along the swizzled row (mode 2)
                    const int ka = MODE == 2 ? ks * 32 : ks * A_TILE, kb = MODE == 2 ? ks * 32 : ks * B_TILE;
#pragma unroll
                    for (int g = 0; g < S; ++g)
#pragma unroll
                        for (int i = 0; i <= g; ++i)
                            umma_i8_acc(tbase + g * N, ad + (uint64_t)((i * TPS * A_TILE + ka) >> 4),
                                        bd + (uint64_t)(((g - i) * TPS * B_TILE + kb) >> 4), idesc);
                }
                umma_commit(empty + s);
            } else {
                mbar_arrive(empty + s);
            }
            if (++s == nstages) { s = 0; ph ^= 1u; }
        }
        if (do_mma) { umma_commit(&done); mbar_wait(&done, 0); }
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int N, int S, int MODE, int KS>
static void run_rate(int what, size_t span_mb, int nstages, int iters) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t span = span_mb << 20;
    uint8_t* src;
    long long* cyc;
    CK(cudaMalloc(&src, span));
    CK(cudaMemset(src, 1, span));
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    constexpr int A_TILE = MODE == 2 ? 128 * 128 : 4096, B_TILE = MODE == 2 ? N * 128 : N * 32;
    constexpr int TPS = MODE == 2 ? 1 : KS;
    constexpr int STAGE = S * TPS * (A_TILE + B_TILE);
    const size_t smem = (size_t)nstages * STAGE + 1024;
    if (smem > 227 * 1024) { printf("skip N=%d S=%d mode=%d: %zu bytes of shared memory\n", N, S, MODE, smem); return; }
    CK(cudaFuncSetAttribute(probe_rate<N, S, MODE, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe_rate<N, S, MODE, KS><<<sms, 128, smem>>>(src, span, iters / 4, what, nstages, cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe_rate<N, S, MODE, KS><<<sms, 128, smem>>>(src, span, iters, what, nstages, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    const int pairs = S * (S + 1) / 2;
    const double bytes = (double)sms * iters * STAGE, macs = (double)sms * iters * KS * pairs * 128.0 * N * 32;
    printf("N=%3d S=%d mode=%d ks/stage=%d stages=%d mma=%d copy=%d span=%4zu MB: %7.3f ms  %6.0f clk/k-step", N, S, MODE, KS, nstages,
           what & 1, (what >> 1) & 1, span_mb, ms, (double)mx / iters / KS);
    if (what & 2) printf("  %5.2f TB/s = %5.1f B/clk/SM", bytes / ms / 1e9, (double)STAGE * iters / (double)mx);
    if (what & 1) printf("  %6.1f int8 TOP/s  %5.1f clk/MMA  fp64-equivalent %5.1f TFLOP/s", 2.0 * macs / ms / 1e9, (double)mx / iters / KS / pairs,
                         2.0 * macs / pairs / ms / 1e9);
    printf("\n");
    cudaFree(src); cudaFree(cyc);
}

int main() {
    // MMA alone
    run_rate<64, 6, 0, 1>(1, 64, 4, 4000);
    run_rate<64, 7, 0, 1>(1, 64, 4, 4000);
    run_rate<80, 6, 0, 1>(1, 64, 4, 4000);
    run_rate<96, 5, 0, 1>(1, 64, 4, 4000);
    run_rate<128, 4, 0, 1>(1, 64, 4, 4000);
    run_rate<256, 2, 0, 1>(1, 64, 4, 4000);
    run_rate<64, 6, 2, 4>(1, 64, 1, 2000);
    run_rate<128, 4, 2, 4>(1, 64, 1, 2000);
    run_rate<256, 2, 2, 4>(1, 64, 1, 2000);
    // fed from L2 (64 MB footprint) and from DRAM (2 GB)
    run_rate<64, 6, 0, 1>(3, 64, 6, 4000);
    run_rate<64, 6, 0, 2>(3, 64, 3, 2000);
    run_rate<64, 7, 0, 1>(3, 64, 5, 4000);
    run_rate<80, 6, 0, 1>(3, 64, 5, 4000);
    run_rate<96, 5, 0, 1>(3, 64, 5, 4000);
    run_rate<128, 4, 0, 1>(3, 64, 6, 4000);
    run_rate<64, 6, 0, 1>(3, 2048, 6, 4000);
    printf("probe2 done\n");
    return 0;
}
