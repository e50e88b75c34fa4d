// Third bring-up probe for the int8 tcgen05 path.  SS-mode MMAs (both operands in shared memory) cannot go below 64 clk
// per instruction at M = 128 (the A tile, 4 KB, is re-read for every instruction), which halves the rate at N = 64.
// This probe checks (a) A-from-TMEM (tcgen05.cp 128x256b of the shared-memory tile image, then the TS form of the MMA)
// against a CPU product, (b) its issue rate with S copies + S (S + 1) / 2 MMAs per k-step, (c) the SS rate and the
// correctness of the 32-byte and 64-byte swizzled tile images.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe3 umma_probe3.cu
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
    const uint64_t lbo = mode == 0 ? (128 >> 4) : 1, sbo = mode == 2 ? (1024 >> 4) : mode == 3 ? (512 >> 4) : (256 >> 4);
    const uint64_t layout = mode == 0 ? 0 : mode == 1 ? 6 : mode == 3 ? 4 : 2;
    d |= lbo << 16;
    d |= sbo << 32;
    d |= 1ull << 46;
    d |= layout << 61;
    return d;
}
__host__ __device__ inline unsigned make_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}


__device__ __forceinline__ void umma_i8(unsigned taddr, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_i8_ts(unsigned taddr, unsigned a_taddr, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(taddr), "r"(a_taddr), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(unsigned taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;\n" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }


// Fourth probe: where does the time of the TS form (A from TMEM) go?  VAR 0: S copies then S (S + 1) / 2 TS MMAs per
// k-step (probe3's order); 1: TS MMAs only (A stale in TMEM); 2: copies only; 3: copies of k-step + 1 issued before the
// MMAs of k-step; 4: SS MMAs; 5: A written by four warps with tcgen05.st (32x32b.x8 per slice) and handed to the issuer
// through mbarriers, two buffers.
template <int N, int S, int VAR>
__global__ void __launch_bounds__(192, 1) probe_rate4(int iters, long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long done, afull[2], aempty[2];
    __shared__ unsigned tslot;
    constexpr int A_TILE = 128 * 32, B_TILE = N * 32;
    constexpr int A_BYTES = S * A_TILE, STAGE = S * (A_TILE + B_TILE);
    constexpr int ACOLS = (S * N + 2 * S * 8 <= 512) ? S * 8 : (512 - S * N) / 2;     // columns of one A buffer
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&done, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(afull + b, 4); mbar_init(aempty + b, 1); }
        fence_barrier_init();
    }
    for (int i = tid; i < STAGE; i += 192) smem[i] = (uint8_t)(i * 7 + 3);
    if (warp == 4) tmem_alloc(&tslot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = tslot;
    long long t0 = clock64();
    if (tid == 128) {
        const unsigned idesc = make_idesc(128, N);
        const uint64_t ad0 = make_desc(smem_u32(smem), 0), bd0 = make_desc(smem_u32(smem) + A_BYTES, 0);
        if (VAR == 3) {
#pragma unroll
            for (int i = 0; i < S; ++i) tmem_cp_128x256b(tbase + S * N + (i * 8) % ACOLS, ad0 + (uint64_t)((i * A_TILE) >> 4));
        }
        for (int it = 0; it < iters; ++it) {
            const unsigned abuf = tbase + S * N + (it & 1) * ACOLS;
            const unsigned anext = tbase + S * N + ((it + 1) & 1) * ACOLS;
            if (VAR == 0 || VAR == 2) {
#pragma unroll
                for (int i = 0; i < S; ++i) tmem_cp_128x256b(abuf + (i * 8) % ACOLS, ad0 + (uint64_t)((i * A_TILE) >> 4));
            }
            if (VAR == 3) {
#pragma unroll
                for (int i = 0; i < S; ++i) tmem_cp_128x256b(anext + (i * 8) % ACOLS, ad0 + (uint64_t)((i * A_TILE) >> 4));
            }
            if (VAR == 5) {
                mbar_wait(afull + (it & 1), (unsigned)(it >> 1) & 1u);
                tc_fence_after();
            }
            if (VAR != 2) {
#pragma unroll
                for (int g = 0; g < S; ++g)
#pragma unroll
                    for (int i = 0; i <= g; ++i) {
                        if (VAR == 4) umma_i8(tbase + g * N, ad0 + (uint64_t)((i * A_TILE) >> 4), bd0 + (uint64_t)(((g - i) * B_TILE) >> 4), idesc, 1);
                        else umma_i8_ts(tbase + g * N, abuf + (i * 8) % ACOLS, bd0 + (uint64_t)(((g - i) * B_TILE) >> 4), idesc, 1);
                    }
            }
            if (VAR == 5) umma_commit(aempty + (it & 1));
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
    } else if (VAR == 5 && warp < 4) {
        // each thread owns one TMEM lane (row of A): 8 words per slice
        unsigned v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = tid * 8 + j;
        for (int it = 0; it < iters; ++it) {
            const int b = it & 1;
            if (it >= 2) { mbar_wait(aempty + b, (unsigned)((it >> 1) - 1) & 1u); tc_fence_after(); }
            const unsigned abuf = tbase + ((unsigned)(warp * 32) << 16) + S * N + b * ACOLS;
#pragma unroll
            for (int i = 0; i < S; ++i)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(abuf + (i * 8) % ACOLS), "r"(v[0]),
                             "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(afull + b);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tbase, 512);
}

template <int N, int S, int VAR>
static void run_rate4(int iters) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* cyc;
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    constexpr int STAGE = S * (128 + N) * 32;
    const size_t smem = (size_t)STAGE + 1024;
    CK(cudaFuncSetAttribute(probe_rate4<N, S, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe_rate4<N, S, VAR><<<sms, 192, smem>>>(iters / 4, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rate4 N=%d S=%d var=%d: CUDA error %s\n", N, S, VAR, cudaGetErrorString(e)); exit(3); }
    CK(cudaEventRecord(e0));
    probe_rate4<N, S, VAR><<<sms, 192, smem>>>(iters, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long c0 = 0;
    CK(cudaMemcpy(&c0, cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    const int pairs = S * (S + 1) / 2;
    printf("N=%3d S=%d var=%d: %7.3f ms  %6.0f clk/k-step (clock64: %6.0f)  %5.1f clk/MMA  %5.2f clk/column\n", N, S, VAR, ms,
           ms * 1e-3 * 1.965e9 / iters, (double)c0 / iters, ms * 1e-3 * 1.965e9 / iters / pairs, ms * 1e-3 * 1.965e9 / iters / N);
    cudaFree(cyc);
}


// Fifth question: the ~47 clk floor per MMA whatever N and whatever the A source.  Is it the read-after-write distance on
// the accumulator?  D consecutive MMAs go to D different accumulators (round robin), TS (A stale in TMEM) or SS.
template <int N, int D, int TS>
__global__ void __launch_bounds__(192, 1) probe_dist(int iters, long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long done;
    __shared__ unsigned tslot;
    constexpr int S = 6, A_TILE = 128 * 32, B_TILE = N * 32, A_BYTES = S * A_TILE, STAGE = S * (A_TILE + B_TILE);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&done, 1); fence_barrier_init(); }
    for (int i = tid; i < STAGE; i += 192) smem[i] = (uint8_t)(i * 7 + 3);
    if (warp == 4) tmem_alloc(&tslot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = tslot;
    if (tid == 128) {
        const unsigned idesc = make_idesc(128, N);
        const uint64_t ad0 = make_desc(smem_u32(smem), 0), bd0 = make_desc(smem_u32(smem) + A_BYTES, 0);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int n = 0; n < 24; ++n) {
                const int g = n % D, i = n % S, j = (n / 2) % S;
                if (TS) umma_i8_ts(tbase + g * N, tbase + S * N + (i * 8) % 32, bd0 + (uint64_t)((j * B_TILE) >> 4), idesc, 1);
                else umma_i8(tbase + g * N, ad0 + (uint64_t)((i * A_TILE) >> 4), bd0 + (uint64_t)((j * B_TILE) >> 4), idesc, 1);
            }
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
    }
    __syncthreads();
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tbase, 512);
}
template <int N, int D, int TS>
static void run_dist(int iters) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* cyc;
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    const size_t smem = (size_t)6 * (128 + N) * 32 + 1024;
    CK(cudaFuncSetAttribute(probe_dist<N, D, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe_dist<N, D, TS><<<sms, 192, smem>>>(iters / 4, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("dist N=%d D=%d ts=%d: CUDA error %s\n", N, D, TS, cudaGetErrorString(e)); exit(3); }
    CK(cudaEventRecord(e0));
    probe_dist<N, D, TS><<<sms, 192, smem>>>(iters, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("N=%3d distance=%d ts=%d: %7.3f ms  %5.1f clk/MMA (tensor time at 8192 MAC/clk: %d)\n", N, D, TS, ms, ms * 1e-3 * 1.965e9 / iters / 24, N / 2);
    cudaFree(cyc);
}

// Sixth question: the production kernel issues the same 21 SS MMAs per k-step at ~68 clk each, this probe at 55.  Which
// difference matters?  bit 0: a commit to an mbarrier after every k-step; bit 1: five stage buffers in rotation;
// bit 2: high-entropy operand bytes; bit 3: other warps spinning on an mbarrier (try_wait) meanwhile.
template <int N, int FLAGS>
__global__ void __launch_bounds__(512, 1) probe_env(int iters, long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long done, stagebar[5], never;
    __shared__ unsigned tslot;
    constexpr int S = 6, A_TILE = 128 * 32, B_TILE = N * 32, A_BYTES = S * A_TILE, STAGE = S * (A_TILE + B_TILE);
    constexpr int NST = (FLAGS & 2) ? 5 : 1;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&done, 1); mbar_init(&never, 1); for (int i = 0; i < 5; ++i) mbar_init(stagebar + i, 1); fence_barrier_init(); }
    unsigned x = tid * 2654435761u + blockIdx.x * 40503u + 12345u;
    for (int i = tid; i < NST * STAGE; i += 512) {
        x = x * 1664525u + 1013904223u;
        smem[i] = (FLAGS & 4) ? (uint8_t)(x >> 24) : (uint8_t)(i * 7 + 3);
    }
    if (warp == 4) tmem_alloc(&tslot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = tslot;
    if (tid == 128) {
        const unsigned idesc = make_idesc(128, N);
        for (int it = 0; it < iters; ++it) {
            const int st = (FLAGS & 2) ? it % 5 : 0;
            const uint64_t ad0 = make_desc(smem_u32(smem) + st * STAGE, 0), bd0 = make_desc(smem_u32(smem) + st * STAGE + A_BYTES, 0);
#pragma unroll
            for (int g = 0; g < S; ++g)
#pragma unroll
                for (int i = 0; i <= g; ++i)
                {
                    umma_i8(tbase + g * N, ad0 + (uint64_t)((i * A_TILE) >> 4), bd0 + (uint64_t)(((g - i) * B_TILE) >> 4), idesc, 1);
                    if (FLAGS & 32) {              // ~40 clk of dependent integer work after every MMA
                        unsigned z = it + g;
#pragma unroll
                        for (int q = 0; q < 10; ++q) asm volatile("mad.lo.u32 %0, %0, 3, 1;" : "+r"(z));
                        if (z == 0x12345u) cycles[1] = z;
                    }
                }
            if (FLAGS & 1) umma_commit(stagebar + st);
            if (FLAGS & 16) {                      // ~200 clk of dependent integer work between two k-steps
                unsigned z = it;
#pragma unroll
                for (int q = 0; q < 50; ++q) asm volatile("mad.lo.u32 %0, %0, 3, 1;" : "+r"(z));
                if (z == 0x12345u) cycles[1] = z;
            }
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
        if (FLAGS & 8) mbar_arrive(&never);
    } else if ((FLAGS & 8) && warp != 4) {
        mbar_wait(&never, 0);
    }
    __syncthreads();
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tbase, 512);
}
template <int N, int FLAGS>
static void run_env(int iters) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* cyc;
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    const size_t smem = (size_t)((FLAGS & 2) ? 5 : 1) * 6 * (128 + N) * 32 + 1024;
    CK(cudaFuncSetAttribute(probe_env<N, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe_env<N, FLAGS><<<sms, 512, smem>>>(iters / 4, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("env N=%d flags=%d: CUDA error %s\n", N, FLAGS, cudaGetErrorString(e)); exit(3); }
    CK(cudaEventRecord(e0));
    probe_env<N, FLAGS><<<sms, 512, smem>>>(iters, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("N=%3d flags=%2d (commit %d, 5 stages %d, random bytes %d, spinning warps %d, gap per k-step %d, gap per MMA %d): %7.3f ms  %6.0f clk/k-step at 1965 MHz  %5.1f clk/MMA\n", N, FLAGS,
           FLAGS & 1, (FLAGS >> 1) & 1, (FLAGS >> 2) & 1, (FLAGS >> 3) & 1, (FLAGS >> 4) & 1, (FLAGS >> 5) & 1, ms, ms * 1e-3 * 1.965e9 / iters, ms * 1e-3 * 1.965e9 / iters / 21);
    cudaFree(cyc);
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    if (only < 0 || only == 0) {
        run_rate4<64, 6, 4>(4000);
        run_rate4<64, 6, 0>(4000);
        run_rate4<64, 6, 1>(4000);
        run_rate4<64, 6, 2>(4000);
        run_rate4<64, 6, 3>(4000);
    }
    if (only < 0 || only == 1) {
        run_rate4<64, 6, 5>(4000);
    }
    if (only < 0 || only == 2) {
        run_rate4<80, 6, 4>(4000);
        run_rate4<80, 6, 1>(4000);
        run_rate4<64, 7, 1>(4000);
        run_rate4<32, 6, 1>(4000);
        run_rate4<32, 6, 4>(4000);
        run_rate4<128, 3, 1>(4000);
        run_rate4<128, 3, 4>(4000);
    }
    if (only < 0 || only == 3) {
        run_dist<64, 1, 1>(4000); run_dist<64, 2, 1>(4000); run_dist<64, 3, 1>(4000); run_dist<64, 4, 1>(4000); run_dist<64, 6, 1>(4000);
        run_dist<64, 1, 0>(4000); run_dist<64, 2, 0>(4000); run_dist<64, 3, 0>(4000); run_dist<64, 6, 0>(4000);
        run_dist<80, 1, 1>(4000); run_dist<80, 2, 1>(4000); run_dist<80, 3, 1>(4000); run_dist<80, 6, 1>(4000);
        run_dist<80, 2, 0>(4000); run_dist<80, 6, 0>(4000);
        run_dist<32, 6, 1>(4000); run_dist<32, 6, 0>(4000); run_dist<16, 6, 1>(4000);
    }
    if (only < 0 || only == 4) {
        run_env<80, 0>(20000); run_env<80, 1>(20000); run_env<80, 2>(20000); run_env<80, 4>(20000); run_env<80, 8>(20000);
        run_env<80, 7>(20000); run_env<80, 15>(20000);
    }
    if (only < 0 || only == 5) {
        run_env<80, 0>(20000); run_env<80, 16>(20000); run_env<80, 32>(20000); run_env<80, 48>(20000);
    }
    printf("probe4 done\n");
    return 0;
}
