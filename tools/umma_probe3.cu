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

// K-major tile images.  mode 0: SWIZZLE_NONE [rows x 32 B]; 1: SWIZZLE_32B [rows x 32 B]; 3: SWIZZLE_64B [rows x 64 B]
// (two k-steps); 2: SWIZZLE_128B [rows x 128 B] (four k-steps).  k = byte along K inside the tile.
__host__ __device__ inline int tile_off(int mode, int r, int k) {
    if (mode == 0) return (r / 8) * 256 + (k / 16) * 128 + (r % 8) * 16 + (k % 16);
    if (mode == 1) return (r / 8) * 256 + (r % 8) * 32 + ((((k / 16) ^ ((r >> 2) & 1))) * 16) + (k % 16);
    if (mode == 3) return (r / 8) * 512 + (r % 8) * 64 + ((((k / 16) ^ ((r >> 1) & 3))) * 16) + (k % 16);
    return (r / 8) * 1024 + (r % 8) * 128 + ((((k / 16) ^ (r % 8))) * 16) + (k % 16);
}
__host__ __device__ inline int tile_kbytes(int mode) { return mode == 2 ? 128 : mode == 3 ? 64 : 32; }

// ---- correctness: D[128 x N] = sum_ks A(ks) B(ks)^T; ts = 1: A goes through TMEM -------------------------------
template <int N>
__global__ void __launch_bounds__(128, 1) probe_correct(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int ksteps, int mode,
                                                        int ts, int* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar;
    __shared__ unsigned tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int kb = tile_kbytes(mode), kpt = kb / 32;             // k-steps per tile image
    const int ntile = (ksteps + kpt - 1) / kpt;
    const int a_bytes = 128 * kb, b_bytes = N * kb;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)ntile * a_bytes;
    for (int i = tid; i < ntile * a_bytes; i += 128) sA[i] = (uint8_t)A[i];
    for (int i = tid; i < ntile * b_bytes; i += 128) sB[i] = (uint8_t)B[i];
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = tslot;
    if (tid == 0) {
        const unsigned idesc = make_idesc(128, N);
        for (int ks = 0; ks < ksteps; ++ks) {
            const unsigned aoff = (unsigned)((ks / kpt) * a_bytes + (ks % kpt) * 32), boff = (unsigned)((ks / kpt) * b_bytes + (ks % kpt) * 32);
            if (ts) {
                const unsigned acol = 256 + (ks & 1) * 8;
                tmem_cp_128x256b(tbase + acol, make_desc(smem_u32(sA) + aoff, mode));
                umma_i8_ts(tbase, tbase + acol, make_desc(smem_u32(sB) + boff, mode), idesc, ks > 0);
            } else {
                umma_i8(tbase, make_desc(smem_u32(sA) + aoff, mode), make_desc(smem_u32(sB) + boff, mode), idesc, ks > 0);
            }
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        int r[16];
        tmem_ld16(tbase + ((unsigned)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) out[(size_t)tid * N + c0 + c] = r[c];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

static int run_correct(int mode, int ts, int ksteps, int N) {
    const int kb = tile_kbytes(mode), kpt = kb / 32, ntile = (ksteps + kpt - 1) / kpt;
    const int a_bytes = 128 * kb, b_bytes = N * kb;
    std::vector<int8_t> hA((size_t)ntile * a_bytes, 0), hB((size_t)ntile * b_bytes, 0);
    std::vector<int8_t> la((size_t)128 * ksteps * 32), lb((size_t)N * ksteps * 32);
    srand(99 + mode * 7 + N + ts);
    for (auto& v : la) v = (int8_t)(rand() % 256 - 128);
    for (auto& v : lb) v = (int8_t)(rand() % 256 - 128);
    for (int ks = 0; ks < ksteps; ++ks) {
        for (int r = 0; r < 128; ++r)
            for (int k = 0; k < 32; ++k) hA[(size_t)(ks / kpt) * a_bytes + tile_off(mode, r, (ks % kpt) * 32 + k)] = la[((size_t)r * ksteps + ks) * 32 + k];
        for (int r = 0; r < N; ++r)
            for (int k = 0; k < 32; ++k) hB[(size_t)(ks / kpt) * b_bytes + tile_off(mode, r, (ks % kpt) * 32 + k)] = lb[((size_t)r * ksteps + ks) * 32 + k];
    }
    int8_t *dA, *dB;
    int* dO;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dO, (size_t)128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0xff, (size_t)128 * N * 4));
    const size_t smem = hA.size() + hB.size() + 1024;
    CK(cudaFuncSetAttribute(probe_correct<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_correct<64><<<1, 128, smem>>>(dA, dB, ksteps, mode, ts, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("correct mode=%d ts=%d: CUDA error %s\n", mode, ts, cudaGetErrorString(e)); exit(3); }
    std::vector<int> hO((size_t)128 * N);
    CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            long ref = 0;
            for (int kk = 0; kk < ksteps * 32; ++kk) ref += (long)la[(size_t)m * ksteps * 32 + kk] * (long)lb[(size_t)n * ksteps * 32 + kk];
            if ((long)hO[(size_t)m * N + n] != ref) {
                if (bad < 3) printf("    mismatch m=%d n=%d got %d want %ld\n", m, n, hO[(size_t)m * N + n], ref);
                ++bad;
            }
        }
    printf("correct mode=%d ts=%d ksteps=%d N=%d : %s (%ld of %d wrong)\n", mode, ts, ksteps, N, bad ? "FAIL" : "ok", bad, 128 * N);
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
    return bad != 0;
}

// ---- rates: S slices, S (S + 1) / 2 MMAs per k-step out of resident shared memory --------------------------------
// TS = 1: the S A-tiles of a k-step are first copied to TMEM (tcgen05.cp), double buffered by k-step parity
template <int N, int S, int MODE, int TS>
__global__ void __launch_bounds__(128, 1) probe_rate(int iters, long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long done;
    __shared__ unsigned tslot;
    constexpr int KB = MODE == 2 ? 128 : MODE == 3 ? 64 : 32, KPT = KB / 32;
    constexpr int A_TILE = 128 * KB, B_TILE = N * KB;
    constexpr int A_BYTES = S * A_TILE, STAGE = S * (A_TILE + B_TILE);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&done, 1); fence_barrier_init(); }
    for (int i = tid; i < STAGE; i += 128) smem[i] = (uint8_t)(i * 7 + 3);
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = tslot;
    long long t0 = clock64();
    if (tid == 32) {
        const unsigned idesc = make_idesc(128, N);
        const uint64_t ad0 = make_desc(smem_u32(smem), MODE), bd0 = make_desc(smem_u32(smem) + A_BYTES, MODE);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < KPT; ++ks) {
                const unsigned abuf = tbase + S * N + ((it * KPT + ks) & 1) * (S * 8);
                if (TS) {
#pragma unroll
                    for (int i = 0; i < S; ++i) tmem_cp_128x256b(abuf + i * 8, ad0 + (uint64_t)((i * A_TILE + ks * 32) >> 4));
                }
#pragma unroll
                for (int g = 0; g < S; ++g)
#pragma unroll
                    for (int i = 0; i <= g; ++i) {
                        if (TS) umma_i8_ts(tbase + g * N, abuf + i * 8, bd0 + (uint64_t)(((g - i) * B_TILE + ks * 32) >> 4), idesc, 1);
                        else umma_i8(tbase + g * N, ad0 + (uint64_t)((i * A_TILE + ks * 32) >> 4), bd0 + (uint64_t)(((g - i) * B_TILE + ks * 32) >> 4), idesc, 1);
                    }
            }
        }
        umma_commit(&done);
        mbar_wait(&done, 0);
    }
    __syncthreads();
    long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int N, int S, int MODE, int TS>
static void run_rate(int iters) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* cyc;
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    constexpr int KB = MODE == 2 ? 128 : MODE == 3 ? 64 : 32, KPT = KB / 32;
    constexpr int STAGE = S * (128 + N) * KB;
    const size_t smem = (size_t)STAGE + 1024;
    CK(cudaFuncSetAttribute(probe_rate<N, S, MODE, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe_rate<N, S, MODE, TS><<<sms, 128, smem>>>(iters / 4, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rate N=%d S=%d mode=%d ts=%d: CUDA error %s\n", N, S, MODE, TS, cudaGetErrorString(e)); exit(3); }
    CK(cudaEventRecord(e0));
    probe_rate<N, S, MODE, TS><<<sms, 128, smem>>>(iters, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const int pairs = S * (S + 1) / 2;
    const double ksteps = (double)iters * KPT, macs = (double)sms * ksteps * pairs * 128.0 * N * 32;
    printf("N=%3d S=%d mode=%d ts=%d: %7.3f ms  %6.0f clk/k-step  %5.1f clk/MMA  %6.1f int8 TOP/s  fp64-equivalent %5.1f TFLOP/s\n", N, S, MODE, TS, ms,
           ms * 1e-3 * 1.965e9 / ksteps, ms * 1e-3 * 1.965e9 / ksteps / pairs, 2.0 * macs / ms / 1e9, 2.0 * macs / pairs / ms / 1e9);
    cudaFree(cyc);
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int fails = 0;
    if (only < 0 || only == 0) {
        fails += run_correct(3, 0, 1, 64);
        fails += run_correct(3, 0, 5, 64);
        fails += run_correct(1, 0, 5, 64);
        fails += run_correct(0, 0, 5, 64);
    }
    if (only < 0 || only == 1) {
        run_rate<64, 6, 0, 0>(4000);
        run_rate<64, 6, 1, 0>(4000);
        run_rate<64, 6, 3, 0>(2000);
        run_rate<64, 6, 2, 0>(1000);
        run_rate<64, 7, 3, 0>(2000);
        run_rate<80, 6, 3, 0>(2000);
    }
    if (only < 0 || only == 2) {
        fails += run_correct(0, 1, 1, 64);
        fails += run_correct(0, 1, 5, 64);
        fails += run_correct(3, 1, 5, 64);
    }
    if (only < 0 || only == 3) {
        run_rate<64, 6, 0, 1>(4000);
        run_rate<64, 6, 3, 1>(2000);
        run_rate<64, 7, 0, 1>(4000);
        run_rate<80, 6, 0, 1>(4000);
    }
    printf("probe3 done, %d correctness failures\n", fails);
    return 0;
}
