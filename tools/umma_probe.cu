// Bring-up probe for the int8 tcgen05 path (Ozaki-split predictive GEMM): checks the shared-memory
// operand layouts / descriptors of tcgen05.mma kind::i8 against a CPU product, and measures what bounds the
// kernel design -- MMA issue rate, bulk-copy (L2 -> shared memory) bandwidth per SM under full-chip load,
// and both together.  Standalone:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
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
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- tcgen05 ---------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(unsigned* smem_slot, unsigned ncols) {      // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {          // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void umma_i8(unsigned taddr, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(taddr), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// K-major operand tile [rows x 32 bytes of K]; layout modes
//   0  SWIZZLE_NONE : 8-row x 16-byte core matrices; offset = (r/8) 256 + (k/16) 128 + (r%8) 16 + k%16   (LBO 128, SBO 256)
//   1  SWIZZLE_32B  : rows of 32 bytes;              offset = (r/8) 256 + (r%8) 32 + (((k/16) ^ ((r>>2)&1)) 16) + k%16   (SBO 256)
//   2  SWIZZLE_128B : rows of 128 bytes (4 k-steps); offset = (r/8) 1024 + (r%8) 128 + (((k/16) ^ (r%8)) 16) + k%16       (SBO 1024)
__host__ __device__ inline int tile_off(int mode, int r, int k) {
    if (mode == 0) return (r / 8) * 256 + (k / 16) * 128 + (r % 8) * 16 + (k % 16);
    if (mode == 1) return (r / 8) * 256 + (r % 8) * 32 + ((((k / 16) ^ ((r >> 2) & 1))) * 16) + (k % 16);
    return (r / 8) * 1024 + (r % 8) * 128 + ((((k / 16) ^ (r % 8))) * 16) + (k % 16);
}
__device__ __forceinline__ uint64_t make_desc(unsigned saddr, int mode) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
    const uint64_t lbo = mode == 0 ? (128 >> 4) : 1, sbo = mode == 2 ? (1024 >> 4) : (256 >> 4);
    const uint64_t layout = mode == 0 ? 0 : mode == 1 ? 6 : 2;
    d |= lbo << 16;
    d |= sbo << 32;
    d |= 1ull << 46;          // descriptor version (Blackwell)
    d |= layout << 61;
    return d;
}
__host__ __device__ inline unsigned make_idesc(int M, int N) {
    // c_format S32 (2) bits [4,6); a_format / b_format INT8 signed (1) bits [7,10) / [10,13); K-major both;
    // n_dim = N >> 3 bits [17,23); m_dim = M >> 4 bits [24,29)
    return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

// ---- test 1: correctness ----------------------------------------------------------------------
// D[128 x N] = sum over ksteps and pairs (i, j), i + j == g, of A_i[128 x 32] B_j[N x 32]^T into accumulator g.
// A: [S][ksteps][tile image 4096 B]; B: [S][ksteps][tile image N*32 B]; out: [S][128][N] int32
template <int N>
__global__ void __launch_bounds__(128, 1) probe_correct(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int S, int ksteps,
                                                        int mode, int* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar;
    __shared__ unsigned tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int a_bytes = mode == 2 ? 128 * 128 : 4096, b_bytes = mode == 2 ? N * 128 : N * 32;
    const int ks_tiles = mode == 2 ? (ksteps + 3) / 4 : ksteps;          // mode 2 holds 4 k-steps per tile
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)S * ks_tiles * a_bytes;
    for (int i = tid; i < S * ks_tiles * a_bytes; i += 128) sA[i] = (uint8_t)A[i];
    for (int i = tid; i < S * ks_tiles * b_bytes; i += 128) sB[i] = (uint8_t)B[i];
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = tslot;
    if (tid == 0) {
        const unsigned idesc = make_idesc(128, N);
        for (int g = 0; g < S; ++g) {
            unsigned acc = 0;
            for (int ks = 0; ks < ksteps; ++ks)
                for (int i = 0; i <= g; ++i) {
                    const int j = g - i;
                    const unsigned aoff = mode == 2 ? (unsigned)((i * ks_tiles + ks / 4) * a_bytes + (ks % 4) * 32) : (unsigned)((i * ks_tiles + ks) * a_bytes);
                    const unsigned boff = mode == 2 ? (unsigned)((j * ks_tiles + ks / 4) * b_bytes + (ks % 4) * 32) : (unsigned)((j * ks_tiles + ks) * b_bytes);
                    umma_i8(tbase + g * N, make_desc(smem_u32(sA) + aoff, mode), make_desc(smem_u32(sB) + boff, mode), idesc, acc);
                    acc = 1;
                }
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int g = 0; g < S; ++g)
        for (int c0 = 0; c0 < N; c0 += 16) {
            int r[16];
            tmem_ld16(tbase + ((unsigned)(warp * 32) << 16) + g * N + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) out[((size_t)g * 128 + tid) * N + c0 + c] = r[c];
        }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// ---- test 2: rates ----------------------------------------------------------------------------
// what & 1: issue 28 MMAs (M=128, N, K=32; 7 x 7 slice pairs, 7 accumulators) per stage
// what & 2: bring each stage with two bulk copies (28 KB + N*32*7 bytes) from `src` (footprint `span` bytes)
// One CTA per SM; thread 0 = copy producer, thread 32 = MMA issuer; `stages`-deep ring.
template <int N>
__global__ void __launch_bounds__(128, 1) probe_rate(const uint8_t* __restrict__ src, size_t span, int iters, int what, int nstages,
                                                     long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long full[8], empty[8], done;
    __shared__ unsigned tslot;
    constexpr int A_BYTES = 7 * 4096, B_BYTES = 7 * N * 32, STAGE = A_BYTES + B_BYTES;
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
        // every CTA walks the footprint from its own offset, stage by stage
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
        for (int it = 0; it < iters; ++it) {
            const int s = it % nstages;
            if (do_copy) { mbar_wait(full + s, (it / nstages) & 1); tc_fence_after(); }
            if (do_mma) {
                const unsigned sa = smem_u32(smem + (size_t)s * STAGE), sb = sa + A_BYTES;
#pragma unroll 1
                for (int g = 0; g < 7; ++g)
                    for (int i = 0; i <= g; ++i)
                        umma_i8(tbase + ((g * N) & 511), make_desc(sa + i * 4096, 0), make_desc(sb + (g - i) * N * 32, 0), idesc, 1);
                umma_commit(empty + s);
            } else {
                asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(smem_u32(empty + s)) : "memory");
            }
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

static int run_correct(int mode, int S, int ksteps, int N) {
    const int a_bytes = mode == 2 ? 128 * 128 : 4096, b_bytes = mode == 2 ? N * 128 : N * 32;
    const int ks_tiles = mode == 2 ? (ksteps + 3) / 4 : ksteps;
    std::vector<int8_t> hA((size_t)S * ks_tiles * a_bytes, 0), hB((size_t)S * ks_tiles * b_bytes, 0);
    std::vector<int8_t> la((size_t)S * 128 * ksteps * 32), lb((size_t)S * N * ksteps * 32);
    srand(1234 + mode * 7 + N);
    for (auto& v : la) v = (int8_t)(rand() % 256 - 128);
    for (auto& v : lb) v = (int8_t)(rand() % 256 - 128);
    for (int s = 0; s < S; ++s)
        for (int ks = 0; ks < ksteps; ++ks) {
            for (int r = 0; r < 128; ++r)
                for (int k = 0; k < 32; ++k) {
                    const size_t t = mode == 2 ? (size_t)(s * ks_tiles + ks / 4) * a_bytes + tile_off(2, r, (ks % 4) * 32 + k)
                                               : (size_t)(s * ks_tiles + ks) * a_bytes + tile_off(mode, r, k);
                    hA[t] = la[(((size_t)s * 128 + r) * ksteps + ks) * 32 + k];
                }
            for (int r = 0; r < N; ++r)
                for (int k = 0; k < 32; ++k) {
                    const size_t t = mode == 2 ? (size_t)(s * ks_tiles + ks / 4) * b_bytes + tile_off(2, r, (ks % 4) * 32 + k)
                                               : (size_t)(s * ks_tiles + ks) * b_bytes + tile_off(mode, r, k);
                    hB[t] = lb[(((size_t)s * N + r) * ksteps + ks) * 32 + k];
                }
        }
    int8_t *dA, *dB;
    int* dO;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dO, (size_t)S * 128 * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0xff, (size_t)S * 128 * N * 4));
    const size_t smem = hA.size() + hB.size() + 1024;
    if (N == 64) {
        CK(cudaFuncSetAttribute(probe_correct<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe_correct<64><<<1, 128, smem>>>(dA, dB, S, ksteps, mode, dO);
    } else {
        CK(cudaFuncSetAttribute(probe_correct<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe_correct<128><<<1, 128, smem>>>(dA, dB, S, ksteps, mode, dO);
    }
    CK(cudaDeviceSynchronize());
    std::vector<int> hO((size_t)S * 128 * N);
    CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int g = 0; g < S; ++g)
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                long ref = 0;
                for (int i = 0; i <= g; ++i)
                    for (int kk = 0; kk < ksteps * 32; ++kk)
                        ref += (long)la[((size_t)i * 128 + m) * ksteps * 32 + kk] * (long)lb[((size_t)(g - i) * N + n) * ksteps * 32 + kk];
                if ((long)hO[((size_t)g * 128 + m) * N + n] != ref) {
                    if (bad < 3) printf("    mismatch g=%d m=%d n=%d got %d want %ld\n", g, m, n, hO[((size_t)g * 128 + m) * N + n], ref);
                    ++bad;
                }
            }
    printf("correct mode=%d S=%d ksteps=%d N=%d : %s (%ld of %d wrong)\n", mode, S, ksteps, N, bad ? "FAIL" : "ok", bad, S * 128 * N);
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
    return bad != 0;
}

template <int N>
static void run_rate(int what, size_t span_mb, int nstages, int iters) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t span = span_mb << 20;
    uint8_t* src;
    long long* cyc;
    CK(cudaMalloc(&src, span));
    CK(cudaMemset(src, 1, span));
    CK(cudaMalloc(&cyc, sms * sizeof(long long)));
    constexpr int STAGE = 7 * 4096 + 7 * N * 32;
    const size_t smem = (size_t)nstages * STAGE + 1024;
    CK(cudaFuncSetAttribute(probe_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe_rate<N><<<sms, 128, smem>>>(src, span, iters / 4, what, nstages, cyc);      // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe_rate<N><<<sms, 128, smem>>>(src, span, iters, what, nstages, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> h(sms);
    CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    const double bytes = (double)sms * iters * STAGE, macs = (double)sms * iters * 28.0 * 128 * N * 32;
    printf("rate N=%d what=%d (mma=%d copy=%d) span=%zu MB stages=%d iters=%d : %.3f ms, %.0f clk/stage", N, what, what & 1, (what >> 1) & 1,
           span_mb, nstages, iters, ms, (double)mx / iters);
    if (what & 2) printf(", %.2f TB/s = %.1f B/clk/SM", bytes / ms / 1e9, (double)STAGE * iters / (double)mx);
    if (what & 1) printf(", %.1f int8 TOP/s (%.1f clk/MMA)", 2.0 * macs / ms / 1e9, (double)mx / iters / 28.0);
    printf("\n");
    cudaFree(src); cudaFree(cyc);
}

int main(int argc, char** argv) {
    int fails = 0;
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    if (only < 0 || only == 0) {
        for (int mode = 0; mode < 3; ++mode) {
            fails += run_correct(mode, 1, 1, 64);
            fails += run_correct(mode, 1, 4, 64);
            fails += run_correct(mode, 3, 5, 64);
            fails += run_correct(mode, 2, 4, 128);
        }
        fails += run_correct(0, 7, 3, 64);
    }
    if (only < 0 || only == 1) {
        run_rate<64>(1, 64, 4, 4000);          // MMA issue rate alone
        run_rate<128>(1, 64, 2, 2000);
        run_rate<64>(2, 16, 4, 4000);          // bulk copies alone, L2-resident footprint
        run_rate<64>(2, 64, 4, 4000);
        run_rate<64>(2, 120, 4, 4000);
        run_rate<64>(2, 1024, 4, 4000);        // DRAM
        run_rate<64>(3, 64, 4, 4000);          // both
        run_rate<64>(3, 64, 5, 4000);
        run_rate<64>(3, 120, 5, 4000);
        run_rate<128>(3, 64, 3, 2000);
    }
    printf("probe done, %d correctness failures\n", fails);
    return 0;
}
