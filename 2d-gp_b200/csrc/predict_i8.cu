// Fused predictive pass on the int8 tensor cores (tcgen05.mma kind::i8, accumulators in TMEM): the product
// V = Z K* of predict.cu (Z = L^-1) computed from balanced base-256 digit slices of both operands
// (Ozaki splitting): Z ~ zu_r sum_i A_i 256^(S-1-i), K* ~ ku sum_j B_j 256^(S-1-j), and
//     V ~ zu_r ku sum_{d < S} 256^(2S-2-d) sum_{i+j=d} A_i B_j^T
// with every A_i B_j^T an exact int8 x int8 -> int32 product.  The S (S + 1) / 2 slice products of a k-step
// are S (S + 1) / 2 tcgen05.mma instructions into S accumulators (one per d) of NC columns each; the fp64
// recombination (Horner in 256), the squares and the column sums happen once per row block, out of TMEM.
// Integer arithmetic is exact, so the result is deterministic and independent of how the grid is cut.
// What is lost is the tail i + j >= S: with S = 6 (S = 7) the variance agrees with the fp64 path to
// ~2e-10 (~1e-12) relative at the conditioning of the BASELINE configurations (tools/ozaki_emulate.py);
// the slice count is chosen at fit time from the conditioning bound and the fp64 DMMA kernel of predict.cu
// remains the path beyond it (gate, below).  Same role as predict.cu: replaces GPy model.predict
// (krig.py:543-544), GP_scripts.getMean + the diagonal of GP_laser.py:129-131.
//
// Why it is faster: the FP64 pipe peaks at 37 TFLOP/s; kind::i8 at M = 128, N = 80 takes ~55 clocks per
// 128 x 80 x 32 product per SM (an SS-mode MMA re-reads both operand tiles from shared memory; whatever N, an
// instruction costs >= 47 clocks: tools/umma_probe4.cu), i.e. 21 instructions per k-step of the fp64-equivalent
// product at S = 6.  Two further facts shape the kernel (same probe): the tensor pipe takes one MMA at a time from a
// warp and queues nothing behind it, so every instruction the issuing warp spends on anything else is pipe idle time
// (hence two issuing warps taking the stages in turn); and a digit slice of a tile that is identically zero --
// covariances of distant points, entries of L^-1 far from the diagonal: a third of the slice products at
// configs[1], 60 % at N = 8192 -- need neither be copied nor multiplied, exactly (tools/sparsity_emulate.py).
//
// CTA (one per SM, persistent over column tiles of NC / 2 grid points x 2 components), by warp (a warp's scheduler
// and TMEM lane quarter are warp % 4):
//   0-7         epilogue: (A) TMEM -> one 64-bit integer per entry, accumulator by accumulator, each handed back to
//               the issuers as soon as it is read; (B) fp64 scale, square, sum over the 128 rows, running column
//               sums, while the tensor core is on the next row block.  Warp w: TMEM lanes 32 (w % 4).., column half w / 4
//   8, 12       MMA issuers: stage n belongs to issuer n % 2; straight-line code per (a, b)
//   15 (14)     producer / scheduler: reads one byte per k-step tile of Z (written at fit time) and of the panel
//               (written by the generators): the leading all-zero slices a, b; drops the k-step when a + b >= S, else
//               copies exactly the slices that take part in a product (A: a .. S-1-b, B: b .. S-1-a) with two bulk
//               copies and writes the stage header (a, b, segment flags) for the issuers; lane-parallel
//   9-11, 13, 14  generators: build the digit slices of the K* panel of the NEXT column tile into the other of two
//               global scratch panels while the tensor core works on the current one, the mean K*^T alpha in fp64 on
//               the way, and the OR of the digits per k-step tile (-> b)
// The kernel runs against the board's power cap (back to back it settles at ~1600 of 1965 MHz): bytes not moved
// come back as clock, which is why only the slices that are used are copied and why the first panel k-steps keep L2
// priority.  Times of bring-up experiments are therefore taken in SM clocks AND in milliseconds (tools/i8_ab.py).
// tests/tools/i8_protocol_sim.py is a discrete-event model of the barrier protocol between these roles.
#include "common.cuh"
#include "linalg.h"
#include "umma.cuh"

namespace gp2d {

constexpr int I8_SMAX = 7;                           // digits stored per entry of Zq
constexpr int I8_GSLOTS = 4;                         // k-steps a CTA generates side by side (x NC / 2 grid points = generator threads)
// An int32 accumulator d sums (d + 1) digit products over k: at most k (2 64 128 + (d - 1) 128 128) for d = S - 1
// (top digits are within [-64, 64]), which passes 2^31 beyond k = 21845 (S = 7).  Row blocks that reach further are
// accumulated in segments of I8_KSEG k-steps (16384 rows: 1.6e9 at worst); the segments add up exactly in the
// 64-bit integers of the epilogue.
constexpr int I8_KSEG = 512;
constexpr int I8_MAX_NPAD = 65536;
constexpr size_t I8_PACE_BYTES = 4096;               // head of the scratch: one arrival counter per row block (<= 512)
constexpr int I8_PACE_WINDOW = 2;                    // row blocks a CTA may be ahead of the slowest one
constexpr long long I8_PACE_LIMIT = 400000;          // clocks a CTA waits for the others at most (pacing is best effort)


// Bring-up watchdog (-DGP2D_I8_WATCHDOG): a wait that does not complete within ~2 s records who was waiting on what
// and releases every later wait of the launch, so a protocol error ends in a report instead of a hung device.
#ifdef GP2D_I8_WATCHDOG
__device__ int g_i8_abort;
__device__ int g_i8_abort_cta = -1;
__device__ unsigned long long g_i8_wd[4];
__device__ unsigned long long g_i8_state[32];        // per warp of the CTA that timed out first: what it was waiting for
__device__ __forceinline__ void i8_wait(unsigned long long* bar, unsigned parity, unsigned tag, unsigned info) {
    const long long t0 = clock64();
    for (unsigned n = 0;; ++n) {
        unsigned ok;
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if ((n & 255u) == 255u) {
            bool out = *(volatile int*)&g_i8_abort != 0;
            if (!out && clock64() - t0 > 4000000000ll) {
                if (atomicCAS(&g_i8_abort_cta, -1, (int)blockIdx.x) == -1) {
                    g_i8_wd[0] = ((unsigned long long)tag << 32) | info;
                    g_i8_wd[1] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;
                    g_i8_wd[2] = parity;
                }
                __threadfence();
                atomicExch(&g_i8_abort, 1);
                out = true;
            }
            if (out) {
                if (*(volatile int*)&g_i8_abort_cta == (int)blockIdx.x)
                    g_i8_state[threadIdx.x >> 5] = ((unsigned long long)parity << 63) | ((unsigned long long)tag << 32) | info;
                return;
            }
        }
    }
}
#define I8_WAIT(bar, parity, tag, info) i8_wait(bar, parity, tag, (unsigned)(info))
#elif defined(GP2D_I8_WAITPROF)
// Bring-up profile (-DGP2D_I8_WAITPROF): SM clocks spent in each kind of wait (tag), summed over the warps' first lanes
__device__ unsigned long long g_i8_waitclk[16];
__device__ __forceinline__ void i8_wait_prof(unsigned long long* bar, unsigned parity, unsigned tag) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    if ((threadIdx.x & 31) == 0) atomicAdd(&g_i8_waitclk[tag & 15], (unsigned long long)(clock64() - t0));
}
extern "C" int gp2d_dbg_i8_waitprof(unsigned long long* out16) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_i8_waitclk, 16 * sizeof(unsigned long long));
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(g_i8_waitclk, z, sizeof(z));
    return 0;
}
#define I8_WAIT(bar, parity, tag, info) i8_wait_prof(bar, parity, tag)
#else
#define I8_WAIT(bar, parity, tag, info) mbar_wait(bar, parity)
#endif

// Stage header, written by the producer for the MMA issuers: bits 0-7 a, 8-15 b (leading all-zero slices of the A and
// of the B tile), then:
constexpr unsigned H_FIRST = 1u << 16;        // first stage of an accumulation segment: taken whole, overwrites the accumulators
constexpr unsigned H_LAST = 1u << 17;         // last stage of the segment
constexpr unsigned H_ITEM_LAST = 1u << 18;    // ... and of the item (column tile)
constexpr unsigned H_FINAL = 1u << 19;        // ... and of the CTA: this issuer is done
constexpr unsigned H_PENULT = 1u << 20;       // the stage before the last of the segment (the other issuer's last one)
constexpr unsigned H_ITEM_PENULT = 1u << 21;  // ... of the last segment of the item
constexpr unsigned H_SECOND = 1u << 22;       // the stage after the first: may not overtake it
constexpr unsigned H_SEGNZ = 1u << 24;        // not the first segment of the CTA: the epilogue has to hand the accumulators back
constexpr unsigned H_SEGPAR = 1u << 25;       // parity of the segment count
constexpr unsigned H_ITEMPAR = 1u << 26;      // parity of the item count (which panel)
constexpr unsigned H_EXIT = 1u << 27;         // no work: the issuer that did not get the final stage leaves

// Work counters of the launches so far (statistics for bench.py, read and cleared by gp2d_dbg_i8_counters; one atomic
// per CTA and launch): slice products issued (MMAs), stages issued, k-steps visited (= stages of the dense schedule).
__device__ unsigned long long g_i8_count[8];         // [6]: clocks the generators spent building panels (first generator thread), summed over CTAs         // + SM clocks from the first to the last stage, summed over the CTAs; CTAs

// bulk copy with an L2 eviction priority: Zq tiles are re-read by every CTA (keep), panel tiles are read once (stream)
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar,
                                              unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
                     smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_keep() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_stream() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}

template <int S, int NC>
struct I8Cfg {
    static constexpr int NG = NC / 2;                 // grid points per column tile
    static constexpr int GSLOTS = I8_GSLOTS;          // k-steps generated side by side
    static constexpr int OBS_BATCH = 16 * GSLOTS;     // observations staged in shared memory at a time
    static constexpr int GT = NG * GSLOTS;            // generator threads: (grid point, k-step slot)
    static constexpr int EPI_WARPS = 8;               // two per TMEM lane quarter: column halves
    static constexpr int MMA_WARPS = 2;               // the tensor pipe idles whenever its issuing warp does anything else: two take turns
    static constexpr int THREADS = 32 * (EPI_WARPS + 1 + MMA_WARPS) + GT;
    // Roles by warp.  A warp's scheduler (and its TMEM lane quarter) is warp % 4.  The issuers are latency critical
    // (the tensor pipe waits for them), so both sit on scheduler 0, which they share with two epilogue warps only; the
    // generators, which run long fp64 sequences, and the producer take the warps of the other three schedulers.
    static constexpr int ISSUER0 = EPI_WARPS, ISSUER1 = EPI_WARPS + 4;
    static constexpr int GEN_WARPS = GT / 32;
    static constexpr int PRODUCER = (GEN_WARPS == 5) ? 15 : 14;
    static_assert(GEN_WARPS == 4 || GEN_WARPS == 5, "role table");
    // generator warp index of a warp (9, 10, 11, 13, 14 -> 0 .. 4), -1 for the others
    __host__ __device__ static constexpr int gen_index(int w) {
        return (w == 9) ? 0 : (w == 10) ? 1 : (w == 11) ? 2 : (w == 13) ? 3 : (w == 14 && GEN_WARPS == 5) ? 4 : -1;
    }
    static constexpr int HC = NC / 2;                 // accumulator columns per epilogue warp
    static constexpr int BTILE = NC * I8_KSTEP;       // bytes of one B slice tile
    static constexpr int STAGE_BYTES = S * (I8_ATILE_BYTES + BTILE);
    static constexpr int STAGES = S == 6 ? 5 : 4;
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    // A stage's "full" barrier is waited on by the issuer n % MMA_WARPS; with an odd ring each issuer would meet a slot at
    // every other use only, and one parity bit cannot tell "my last use" from "the use in between has not landed yet"
    // (copies of different stages complete out of order).  So the full barriers are indexed by n % NFULL, NFULL a
    // multiple of both the ring and the issuer count: every barrier then belongs to one issuer, use after use.
    static constexpr int NFULL = MMA_WARPS * STAGES;
    static constexpr int NBARS = NFULL + STAGES + 2 + 4 + 2 + 8;   // full, empty, acc_full, (unused), panel_full[2], panel_empty[2], first_done, pad, acc_empty[8]
    // doubles after the barriers: observation stage, mean partials [GSLOTS][NG][2], column sums [4][NC], parameters,
    // digit ORs of the generators [2][GSLOTS], stage headers [STAGES] (32-bit)
    static constexpr int TAIL_DOUBLES = 6 * OBS_BATCH + GSLOTS * NG * 2 + 4 * NC + 64 + 16 + 8 + 64;      // ... and a copy of the exp table
    static_assert(STAGES <= 16 && GSLOTS <= 8, "stage headers, digit ORs");
    static constexpr int BARS_BYTES = (NBARS * 8 + 16 + 15) / 16 * 16;     // barriers, TMEM slot; the observation stage behind it is read 16 bytes at a time
    static constexpr int SMEM_BYTES = RING_BYTES + BARS_BYTES + TAIL_DOUBLES * 8;
    static constexpr int TMEM_COLS = 512;
    static_assert(S * NC <= TMEM_COLS, "S accumulators of NC columns must fit TMEM");
    static_assert(NC % 16 == 0, "tcgen05.mma M = 128 needs N % 16 == 0");
    static_assert(HC % 8 == 0, "an epilogue warp reads its columns eight at a time");
    static_assert(THREADS % 32 == 0 && GT % 32 == 0, "whole warps");
    static_assert(SMEM_BYTES <= 232448, "shared memory");
};

struct PredictI8Args {
    const int8_t* Zq;             // [k-step of the lower triangle][I8_SMAX][4096]: digit slices of Z, tile images
    const uint8_t* zlead;         // [k-step of the lower triangle]: leading all-zero slices of that tile
    const double* zunit;          // [npad] value of one unit of digit I8_SMAX - 1 of row r
    int npad;
    const double* alpha;          // pair-interleaved, zero padded
    const double* X; int N;
    HelmParams hp;
    const double* Xs; int M;
    long out_stride;
    double kss, kss1, var_add;
    double kscale;                // K* entries are quantised as rint(k kscale), |k kscale| <= 2^(8 S - 2)
    double cscale;                // ku 256^(S-1) 256^(I8_SMAX-S): v = Horner zunit[r] cscale
    double* mean; double* var;
    uint8_t* scratch;             // per CTA: 2 panels
    size_t panel_bytes, cta_bytes, lead_off;     // a panel: digit slices, then at lead_off one byte per k-step (leading all-zero slices)
    int ntiles;
    int keep_ks;                  // panel k-steps below this are copied with L2 evict-last priority (see the producer)
    int* pace;                    // [npad / 128] row-block arrival counters of this launch (zeroed), or null: see the producer
    const int* gate;              // slice count chosen at fit time; the kernel runs only when it equals S
    int dbg;                      // bring-up knobs: 8 = copy and multiply the all-zero slices too (same results bit for bit);
                                  // timing experiments with wrong results: 1 all CTAs stream panel 0, 2 no epilogue math, 4 no generation, 16 no MMAs, 32 no copies; 64 no pacing, 128 pacing whatever the size
};

// ---------------------------------------------------------------------------------------------------
// fit side: digit slices of Z = L^-1
// ---------------------------------------------------------------------------------------------------
// one warp per row: power-of-two scale from max |Z[r][0 .. (r / 128 + 1) 128)| (upper part of the diagonal
// tile is exactly zero).  zunit[r] = 2^(e - 55), zqs[r] = 2^(55 - e) with max / 2^e in [1/4, 1/2).
__global__ void __launch_bounds__(256) i8_rowscale_kernel(const double* __restrict__ Z, long ldz, int npad,
                                                          double* __restrict__ zunit, double* __restrict__ zqs, long bstride) {
    Z += (long)blockIdx.y * bstride;
    zunit += (long)blockIdx.y * bstride;
    zqs += (long)blockIdx.y * bstride;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= npad) return;
    const int ncol = (row / TILE + 1) * TILE;
    const double* zr = Z + (long)row * ldz;
    double mx = 0.0;
    for (int c = 2 * lane; c < ncol; c += 64) {
        const double2 v = *reinterpret_cast<const double2*>(zr + c);
        mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
        int ex = 0;
        if (mx > 0.0 && isfinite(mx)) frexp(mx, &ex); else ex = 0;      // a non-finite factor is reported through info
        const int e = ex + 1;
        zunit[row] = scalbn(1.0, e - (8 * I8_SMAX - 1));
        zqs[row] = scalbn(1.0, (8 * I8_SMAX - 1) - e);
    }
}

// Leading all-zero digits of a set of biased values: y = OR over the set of (q + BIAS) ^ BIAS has byte p zero exactly
// when digit p of every value is zero, so the count of zero top bytes of y (out of ND) is the number of leading
// (most significant) slices of the set that are identically zero.
__device__ __forceinline__ int i8_lead_zero_slices(unsigned long long y, int nd) {
    const int used = y ? (64 - __clzll((long long)y) + 7) / 8 : 0;
    return nd - used;
}

// one CTA per lower 128 x 128 tile (I, J): 4 k-steps x 7 slices of 4 KB tile images.  Thread (r, h): row r,
// k-steps 2h, 2h+1.  Also writes, per k-step tile, how many of its leading slices are all zero (entries far from
// the diagonal are small against the row unit): the predictive kernel neither copies nor multiplies those.
__global__ void __launch_bounds__(256) i8_quantize_kernel(const double* __restrict__ Z, long ldz, const double* __restrict__ zqs,
                                                          int8_t* __restrict__ Zq, uint8_t* __restrict__ lead, long bstride) {
    __shared__ unsigned long long s_or[4];
    Z += (long)blockIdx.y * bstride;
    zqs += (long)blockIdx.y * bstride;
    Zq += (long)blockIdx.y * bstride * (long)sizeof(double);
    lead += (long)blockIdx.y * bstride * (long)sizeof(double);
    if (threadIdx.x < 4) s_or[threadIdx.x] = 0ull;
    __syncthreads();
    const int t = blockIdx.x;
    int I = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(I + 1) * (I + 2) / 2 <= t) ++I;
    while ((long)I * (I + 1) / 2 > t) --I;
    const int J = t - I * (I + 1) / 2;
    const int r = threadIdx.x & 127, h = threadIdx.x >> 7;
    const double* src = Z + ((long)I * TILE + r) * ldz + (long)J * TILE + h * 64;
    const double qs = zqs[I * TILE + r];
    constexpr long long BIAS = i8_digit_bias(I8_SMAX);
    const size_t ks0 = (size_t)2 * I * (I + 1) + 4 * (size_t)J + 2 * h;
    unsigned long long yor[2] = {0ull, 0ull};
#pragma unroll 1
    for (int c16 = 0; c16 < 4; ++c16) {          // 16 consecutive k: one 16-byte row chunk of every slice image
        long long q[16];
        unsigned long long y = 0ull;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const double2 v = *reinterpret_cast<const double2*>(src + c16 * 16 + j);
            q[j] = __double2ll_rn(v.x * qs) + BIAS;
            q[j + 1] = __double2ll_rn(v.y * qs) + BIAS;
            y |= (unsigned long long)(q[j] ^ BIAS) | (unsigned long long)(q[j + 1] ^ BIAS);
        }
        yor[c16 >> 1] |= y;
        int8_t* dst = Zq + (ks0 + (c16 >> 1)) * (size_t)(I8_SMAX * I8_ATILE_BYTES) + i8_tile_off(r, (c16 & 1) * 16);
#define GP2D_I8_STORE(P)                                                                                          \
        *reinterpret_cast<uint4*>(dst + (I8_SMAX - 1 - P) * I8_ATILE_BYTES) =                                     \
            make_uint4(i8_digit_word<P>(q[0], q[1], q[2], q[3]), i8_digit_word<P>(q[4], q[5], q[6], q[7]),        \
                       i8_digit_word<P>(q[8], q[9], q[10], q[11]), i8_digit_word<P>(q[12], q[13], q[14], q[15]));
        GP2D_I8_STORE(0) GP2D_I8_STORE(1) GP2D_I8_STORE(2) GP2D_I8_STORE(3) GP2D_I8_STORE(4) GP2D_I8_STORE(5) GP2D_I8_STORE(6)
#undef GP2D_I8_STORE
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        unsigned long long y = yor[e];
#pragma unroll
        for (int o = 16; o; o >>= 1) y |= __shfl_xor_sync(0xffffffffu, y, o);
        if ((threadIdx.x & 31) == 0 && y) atomicOr(&s_or[2 * h + e], y);
    }
    __syncthreads();
    if (threadIdx.x < 4)
        lead[(size_t)2 * I * (I + 1) + 4 * (size_t)J + threadIdx.x] = (uint8_t)i8_lead_zero_slices(s_or[threadIdx.x], I8_SMAX);
}

__global__ void i8_set_gate_kernel(int* gate, int s) { *gate = s; }

// digit slices of the lower triangle, then one byte per k-step tile (leading all-zero slices)
static size_t i8_zq_slice_bytes(int npad) {
    const size_t nb = (size_t)npad / TILE;
    return 2 * nb * (nb + 1) * (size_t)(I8_SMAX * I8_ATILE_BYTES);
}
size_t i8_zq_bytes(int npad) {
    const size_t nb = (size_t)npad / TILE;
    return i8_zq_slice_bytes(npad) + ((2 * nb * (nb + 1) + 255) / 256) * 256;
}
int i8_max_npad() { return I8_MAX_NPAD; }

// zqs: npad doubles of scratch; batch > 1: problem b at every pointer + b bstride doubles
cudaError_t i8_quantize_lower(const double* Z, long ldz, int npad, double* zunit, double* zqs, int8_t* Zq, cudaStream_t st,
                              int batch, long bstride) {
    const int nb = npad / TILE;
    const long bs = batch > 1 ? bstride : 0;
    i8_rowscale_kernel<<<dim3((npad + 7) / 8, batch), 256, 0, st>>>(Z, ldz, npad, zunit, zqs, bs);
    i8_quantize_kernel<<<dim3(nb * (nb + 1) / 2, batch), 256, 0, st>>>(Z, ldz, zqs, Zq, reinterpret_cast<uint8_t*>(Zq) + i8_zq_slice_bytes(npad), bs);
    return cudaGetLastError();
}
cudaError_t i8_set_gate(int* gate, int s, cudaStream_t st) {
    i8_set_gate_kernel<<<1, 1, 0, st>>>(gate, s);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// predict side
// ---------------------------------------------------------------------------------------------------
// exp(x), x <= 0, without a table: the generators run beside a tensor core that keeps the shared-memory / L1 data pipe
// 70 % busy, where every load of theirs queues (they take four times as long as on an idle SM and are then what the
// column tile waits for: tools/i8_knobs.py, "generators busy"); fp64 arithmetic is what the SM has to spare.
// n = rint(4 x / ln 2), r = x - n ln2 / 4 in two pieces (|r| <= 0.0867), Taylor to r^9 (6e-18), 2^((n & 3) / 4) from
// three constants, 2^(n >> 2) into the exponent: 1 ulp, like helmholtz.cuh's table version.
__device__ __forceinline__ double exp_neg_sel(double x) {
    x = fmax(x, -708.0);
    const double nd = rint(x * 0x1.71547652b82fep+2);
    const int ni = (int)nd;
    double r = fma(-nd, 0x1.62e42fefa0000p-3, x);
    r = fma(-nd, 0x1.cf79abc9e3b3ap-42, r);
    const double r2 = r * r;
    // Estrin on 1 + r + r^2/2 + ... + r^9/9!
    const double c01 = 1.0 + r, c23 = fma(r, 1.0 / 6.0, 0.5), c45 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    const double c67 = fma(r, 1.0 / 5040.0, 1.0 / 720.0), c89 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0);
    const double r4 = r2 * r2;
    const double lo = fma(r2, c23, c01), mid = fma(r2, c67, c45);
    const double pol = fma(r4, fma(r4, c89, mid), lo);
    const int j = ni & 3;
    const double c = j == 0 ? 1.0 : j == 1 ? 0x1.306fe0a31b715p+0 : j == 2 ? 0x1.6a09e667f3bcdp+0 : 0x1.ae89f995ad3adp+0;
    return pol * __longlong_as_double(__double_as_longlong(c) + ((long long)(ni >> 2) << 52));
}

// Generator: thread (g, slot) owns grid point gp0 + g and the k-steps slot, slot + 8, ... (16 observations =
// 32 rows of K* each).  Per k-step it forms 16 2x2 blocks, quantises the 48 distinct values and stores, for
// each digit and each of its two columns, the two 16-byte row chunks of the slice image.
template <int S, bool SAME_LEN, bool HAS_T>
__device__ __forceinline__ void i8_generate_kstep(const HelmParams& hp, const HelmPoint& gpt, double wg, int nvalid, int o0,
                                                  const double* __restrict__ stage, double kscale, uint8_t* __restrict__ dst,
                                                  int btile, double& m0, double& m1, unsigned long long& yor,
                                                  const double* __restrict__ tab) {
    constexpr long long BIAS = i8_digit_bias(S);
    const double wdk = hp.w_df * wg * kscale, wck = hp.w_cf * wg * kscale;      // wg is 0 or 1, kscale a power of two: exact
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {                 // 8 observations = 16 consecutive k
        unsigned wa[S][4], wb[S][4];               // column 2g / 2g+1, digit p, word w
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            long long qa[2], qb[2], qc[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int ol = h * 8 + w * 2 + e;
                // (a, b), (t, -), (alpha_u, alpha_v): three 16-byte loads per observation
                const double2 s01 = *reinterpret_cast<const double2*>(stage + ol * 6);
                const double2 s23 = *reinterpret_cast<const double2*>(stage + ol * 6 + 2);
                const double2 s45 = *reinterpret_cast<const double2*>(stage + ol * 6 + 4);
                const double d1 = s01.x - gpt.a, d2 = s01.y - gpt.b;
                const double a = d1 * d1, b = d2 * d2, c = d1 * d2, r2 = a + b;
                const double E = exp_neg(-0.5 * hp.s_df * r2, tab);
                const double F = SAME_LEN ? E : exp_neg(-0.5 * hp.s_cf * r2, tab);
                // The quantisation scale (a power of two) and the validity of the pair ride on the two kernel weights:
                // fp64 instructions are what this kernel cannot afford (they only issue in the gaps the tensor pipe
                // leaves and push its work back by more than their own time: tools/i8_knobs.py), and the block below
                // then yields k kscale directly -- the same bits as scaling afterwards, zero for a padded pair.
                double wd = (o0 + ol) < nvalid ? wdk : 0.0, wc = (o0 + ol) < nvalid ? wck : 0.0;
                if (HAS_T) {
                    const double dt = s23.x - gpt.t;
                    const double tf = hp.tvar * exp_neg(-hp.thalf * dt * dt, tab);
                    wd *= tf; wc *= tf;
                }
                const double ee = wd * E, ff = wc * F;
                const double es = ee * hp.s_df, fs = ff * hp.s_cf;
                const double k11 = fma(-b, es, ee) + fma(-a, fs, ff);
                const double k22 = fma(-a, es, ee) + fma(-b, fs, ff);
                const double k12 = c * (es - fs);
                m0 = fma(k11, s45.x, fma(k12, s45.y, m0));           // scaled by kscale: undone once per grid point
                m1 = fma(k12, s45.x, fma(k22, s45.y, m1));
                qa[e] = __double2ll_rn(k11) + BIAS;
                qb[e] = __double2ll_rn(k12) + BIAS;
                qc[e] = __double2ll_rn(k22) + BIAS;
                yor |= (unsigned long long)((qa[e] ^ BIAS) | (qb[e] ^ BIAS) | (qc[e] ^ BIAS));
            }
            // rows k = 2o, 2o+1, 2o+2, 2o+3: column 2g holds (k11, k12) per observation, column 2g+1 (k12, k22)
#define GP2D_I8_WORDS(P)                                                              \
            if (P < S) {                                                              \
                wa[P < S ? P : 0][w] = i8_digit_word<P>(qa[0], qb[0], qa[1], qb[1]);  \
                wb[P < S ? P : 0][w] = i8_digit_word<P>(qb[0], qc[0], qb[1], qc[1]);  \
            }
            GP2D_I8_WORDS(0) GP2D_I8_WORDS(1) GP2D_I8_WORDS(2) GP2D_I8_WORDS(3) GP2D_I8_WORDS(4) GP2D_I8_WORDS(5) GP2D_I8_WORDS(6)
#undef GP2D_I8_WORDS
        }
#pragma unroll
        for (int p = 0; p < S; ++p) {
            // slice index S - 1 - p (most significant first); columns 2g, 2g+1 are adjacent 16-byte rows of the image
            uint4* d = reinterpret_cast<uint4*>(dst + (S - 1 - p) * btile + h * 128);
            __stcs(d, make_uint4(wa[p][0], wa[p][1], wa[p][2], wa[p][3]));        // read once, by a bulk copy: do not keep in L2
            __stcs(d + 1, make_uint4(wb[p][0], wb[p][1], wb[p][2], wb[p][3]));
        }
    }
}

// Leading all-zero slices of the k-step tile generated in batch b by the threads of one slot: the slot's first thread
// turns the OR of the batch into the panel's byte for that k-step and clears it for batch b + 2.
template <int S>
__device__ __forceinline__ void i8_flush_lead(unsigned long long* sh_or, uint8_t* plead, int b, int slot, int nks) {
    unsigned long long* o = sh_or + (b & 1) * I8_GSLOTS + slot;
    const int ks = b * I8_GSLOTS + slot;
    if (ks < nks) plead[ks] = (uint8_t)i8_lead_zero_slices(*o, S);
    *o = 0ull;
}

template <int S, int NC, bool SAME_LEN, bool HAS_T>
__device__ __forceinline__ void i8_generate_item(const PredictI8Args& p, const HelmParams& hp, uint8_t* __restrict__ panel,
                                                 double* __restrict__ stage, unsigned long long* sh_or, const double* __restrict__ tab,
                                                 int gp0, int gtid, double& mu0, double& mu1) {
    using C = I8Cfg<S, NC>;
    const int g = gtid % C::NG, slot = gtid / C::NG;
    const int gj = gp0 + g;
    const double wg = gj < p.M ? 1.0 : 0.0;
    const HelmPoint gpt = helm_point(hp, p.Xs, gj < p.M ? gj : 0);
    const int nobs_pad = p.npad >> 1;
    // the two image rows (columns 2g, 2g+1 of the tile) of this thread inside a slice tile
    const int rowoff = ((2 * g) >> 3) * 256 + ((2 * g) & 7) * 16;
    double m0 = 0.0, m1 = 0.0;
    uint8_t* plead = panel + p.lead_off;
    const int nks_all = p.npad / I8_KSTEP;
    int batch = 0;
    for (int ob = 0; ob < nobs_pad; ob += C::OBS_BATCH, ++batch) {
        named_barrier(2, C::GT);                  // the previous batch has been read, its digit ORs are complete
        if (batch > 0 && g == 0) i8_flush_lead<S>(sh_or, plead, batch - 1, slot, nks_all);
        __syncwarp();                             // bar.sync below is warp-aligned: reconverge after the divergent flush
        if (gtid < C::OBS_BATCH) {
            const int o = ob + gtid;
            const bool ov = o < p.N;
            const HelmPoint q = helm_point(hp, p.X, ov ? o : 0);
            stage[gtid * 6 + 0] = q.a;
            stage[gtid * 6 + 1] = q.b;
            stage[gtid * 6 + 2] = q.t;
            stage[gtid * 6 + 3] = 0.0;
            stage[gtid * 6 + 4] = ov ? __ldg(p.alpha + 2 * o) : 0.0;
            stage[gtid * 6 + 5] = ov ? __ldg(p.alpha + 2 * o + 1) : 0.0;
        }
        named_barrier(2, C::GT);
        const int o0 = ob + slot * 16;
        if (o0 < nobs_pad) {
            const int ks = o0 >> 4;
            unsigned long long yor = 0ull;
            i8_generate_kstep<S, SAME_LEN, HAS_T>(hp, gpt, wg, p.N, o0, stage + slot * 16 * 6, p.kscale,
                                                  panel + (size_t)ks * (S * C::BTILE) + rowoff, C::BTILE, m0, m1, yor, tab);
            if (yor) atomicOr(sh_or + (batch & 1) * I8_GSLOTS + slot, yor);
        }
    }
    named_barrier(2, C::GT);
    if (g == 0) i8_flush_lead<S>(sh_or, plead, batch - 1, slot, nks_all);
    __syncwarp();
    const double unscale = 1.0 / p.kscale;        // a power of two
    mu0 = m0 * unscale;
    mu1 = m1 * unscale;
}

// The MMAs of one k-step whose A tile has A and whose B tile has B leading all-zero slices: the slice pairs (i, j),
// i >= A, j >= B, i + j < S, accumulator i + j.  A and B are compile-time so that the issuing lane runs straight-line
// code (descriptor = stage base + constant): a single warp issues every MMA of the SM, and predicates per MMA
// (~270 instructions per k-step) cost as much time as the products themselves.  FIRST: the k-step that starts an
// accumulation segment (taken whole), whose first product per accumulator overwrites.
template <int S, int NC, int A, int B, bool FIRST>
__device__ __forceinline__ void i8_issue(unsigned tbase, unsigned a_lo, unsigned b_lo) {
    constexpr unsigned IDESC = i8_idesc(128, NC);
    constexpr int BT16 = (NC * I8_KSTEP) >> 4, AT16 = I8_ATILE_BYTES >> 4;
#pragma unroll
    for (int d = A + B; d < S; ++d)
#pragma unroll
        for (int i = A; i <= d - B; ++i)
            umma_i8_ss(tbase + d * NC, i8_desc(a_lo + (unsigned)(i * AT16)), i8_desc(b_lo + (unsigned)((d - i) * BT16)), IDESC,
                       (FIRST && i == A) ? 0u : 1u);
}
// the products of accumulator d of a first (whole) k-step: slice pairs (i, d - i), the first one overwrites
template <int S, int NC>
__device__ __forceinline__ void i8_issue_first(int d, unsigned tbase, unsigned a_lo, unsigned b_lo) {
    constexpr unsigned IDESC = i8_idesc(128, NC);
    constexpr int BT16 = (NC * I8_KSTEP) >> 4, AT16 = I8_ATILE_BYTES >> 4;
#pragma unroll
    for (int i = 0; i < S; ++i)
        if (i <= d)
            umma_i8_ss(tbase + d * NC, i8_desc(a_lo + (unsigned)(i * AT16)), i8_desc(b_lo + (unsigned)((d - i) * BT16)), IDESC, i == 0 ? 0u : 1u);
}

template <int S, int NC, int A>
__device__ __forceinline__ void i8_issue_b(int b, unsigned tbase, unsigned a_lo, unsigned b_lo) {
    switch (b) {
        case 0: if (A + 0 < S) i8_issue<S, NC, A, (A + 0 < S ? 0 : 0), false>(tbase, a_lo, b_lo); break;
        case 1: if (A + 1 < S) i8_issue<S, NC, A, (A + 1 < S ? 1 : 0), false>(tbase, a_lo, b_lo); break;
        case 2: if (A + 2 < S) i8_issue<S, NC, A, (A + 2 < S ? 2 : 0), false>(tbase, a_lo, b_lo); break;
        case 3: if (A + 3 < S) i8_issue<S, NC, A, (A + 3 < S ? 3 : 0), false>(tbase, a_lo, b_lo); break;
        case 4: if (A + 4 < S) i8_issue<S, NC, A, (A + 4 < S ? 4 : 0), false>(tbase, a_lo, b_lo); break;
        case 5: if (A + 5 < S) i8_issue<S, NC, A, (A + 5 < S ? 5 : 0), false>(tbase, a_lo, b_lo); break;
        case 6: if (A + 6 < S) i8_issue<S, NC, A, (A + 6 < S ? 6 : 0), false>(tbase, a_lo, b_lo); break;
        default: break;
    }
}
template <int S, int NC>
__device__ __forceinline__ void i8_issue_ab(int a, int b, unsigned tbase, unsigned a_lo, unsigned b_lo) {
    switch (a) {
        case 0: i8_issue_b<S, NC, 0>(b, tbase, a_lo, b_lo); break;
        case 1: i8_issue_b<S, NC, 1>(b, tbase, a_lo, b_lo); break;
        case 2: i8_issue_b<S, NC, 2>(b, tbase, a_lo, b_lo); break;
        case 3: i8_issue_b<S, NC, 3>(b, tbase, a_lo, b_lo); break;
        case 4: i8_issue_b<S, NC, 4>(b, tbase, a_lo, b_lo); break;
        case 5: i8_issue_b<S, NC, 5>(b, tbase, a_lo, b_lo); break;
        case 6: i8_issue_b<S, NC, (S > 6 ? 6 : 0)>(S > 6 ? b : S, tbase, a_lo, b_lo); break;
        default: break;
    }
}

template <int S, int NC>
__global__ void __launch_bounds__(I8Cfg<S, NC>::THREADS, 1) predict_i8_kernel(const __grid_constant__ PredictI8Args p) {
    using C = I8Cfg<S, NC>;
    if (*p.gate != S) return;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* ring = smem_raw;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_raw + C::RING_BYTES);
    unsigned long long* full = bars;
    unsigned long long* empty = bars + C::NFULL;
    unsigned long long* acc_full = empty + C::STAGES;
    unsigned long long* acc_empty = acc_full + 8;          // [S]: one per accumulator, handed back as soon as it is drained
    unsigned long long* panel_full = acc_full + 2;
    unsigned long long* panel_empty = acc_full + 4;
    unsigned long long* first_done = acc_full + 6;
    unsigned* tslot = reinterpret_cast<unsigned*>(bars + C::NBARS);
    double* sh_stage = reinterpret_cast<double*>(smem_raw + C::RING_BYTES + C::BARS_BYTES);
    double* sh_mu = sh_stage + 6 * C::OBS_BATCH;          // [GSLOTS][NG][2]
    double* sh_red = sh_mu + C::GSLOTS * C::NG * 2;       // [4][NC]
    double* sh_par = sh_red + 4 * NC;
    unsigned long long* sh_or = reinterpret_cast<unsigned long long*>(sh_par + 64);      // [2][8]
    volatile unsigned* sh_hdr = reinterpret_cast<volatile unsigned*>(sh_or + 16);          // [STAGES]
    double* sh_tab = reinterpret_cast<double*>(sh_or + 16) + 8;                             // exp_neg's table: the L1 left beside
                                                                                           // 208 KB of shared memory does not keep it

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = p.npad / TILE;
    uint8_t* panels = p.scratch + (size_t)blockIdx.x * p.cta_bytes;

    if (tid == 0) {
        for (int s = 0; s < C::NFULL; ++s) mbar_init(full + s, 1);
        for (int s = 0; s < C::STAGES; ++s) mbar_init(empty + s, 1);
        mbar_init(acc_full, C::MMA_WARPS);
        for (int d = 0; d < S; ++d) mbar_init(acc_empty + d, C::EPI_WARPS);
        mbar_init(first_done, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(panel_full + b, 1); mbar_init(panel_empty + b, C::MMA_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < (int)(sizeof(HelmParams) / sizeof(double))) sh_par[tid] = reinterpret_cast<const double*>(&p.hp)[tid];
    if (tid >= 64 && tid < 80) sh_or[tid - 64] = 0ull;
    if (tid >= 128 && tid < 192) sh_tab[tid - 128] = EXP2_TAB[tid - 128];
    if (warp == C::ISSUER0) tmem_alloc(tslot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = *tslot;

    if (warp < C::EPI_WARPS) {
        // ------------------------------ epilogue ------------------------------------------------
        // Warp w reads TMEM lanes 32 (w % 4) .. (rows of the row block) and the column half w / 4.  Two phases per row
        // block: (A) drain the accumulators, one after the other, into one 64-bit integer per entry -- digit sum d has the
        // weight 256^(S-1-d); the total keeps bits SH and up (the low digit sums are rounded in: far below the products
        // i + j >= S that the slicing leaves out) -- and hand each accumulator back to the MMA issuers as soon as it is
        // read; (B) convert, scale by the row unit, square and sum over the rows while the tensor core works on the
        // next row block.  Segments of a long row block add up exactly in the integers.
        constexpr int HC = C::HC, NG8 = HC / 8;
        constexpr int SH = S == 6 ? 8 : 14;               // S = 6: |total| < 2^60 over four segments; S = 7: < 2^62
        const int q = warp & 3, half = warp >> 2;
        const double cs = p.cscale * (double)(1 << SH);
        const unsigned tcol0 = tbase + ((unsigned)(q * 32) << 16) + (unsigned)(half * HC);
        unsigned ph = 0;
        for (int item = blockIdx.x; item < p.ntiles; item += gridDim.x) {
            double colsum[NG8];
#pragma unroll
            for (int g = 0; g < NG8; ++g) colsum[g] = 0.0;
            for (int rb = 0; rb < nb; ++rb) {
                const double rs = __ldg(p.zunit + rb * TILE + q * 32 + lane) * cs;
                const int nseg = (4 * (rb + 1) + I8_KSEG - 1) / I8_KSEG;
                long long t64[HC];
                for (int seg = 0; seg < nseg; ++seg) {
                    I8_WAIT(acc_full, ph, 1, (item << 8) | rb);
                    ph ^= 1u;
                    tc_fence_after();
                    // accumulator by accumulator (digit sum d has weight 256^(S-1-d); the total keeps bits SH and up, the
                    // low digits rounded in), each handed back to the issuers as soon as it is in the integers: the first
                    // stage of the next segment writes them in the same order
#pragma unroll
                    for (int d = 0; d < S; ++d) {
                        constexpr int NQ = HC / 8;        // two batches of loads: half the columns each (registers)
                        if (!(p.dbg & 2)) {
                            const int sh = 8 * (S - 1 - d) - SH;      // compile time after unrolling
#pragma unroll
                            for (int hb = 0; hb < 2; ++hb) {
                                int r[NQ][4];
#pragma unroll
                                for (int g = 0; g < NQ; ++g) tmem_ld4(tcol0 + d * NC + (hb * NQ + g) * 4, r[g]);
                                tmem_ld_wait();
#pragma unroll
                                for (int g = 0; g < NQ; ++g)
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const long long v = sh >= 0 ? (long long)r[g][j] * (1ll << (sh >= 0 ? sh : 0))
                                                                    : (((long long)r[g][j] + (1ll << (sh < 0 ? -sh - 1 : 0))) >> (sh < 0 ? -sh : 0));
                                        const int c = (hb * NQ + g) * 4 + j;
                                        t64[c] = (d == 0 && seg == 0) ? v : t64[c] + v;
                                    }
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty + d);
                    }
                }
                if (p.dbg & 2) continue;
#pragma unroll
                for (int g = 0; g < NG8; ++g) {
                    double t[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const double v = (double)t64[g * 8 + j] * rs; t[j] = v * v; }
                    // sum over the 32 rows of the warp, 8 columns at once: halve the columns with each exchange
#pragma unroll
                    for (int w = 4; w >= 1; w >>= 1) {
                        const bool up = (lane & (2 * w)) != 0;
#pragma unroll
                        for (int j = 0; j < w; ++j) {
                            const double keep = up ? t[j + w] : t[j], send = up ? t[j] : t[j + w];
                            t[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * w);
                        }
                    }
                    t[0] += __shfl_xor_sync(0xffffffffu, t[0], 1);
                    t[0] += __shfl_xor_sync(0xffffffffu, t[0], 16);
                    colsum[g] += t[0];                // column g 8 + ((lane >> 1) & 7)
                }
            }
            // columns of the item: sum of the four lane quarters in fixed order
            if (!(lane & 17)) {
#pragma unroll
                for (int g = 0; g < NG8; ++g) sh_red[q * NC + half * HC + g * 8 + (lane >> 1)] = colsum[g];
            }
            __syncwarp();                             // bar.sync is warp-aligned: reconverge after the divergent store
            named_barrier(1, 32 * C::EPI_WARPS);
            if (tid < NC) {
                const int pj = tid >> 1, cc = tid & 1;
                const int j = item * C::NG + pj;
                if (j < p.M) {
                    const double ss = ((sh_red[tid] + sh_red[NC + tid]) + sh_red[2 * NC + tid]) + sh_red[3 * NC + tid];
                    double v = (cc ? p.kss1 : p.kss) - ss;
                    v = v < 0.0 ? 0.0 : v;
                    p.var[(long)cc * p.out_stride + j] = v + p.var_add;
                }
            }
            __syncwarp();
            named_barrier(1, 32 * C::EPI_WARPS);
        }
    } else if (warp == C::PRODUCER) {
        // ------------------------------ producer / scheduler ----------------------------------------
        // Per accumulation segment (the k-steps of a row block, at most I8_KSEG): lane l of chunk c reads the
        // leading-zero-slice bytes of k-step 32 c + l of the Z row block and of the panel (all chunks' loads in flight
        // together).  A k-step whose non-zero slices cannot meet (a + b >= S) is dropped, the first one of the segment is
        // always taken whole (it initialises the accumulators).  Every live lane then emits ITS stage by itself, as
        // soon as the ring slot of that stage is free: header for the issuers (a, b, position in the segment), barrier,
        // two bulk copies of exactly the slices that take part in a product.  A stage costs this warp one pass of a
        // short divergent loop, not a round of shuffles and queue bookkeeping through lane 0 -- in the sparse regime
        // (10 products per stage) the serial version was what the issuers waited for.
        {
            int it = 0;
            long fills = 0, segs = 0;
            unsigned long long n_mma = 0, n_stage = 0, n_kstep = 0;       // per lane; summed at the end
            const long long clk0 = clock64();
            const bool noskip = (p.dbg & (8 | 4)) != 0;
            const unsigned long long pol_keep = l2_policy_keep(), pol_stream = l2_policy_stream();
            // Panel k-step ks is read once per row block rb >= ks / 4: the first k-steps of the panel are re-read by every
            // row block, the last ones by a few.  As many of the first as fit the L2 beside Zq are kept there.
            const int keep_ks = (p.dbg >> 8) & 0xff ? ((p.dbg >> 8) & 0xff) - 1 : p.keep_ks;
            const unsigned lt_mask = (1u << lane) - 1u;
            int pace_timeouts = 0;
            // is the ring slot of stage n free (the products of stage n - STAGES have completed)?
            auto slot_free = [&](long n) -> bool {
                return n < C::STAGES || mbar_test(empty + (int)(n % C::STAGES), ((unsigned)(n / C::STAGES) & 1u) ^ 1u);
            };
            // stage n (this lane's), its slot being free: header, barrier, copies.  ks < 0: header only (w = S: no
            // products).  w packs the leading-zero counts a | b << 4.
            auto emit_lane = [&](long n, const int8_t* za, const uint8_t* pb, int ks, unsigned w, unsigned flags) {
                const int rs = (int)(n % C::STAGES);
                const int a = (int)(w & 15u), b = (int)(w >> 4);
                unsigned long long* fb = full + (int)(n % C::NFULL);
                uint8_t* st = ring + rs * C::STAGE_BYTES;
                sh_hdr[rs] = (unsigned)a | ((unsigned)b << 8) | flags;
                if (ks < 0 || (p.dbg & 32)) {
                    mbar_arrive(fb);
                } else {
                    // slice pairs (i, j), i >= a, j >= b, i + j < S: A slices a .. S-1-b, B slices b .. S-1-a
                    const int ns = S - a - b;
                    const unsigned abytes = (unsigned)ns * I8_ATILE_BYTES, bbytes = (unsigned)ns * C::BTILE;
                    mbar_arrive_expect_tx(fb, abytes + bbytes);
                    bulk_g2s_hint(st + a * I8_ATILE_BYTES, za + (size_t)ks * (I8_SMAX * I8_ATILE_BYTES) + a * I8_ATILE_BYTES, abytes, fb, pol_keep);
                    bulk_g2s_hint(st + S * I8_ATILE_BYTES + b * C::BTILE, pb + (size_t)ks * (S * C::BTILE) + b * C::BTILE, bbytes, fb,
                                  ks < keep_ks ? pol_keep : pol_stream);
                    n_mma += (unsigned)(ns * (ns + 1) / 2);
                    ++n_stage;
                }
            };
            bool final_item = false;
            for (int item = blockIdx.x; item < p.ntiles; item += gridDim.x, ++it) {
                const int b = it & 1;
                final_item = item + (int)gridDim.x >= p.ntiles;
                I8_WAIT(panel_full + b, (unsigned)(it >> 1) & 1u, 2, it);
                fence_proxy_async();
                const uint8_t* pb = (p.dbg & 1) ? p.scratch : panels + (size_t)b * p.panel_bytes;
                const volatile uint8_t* plead = pb + p.lead_off;
                const unsigned item_flags = b ? H_ITEMPAR : 0u;
                for (int rb = 0; rb < nb; ++rb) {
                    // Pacing.  Every CTA walks the same row blocks of Zq, once per column tile.  When Zq is larger than
                    // the L2 and the CTAs drift apart, each of them streams it from HBM by itself (N = 8192: 0.8 GB per
                    // column tile, the kernel is HBM bound); kept within a window of row blocks of each other they share
                    // one pass through the L2.  A CTA counts itself in at row block rb and waits, for a bounded time,
                    // until every CTA has reached row block rb - window of this round of column tiles.
                    if (p.pace) {
                        if (lane == 0) atomicAdd(p.pace + rb, 1);
                        const int pace_window = ((p.dbg >> 12) & 15) ? ((p.dbg >> 12) & 15) : I8_PACE_WINDOW;
                        if (rb >= pace_window && pace_timeouts < 4) {
                            const long need_l = (long)(it + 1) * (long)gridDim.x;
                            const int need = need_l < (long)p.ntiles ? (int)need_l : p.ntiles;
                            const volatile int* pc = p.pace + (rb - pace_window);
                            const long long t0 = clock64();
                            for (;;) {
                                int seen = 0;
                                if (lane == 0) seen = *pc;
                                seen = __shfl_sync(0xffffffffu, seen, 0);
                                if (seen >= need) break;
                                // best effort: CTAs that are not co-resident (another kernel holds SMs) never arrive; after a
                                // few time-outs this CTA stops waiting for the rest of the launch
                                if (clock64() - t0 > I8_PACE_LIMIT) { ++pace_timeouts; break; }
                                __nanosleep(256);
                            }
                        }
                    }
                    const size_t kbase = (size_t)2 * rb * (rb + 1);
                    const int8_t* za = p.Zq + kbase * (size_t)(I8_SMAX * I8_ATILE_BYTES);
                    const uint8_t* zl = p.zlead + kbase;
                    const int nks = 4 * (rb + 1);
                    for (int ks0 = 0; ks0 < nks; ks0 += I8_KSEG, ++segs) {
                        const int ks1 = ks0 + I8_KSEG < nks ? ks0 + I8_KSEG : nks;
                        const bool item_last = rb == nb - 1 && ks1 == nks;
                        if (lane == 0) n_kstep += (unsigned)(ks1 - ks0);
                        constexpr int NCH = I8_KSEG / 32;
                        // pass 1: the lead bytes of the whole segment; wv[c]: this lane's k-step of chunk c (0xff: dead)
                        unsigned wv[NCH], live_mask[NCH];
                        int L = 0;
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            const int ks = ks0 + 32 * c + lane;
                            unsigned w = 0xffu;
                            if (ks < ks1) {
                                const unsigned a = __ldg(zl + ks), bb = plead[ks];
                                w = (ks == ks0 || noskip) ? 0u : (a + bb < (unsigned)S ? (a | (bb << 4)) : 0xffu);
                            }
                            wv[c] = w;
                        }
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            live_mask[c] = (ks0 + 32 * c < ks1) ? __ballot_sync(0xffffffffu, wv[c] != 0xffu) : 0u;
                            L += __popc(live_mask[c]);
                        }
                        // pass 2: every live lane emits its stage; j: its position among the live k-steps of the segment
                        const unsigned first_flags = item_flags | H_FIRST | (segs > 0 ? H_SEGNZ : 0u) | (((segs - 1) & 1) ? H_SEGPAR : 0u);
                        const unsigned second_flags = item_flags | H_SECOND | ((segs & 1) ? H_SEGPAR : 0u);
                        const unsigned endf = H_LAST | (item_last ? H_ITEM_LAST | (final_item ? H_FINAL : 0u) : 0u);
                        const unsigned penf = H_PENULT | (item_last ? H_ITEM_PENULT : 0u);
                        int before = 0;
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            if (live_mask[c] == 0u) continue;               // uniform
                            const int j = before + __popc(live_mask[c] & lt_mask);
                            unsigned f = j == 0 ? first_flags : j == 1 ? second_flags : item_flags;
                            if (j == L - 1 && L >= 2) f |= endf;
                            if (j == L - 2 || L == 1) f |= penf;
                            // rounds: whoever finds its slot free emits; the warp stays converged between the rounds, so a
                            // lane waiting for a slot never keeps back one whose slot is free (its stage may be what the
                            // issuers need next to free the other's).  A lane may look at its slot's barrier only once the
                            // stage STAGES before it has been emitted: one parity bit tells "the previous use is over"
                            // from "not yet", not from the uses before that -- so the window is the STAGES stages after
                            // the emitted prefix.
                            const int rank = __popc(live_mask[c] & lt_mask);
                            bool done = wv[c] == 0xffu;
                            for (;;) {
                                const unsigned rem = live_mask[c] & ~__ballot_sync(0xffffffffu, done);
                                if (rem == 0u) break;
                                const int prefix = __popc(live_mask[c] & ((rem & (0u - rem)) - 1u));      // live lanes below the first one not done
                                if (!done && rank < prefix + C::STAGES && slot_free(fills + j)) {
                                    emit_lane(fills + j, za, pb, ks0 + 32 * c + lane, wv[c], f);
                                    done = true;
                                }
                            }
                            before += __popc(live_mask[c]);
                        }
                        fills += L;
                        if (L == 1) {
                            // A segment always has two stages, so that both issuers take part in every segment and neither
                            // can run a whole segment ahead (the barriers carry one parity bit): an empty second stage.
                            bool done = lane != 0;
                            while (!__all_sync(0xffffffffu, done)) {
                                if (!done && slot_free(fills)) { emit_lane(fills, za, pb, -1, (unsigned)S, second_flags | endf); done = true; }
                            }
                            ++fills;
                        }
                    }
                }
            }
            {   // for the issuer that did not get the final stage
                bool done = lane != 0;
                while (!__all_sync(0xffffffffu, done)) {
                    if (!done && slot_free(fills)) { emit_lane(fills, nullptr, nullptr, -1, 0u, H_EXIT); done = true; }
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                n_mma += __shfl_xor_sync(0xffffffffu, n_mma, o);
                n_stage += __shfl_xor_sync(0xffffffffu, n_stage, o);
            }
            if (lane == 0) {
                atomicAdd(&g_i8_count[0], n_mma);
                atomicAdd(&g_i8_count[1], n_stage);
                atomicAdd(&g_i8_count[2], n_kstep);
                if (blockIdx.x == 0) atomicAdd(&g_i8_count[3], 1ull);
                atomicAdd(&g_i8_count[4], (unsigned long long)(clock64() - clk0));
                atomicAdd(&g_i8_count[5], 1ull);
            }
        }
    } else if (warp == C::ISSUER0 || warp == C::ISSUER1) {
        // ------------------------------ MMA issuers -----------------------------------------------
        // The tensor pipe takes one MMA at a time from a warp and queues nothing: every instruction its issuing warp
        // spends between two MMAs (waiting for the next stage, reading its header, the commit) is time the pipe idles
        // (tools/umma_probe4.cu: 1:1).  So two warps take the stages in turn, w issues stage n = w, w + 2, ..: while
        // one does its bookkeeping the other's MMAs fill the pipe.  The products of two stages may then interleave,
        // which integer accumulation does not notice; what has to be ordered is the start and the end of an
        // accumulation segment: the stage after the first waits until the first one (which overwrites) has been
        // issued (first_done), and both warps commit to acc_full / arrive on panel_empty when their part is done.
        // Each warp walks the loop whole (descriptors are warp-uniform); one elected lane issues.
        {
            const int w = warp == C::ISSUER0 ? 0 : 1;
            const unsigned ring_lo = i8_desc_lo(smem_u32(ring));
            const bool nomma = (p.dbg & 16) != 0;
            // stage n = w, w + MMA_WARPS, ..: ring slot n % STAGES, full barrier n % NFULL in its phase n / NFULL, kept as
            // small counters (64-bit divisions by constants in this loop are time the tensor pipe waits)
            int rs = w % C::STAGES, fi = w % C::NFULL;
            unsigned fph = 0;
            for (unsigned n = (unsigned)w;; n += C::MMA_WARPS) {
                I8_WAIT(full + fi, fph, 4, n);
                const unsigned h = sh_hdr[rs];
                if (h & H_EXIT) break;
                const int a = (int)(h & 0xffu), b = (int)((h >> 8) & 0xffu);
                const unsigned segpar = (h & H_SEGPAR) ? 1u : 0u;
                if (h & H_SECOND) I8_WAIT(first_done, segpar, 7, (unsigned)n);
                tc_fence_after();
                const unsigned a_lo = ring_lo + (unsigned)((rs * C::STAGE_BYTES) >> 4);
                const unsigned b_lo = a_lo + (unsigned)((S * I8_ATILE_BYTES) >> 4);
                if (h & H_FIRST) {
                    // the first stage of a segment overwrites the accumulators: each as soon as the epilogue has drained it
#pragma unroll
                    for (int d = 0; d < S; ++d) {
                        if (h & H_SEGNZ) {
                            I8_WAIT(acc_empty + d, segpar, 5, (unsigned)n);
                            tc_fence_after();
                        }
                        if (!nomma && elect_one_sync()) i8_issue_first<S, NC>(d, tbase, a_lo, b_lo);
                        __syncwarp();
                    }
                }
                if (elect_one_sync()) {
                    if (nomma || (h & H_FIRST)) {
                    } else {
                        i8_issue_ab<S, NC>(a, b, tbase, a_lo, b_lo);
                    }
                    umma_commit(empty + rs);
                    if (h & H_FIRST) mbar_arrive(first_done);
                    if (h & (H_LAST | H_PENULT)) umma_commit(acc_full);
                    if (h & (H_ITEM_LAST | H_ITEM_PENULT)) {      // every copy this warp waited for out of the item's panel has landed
                        mbar_arrive(panel_empty + ((h & H_ITEMPAR) ? 1 : 0));
                    }
                }
                __syncwarp();
                if (h & H_FINAL) break;
                rs += C::MMA_WARPS;
                if (rs >= C::STAGES) rs -= C::STAGES;
                fi += C::MMA_WARPS;
                if (fi >= C::NFULL) { fi -= C::NFULL; fph ^= 1u; }
            }
        }
    } else {
        // ------------------------------ generators ------------------------------------------------
        const int gtid = C::gen_index(warp) * 32 + lane;
        const HelmParams& hp = *reinterpret_cast<const HelmParams*>(sh_par);
        int it = 0;
        for (int item = blockIdx.x; item < p.ntiles; item += gridDim.x, ++it) {
            const int b = it & 1;
            // the whole first warp waits (the others wait for it at the named barrier): no divergence ahead of bar.sync
            if (it >= 2 && gtid < 32) I8_WAIT(panel_empty + b, (unsigned)((it >> 1) - 1) & 1u, 6, it);
            uint8_t* panel = panels + (size_t)b * p.panel_bytes;
            double mu0 = 0.0, mu1 = 0.0;
            const int gp0 = item * C::NG;
            const long long gclk0 = clock64();
            if (p.dbg & 4) {
                named_barrier(2, C::GT);
            } else if (hp.has_t) {
                if (hp.same_len) i8_generate_item<S, NC, true, true>(p, hp, panel, sh_stage, sh_or, sh_tab, gp0, gtid, mu0, mu1);
                else i8_generate_item<S, NC, false, true>(p, hp, panel, sh_stage, sh_or, sh_tab, gp0, gtid, mu0, mu1);
            } else {
                if (hp.same_len) i8_generate_item<S, NC, true, false>(p, hp, panel, sh_stage, sh_or, sh_tab, gp0, gtid, mu0, mu1);
                else i8_generate_item<S, NC, false, false>(p, hp, panel, sh_stage, sh_or, sh_tab, gp0, gtid, mu0, mu1);
            }
            if (gtid == 0) atomicAdd(&g_i8_count[6], (unsigned long long)(clock64() - gclk0));
            fence_proxy_async();                  // the panel is read back by bulk (async-proxy) copies
            const int g = gtid % C::NG, slot = gtid / C::NG;
            sh_mu[(slot * C::NG + g) * 2 + 0] = mu0;
            sh_mu[(slot * C::NG + g) * 2 + 1] = mu1;
            named_barrier(2, C::GT);
            if (gtid == 0) mbar_arrive(panel_full + b);
            if (gtid < NC) {
                const int pj = gtid >> 1, cc = gtid & 1;
                const int j = gp0 + pj;
                if (j < p.M) {
                    double m = 0.0;
#pragma unroll
                    for (int s = 0; s < C::GSLOTS; ++s) m += sh_mu[(s * C::NG + pj) * 2 + cc];
                    p.mean[(long)cc * p.out_stride + j] = m;
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == C::ISSUER0) tmem_dealloc(tbase, C::TMEM_COLS);
}

template <int S, int NC>
static size_t i8_panel_lead_off(int npad) { return (size_t)(npad / I8_KSTEP) * (S * I8Cfg<S, NC>::BTILE); }
template <int S, int NC>
static size_t i8_panel_bytes(int npad) {
    // digit slices, one byte per k-step; + 5 KB: consecutive panels must not sit at the same offset modulo a power
    // of two (see predict.cu)
    return i8_panel_lead_off<S, NC>(npad) + (size_t)round_up(npad / I8_KSTEP, 128) + 5120;
}
template <int S, int NC>
static size_t i8_cta_bytes(int npad) { return 2 * i8_panel_bytes<S, NC>(npad); }
size_t predict_i8_min_scratch_bytes(int npad) {          // one CTA
    const size_t a = i8_cta_bytes<6, 80>(npad), b = i8_cta_bytes<7, 64>(npad);
    return I8_PACE_BYTES + (a > b ? a : b);
}
size_t predict_i8_scratch_bytes(int npad) {
    const size_t a = i8_cta_bytes<6, 80>(npad), b = i8_cta_bytes<7, 64>(npad);
    return I8_PACE_BYTES + (size_t)predict_max_ctas() * (a > b ? a : b);
}

template <int S, int NC>
static cudaError_t predict_i8_launch(PredictI8Args a, double kmax, uint8_t* scratch, size_t scratch_bytes, cudaStream_t st) {
    using C = I8Cfg<S, NC>;
    static PerDeviceOnce once;
    const int slot = once.pending();
    if (slot >= 0) {
        cudaError_t e = cudaFuncSetAttribute(predict_i8_kernel<S, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        once.done[slot] = true;
    }
    // |K*| <= kmax: scale so that |k| kscale <= 2^(8 S - 2) (top digit within [-64, 64])
    int ex = 0;
    frexp(kmax, &ex);
    const int e = ex + 1;
    a.kscale = scalbn(1.0, (8 * S - 1) - e);
    a.cscale = scalbn(1.0, e - (8 * S - 1)) * scalbn(1.0, 8 * (S - 1)) * scalbn(1.0, 8 * (I8_SMAX - S));
    a.ntiles = (a.M + C::NG - 1) / C::NG;
    a.panel_bytes = i8_panel_bytes<S, NC>(a.npad);
    a.lead_off = i8_panel_lead_off<S, NC>(a.npad);
    a.cta_bytes = i8_cta_bytes<S, NC>(a.npad);
    if (scratch_bytes < I8_PACE_BYTES) return cudaErrorInvalidValue;
    a.scratch = scratch + I8_PACE_BYTES;
    long grid = a.ntiles;
    if (grid > predict_max_ctas()) grid = predict_max_ctas();
    const long panels = (long)((scratch_bytes - I8_PACE_BYTES) / a.cta_bytes);
    if (grid > panels) grid = panels;
    if (grid <= 0) return cudaErrorInvalidValue;
    const long per = (a.ntiles + grid - 1) / grid;
    grid = (a.ntiles + per - 1) / per;
    // pacing pays when the S slices of Z do not fit the L2 beside the panel stream and there are CTAs to keep together
    const bool big = (size_t)S * a.npad * (size_t)a.npad / 2 > ((size_t)64 << 20);
    a.pace = (((big && !(a.dbg & 64)) || (a.dbg & 128)) && grid > 1) ? reinterpret_cast<int*>(scratch) : nullptr;
    {   // L2 budget for panel tiles: what 96 MB leave beside the S slices of Z, shared by the CTAs' current panels
        const double zq = (double)S * a.npad * (double)a.npad / 2.0, budget = 96.0 * 1048576.0 - zq;
        const double per_kstep = (double)grid * S * C::BTILE;
        a.keep_ks = budget > 0.0 ? (int)(budget / per_kstep) : 0;
    }
    predict_i8_kernel<S, NC><<<(unsigned)grid, C::THREADS, C::SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

#ifdef GP2D_I8_WATCHDOG
extern "C" int gp2d_dbg_i8_watchdog(unsigned long long* out36) {      // [4] first time-out, [32] warp states of its CTA
    int ab = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&ab, g_i8_abort, sizeof(int));
    cudaMemcpyFromSymbol(out36, g_i8_wd, 4 * sizeof(unsigned long long));
    cudaMemcpyFromSymbol(out36 + 4, g_i8_state, 32 * sizeof(unsigned long long));
    const int zero = 0, minus = -1;
    unsigned long long z[32] = {};
    cudaMemcpyToSymbol(g_i8_abort, &zero, sizeof(int));
    cudaMemcpyToSymbol(g_i8_abort_cta, &minus, sizeof(int));
    cudaMemcpyToSymbol(g_i8_state, z, sizeof(z));
    return ab;
}
#endif

// out4: slice products (tcgen05.mma instructions of 128 x NC x 32) issued, stages issued, k-steps of the dense schedule,
// launches -- since the last call; clears them
extern "C" int gp2d_dbg_i8_counters(unsigned long long* out4) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out4, g_i8_count, 4 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    const unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    return cudaMemcpyToSymbol(g_i8_count, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}
// the same and, in out6[4], out6[5]: SM clocks the producers spent (summed over the CTAs of all launches), CTAs -- times
// in clocks do not move with the power cap the way milliseconds do
extern "C" int gp2d_dbg_i8_counters6(unsigned long long* out6) {      // out6[6]: clocks the generators were busy (one thread per CTA)
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out6, g_i8_count, 7 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    const unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    return cudaMemcpyToSymbol(g_i8_count, z, sizeof(z)) == cudaSuccess ? 0 : -1;
}

static thread_local int g_i8_dbg = 0;
void set_i8_debug(int v) { g_i8_dbg = v; }

// Launches both slice-count variants; each returns at once unless *gate names it (the choice was made at fit
// time, on the device side of the stream, so no host synchronisation is needed here).
cudaError_t predict_fused_i8(const int8_t* Zq, const double* zunit, const int* gate, int npad, const double* alpha_int,
                             const double* X, int N, const HelmParams& hp, const double* Xs, int M, long out_stride,
                             double var_add, double* mean, double* var, void* scratch, size_t scratch_bytes, cudaStream_t st,
                             int only_s) {
    if (M <= 0) return cudaSuccess;
    PredictI8Args a{};
    a.Zq = Zq; a.zlead = reinterpret_cast<const uint8_t*>(Zq) + i8_zq_slice_bytes(npad); a.zunit = zunit; a.gate = gate; a.npad = npad; a.alpha = alpha_int; a.X = X; a.N = N; a.hp = hp;
    a.Xs = Xs; a.M = M; a.out_stride = out_stride;
    a.kss = a.kss1 = hp.tvar * (hp.w_df + hp.w_cf);
    a.var_add = var_add; a.mean = mean; a.var = var;
    a.dbg = g_i8_dbg;
    // every entry of the 2x2 block is bounded by the prior variance k** (helmholtz.cuh)
    const double kmax = a.kss;
    // row-block arrival counters of the pacing (head of the scratch)
    cudaError_t e = scratch_bytes >= I8_PACE_BYTES ? cudaMemsetAsync(scratch, 0, I8_PACE_BYTES, st) : cudaErrorInvalidValue;
    if (e != cudaSuccess) return e;
    if (only_s == 0 || only_s == 6) e = predict_i8_launch<6, 80>(a, kmax, (uint8_t*)scratch, scratch_bytes, st);
    if (e != cudaSuccess) return e;
    if (only_s == 0 || only_s == 7) e = predict_i8_launch<7, 64>(a, kmax, (uint8_t*)scratch, scratch_bytes, st);
    return e;
}

}  // namespace gp2d
