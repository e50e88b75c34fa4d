// tcgen05 (5th-generation tensor core) helpers for the int8-sliced predictive kernel (predict_i8.cu):
// TMEM allocation, kind::i8 MMA with both operands in shared memory, commit -> mbarrier, TMEM loads, and
// the shared-memory operand image / descriptor the MMA reads.
//
// Operand tiles are K-major [rows x 32 bytes] int8 in the canonical SWIZZLE_NONE layout: 8-row x 16-byte
// core matrices, the two 16-byte halves of K 128 bytes apart (leading byte offset), 8-row groups 256 bytes
// apart (stride byte offset).  tools/umma_probe*.cu checked this image (and the 32/64/128-byte swizzled
// ones) against a CPU product on a B200 and measured the issue rates the kernel design rests on
// (profiles/r02a_umma_probe*.log): an SS-mode MMA re-reads its A tile (4 KB) from shared memory, so at
// M = 128 one instruction costs max(N / 2, (4096 + 32 N) / 128) clocks whatever the swizzle.
#pragma once
#include <stdint.h>

#include "pipeline.cuh"

namespace gp2d {

constexpr int I8_KSTEP = 32;                       // K of one tcgen05.mma kind::i8
constexpr int I8_ATILE_BYTES = 128 * I8_KSTEP;     // 128-row operand tile image

// byte offset of element (row r, k) inside a [rows x 32] tile image
__host__ __device__ __forceinline__ int i8_tile_off(int r, int k) {
    return (r >> 3) * 256 + (k >> 4) * 128 + (r & 7) * 16 + (k & 15);
}

// shared-memory matrix descriptor: start address, LBO = 128, SBO = 256 (16-byte units), version 1, no swizzle
constexpr unsigned I8_DESC_HI = (256u >> 4) | (1u << 14);
__device__ __forceinline__ unsigned i8_desc_lo(unsigned saddr) { return ((saddr & 0x3FFFFu) >> 4) | ((128u >> 4) << 16); }
__device__ __forceinline__ uint64_t i8_desc(unsigned lo) { return ((uint64_t)I8_DESC_HI << 32) | (uint64_t)lo; }

// instruction descriptor: D = S32, A = B = signed int8, both K-major, M x N
__host__ __device__ constexpr unsigned i8_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(unsigned* smem_slot, unsigned ncols) {      // one full warp; ncols a power of two >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {          // the warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// D[tmem] (+)= A[smem] B[smem]^T, issued by ONE thread
__device__ __forceinline__ void umma_i8_ss(unsigned d_tmem, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the mbarrier gets one arrival when every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane (warp w of the CTA reads lanes 32 (w % 4) ..)
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(unsigned taddr, int (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, int (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one_sync() {
    unsigned pred;
    asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void named_barrier(int id, int threads) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory"); }

// Balanced base-256 digits.  q is a signed fixed-point integer of at most 8 S - 2 bits; adding 0x80 to each
// of its S low bytes makes byte p of the sum equal digit_p + 128 with digit_p in [-128, 127] (the carries
// of the classical "subtract 256 when the byte is >= 128" recurrence are exactly those of this addition),
// so digit_p as an int8 is that byte with its top bit flipped.  Dropping low digits of a balanced
// representation rounds to nearest: the first S' digits of an S-digit number are its S'-digit rounding.
__host__ __device__ constexpr long long i8_digit_bias(int S) {
    long long b = 0;
    for (int p = 0; p < S; ++p) b |= 0x80ll << (8 * p);
    return b;
}
// word holding byte P of four biased values (x0 lowest byte ... x3 highest), top bits flipped -> four int8 digits
template <int P>
__device__ __forceinline__ unsigned i8_digit_word(long long x0, long long x1, long long x2, long long x3) {
    const unsigned a0 = P < 4 ? (unsigned)x0 : (unsigned)(x0 >> 32), a1 = P < 4 ? (unsigned)x1 : (unsigned)(x1 >> 32);
    const unsigned a2 = P < 4 ? (unsigned)x2 : (unsigned)(x2 >> 32), a3 = P < 4 ? (unsigned)x3 : (unsigned)(x3 >> 32);
    constexpr unsigned B = P & 3, SEL = B | ((4u + B) << 4);
    const unsigned t01 = __byte_perm(a0, a1, SEL), t23 = __byte_perm(a2, a3, SEL);
    return __byte_perm(t01, t23, 0x5410) ^ 0x80808080u;
}

}  // namespace gp2d
