// Deterministic block / grid reductions (fixed summation order: no fp64 atomics).
#pragma once
#include <cuda_runtime.h>

namespace gp2d {

// block-wide sum of NV values per thread; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* sh /* [NV][32] */) {
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        if (lane == 0) sh[q * 32 + wp] = v[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            double s = 0.0;
            for (int i = 0; i < nw; ++i) s += sh[q * 32 + i];
            v[q] = s;
        }
    }
}

// batch: CTA b reduces partial + b bstride into out + b bstride
template <int NV>
__global__ void final_reduce_kernel(const double* __restrict__ partial, int count, double* __restrict__ out, long bstride = 0) {
    partial += (long)blockIdx.x * bstride;
    out += (long)blockIdx.x * bstride;
    __shared__ double sh[NV * 32];
    double acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x)
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] += partial[(long)i * NV + q];
    block_reduce<NV>(acc, sh);
    if (threadIdx.x == 0)
#pragma unroll
        for (int q = 0; q < NV; ++q) out[q] = acc[q];
}

// one CTA per term: out[q][0..NV) = sum over that term's `count` partial rows
template <int NV>
__global__ void strided_final_reduce_kernel(const double* __restrict__ partial, int count, double* __restrict__ out, int out_stride) {
    __shared__ double sh[NV * 32];
    const double* src = partial + (size_t)blockIdx.x * count * NV;
    double acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x)
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] += src[(long)i * NV + q];
    block_reduce<NV>(acc, sh);
    if (threadIdx.x == 0)
#pragma unroll
        for (int q = 0; q < NV; ++q) out[(long)blockIdx.x * out_stride + q] = acc[q];
}

}  // namespace gp2d
