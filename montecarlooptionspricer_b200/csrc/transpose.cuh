// transpose.cuh -- out[c][r] = (TOut) in[r][c] through a padded 32x33 shared tile; both sides coalesced.
// Used for host-layout [path][step] <-> slab [step][path] and for draw tables [path][slot] <-> [slot][path].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) mcp_transpose_kernel(const TIn* __restrict__ in, int64_t in_ld, int64_t R, int64_t C,
                                                            TOut* __restrict__ out, int64_t out_ld, int64_t tiles_c) {
    __shared__ TOut tile[32][33];
    const int64_t t = blockIdx.x;
    const int64_t r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int64_t r = r0 + ty + k, c = c0 + tx;
        if (r < R && c < C) tile[ty + k][tx] = (TOut)in[r * in_ld + c];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
        const int64_t c = c0 + ty + k, r = r0 + tx;
        if (r < R && c < C) out[c * out_ld + r] = tile[tx][ty + k];
    }
}

template <typename TIn, typename TOut>
static inline void mcp_launch_transpose(cudaStream_t stream, const TIn* in, int64_t in_ld, int64_t R, int64_t C, TOut* out,
                                        int64_t out_ld) {
    const int64_t tiles_r = (R + 31) / 32, tiles_c = (C + 31) / 32;
    mcp_transpose_kernel<TIn, TOut><<<(unsigned)(tiles_r * tiles_c), 256, 0, stream>>>(in, in_ld, R, C, out, out_ld, tiles_c);
}
