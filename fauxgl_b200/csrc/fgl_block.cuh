// fgl_block.cuh -- warp/block scan primitives shared by the kernels.
#pragma once
#include <cstdint>

namespace fgl {

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread; returns the exclusive
// prefix and writes the block total to *total (valid for all threads).
// smem needs THREADS/32+1 words.  Ends with a barrier, so smem can be reused.
template <int THREADS>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *smem, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v);
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < THREADS / 32 ? smem[lane] : 0;
        uint32_t si = warp_incl_scan(s);
        if (lane < THREADS / 32) smem[lane] = si - s;
        if (lane == THREADS / 32 - 1) smem[THREADS / 32] = si;
    }
    __syncthreads();
    uint32_t r = smem[warp] + incl - v;
    *total = smem[THREADS / 32];
    __syncthreads();
    return r;
}

// The same with ONE barrier: the warp totals go to one of two alternating buffers (smem[2][THREADS/32]) and
// every thread adds up the totals of the warps before its own.  `parity` must alternate between successive
// calls that share the buffers (the barrier of call k orders the reads of call k-1 before the writes of k+1).
template <int THREADS>
__device__ __forceinline__ uint32_t block_excl_scan1(uint32_t v, uint32_t (*smem)[THREADS / 32], uint32_t &parity,
                                                     uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t incl = warp_incl_scan(v);
    uint32_t *buf = smem[parity & 1u];
    parity ^= 1u;
    if (lane == 31) buf[warp] = incl;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) {
        const uint32_t t = buf[w];
        if (w < warp) before += t;
        all += t;
    }
    *total = all;
    return before + incl - v;
}

}  // namespace fgl
