// scan.cuh — device-wide exclusive prefix sum (u64 accumulators), three-phase (tile reduce, recursive scan of tile sums,
// tile down-sweep). Every variable-length stage of the engine (partition sizes, index entry sizes, compressed chunk
// sizes, head flags) is "size pass -> exclusive scan -> emit pass", so this is the glue between all kernels.
#pragma once
#include "common.cuh"

namespace b200c {

enum { SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS };

__device__ __forceinline__ uint64_t block_exclusive_scan_u64(uint64_t v, uint64_t* total, uint64_t* s_warp /*[8+1]*/) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = (lane < SCAN_THREADS / 32) ? s_warp[lane] : 0;
        uint64_t wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(FULL_MASK, wi, d); if (lane >= d) wi += t; }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = wi - w;
        if (lane == SCAN_THREADS / 32 - 1) s_warp[SCAN_THREADS / 32] = wi;
    }
    __syncthreads();
    uint64_t res = incl - v + s_warp[wid];
    *total = s_warp[SCAN_THREADS / 32];
    return res;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const TIn* __restrict__ in, uint64_t n, uint64_t* __restrict__ tile_sums) {
    __shared__ uint64_t s_warp[SCAN_THREADS / 32 + 1];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint64_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x;   // strided: coalesced
        if (i < n) sum += (uint64_t)in[i];
    }
    uint64_t total;
    block_exclusive_scan_u64(sum, &total, s_warp);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(const TIn* __restrict__ in, uint64_t n, const uint64_t* __restrict__ tile_offs,
                                                            uint64_t* __restrict__ out, uint64_t* __restrict__ grand_total) {
    __shared__ uint64_t s_warp[SCAN_THREADS / 32 + 1];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;   // blocked per thread
    uint64_t v[SCAN_ITEMS]; uint64_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { uint64_t i = base + k; v[k] = (i < n) ? (uint64_t)in[i] : 0; sum += v[k]; }
    uint64_t total;
    uint64_t off = block_exclusive_scan_u64(sum, &total, s_warp) + (tile_offs ? tile_offs[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { uint64_t i = base + k; if (i < n) out[i] = off; off += v[k]; }
    if (grand_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *grand_total = off;
}

} // namespace b200c
