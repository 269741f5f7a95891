// scan.cuh — device-wide exclusive prefix sum (reduce-then-scan, three launches).
// Used for every ordered compaction / CSR offset in the level-synchronous pool build.
#pragma once
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ i64 block_excl_scan(i64 x, i64* total, i64* smem /*[SCAN_THREADS/32 + 1]*/) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    i64 y = x;
    for (int o = 1; o < 32; o <<= 1) { i64 t = __shfl_up_sync(FULL, y, o); if (lane >= o) y += t; }
    if (lane == 31) smem[w] = y;
    __syncthreads();
    if (w == 0) {
        i64 s = lane < SCAN_THREADS / 32 ? smem[lane] : 0;
        i64 z = s;
        for (int o = 1; o < 32; o <<= 1) { i64 t = __shfl_up_sync(FULL, z, o); if (lane >= o) z += t; }
        if (lane < SCAN_THREADS / 32) smem[lane] = z - s;
        if (lane == 31) smem[SCAN_THREADS / 32] = z;
    }
    __syncthreads();
    i64 r = smem[w] + y - x;
    *total = smem[SCAN_THREADS / 32];
    __syncthreads();
    return r;
}

template <typename Tin>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const Tin* __restrict__ in, i64 n, i64* __restrict__ partial) {
    __shared__ i64 smem[SCAN_THREADS / 32 + 1];
    i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_ITEMS;
    i64 s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) s += (i64)in[base + k];
    i64 tot;
    block_excl_scan(s, &tot, smem);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// one block: exclusive scan of the per-tile totals (any number of tiles), grand total to *total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_partials(i64* __restrict__ partial, int nb, i64* __restrict__ total) {
    __shared__ i64 smem[SCAN_THREADS / 32 + 1];
    i64 carry = 0;
    for (int b0 = 0; b0 < nb; b0 += SCAN_THREADS) {
        int b = b0 + threadIdx.x;
        i64 x = b < nb ? partial[b] : 0;
        i64 tot;
        i64 e = block_excl_scan(x, &tot, smem);
        if (b < nb) partial[b] = carry + e;
        carry += tot;
    }
    if (threadIdx.x == 0) *total = carry;
}

template <typename Tin>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(const Tin* __restrict__ in, i64 n, const i64* __restrict__ partial,
                                                            i64* __restrict__ out) {
    __shared__ i64 smem[SCAN_THREADS / 32 + 1];
    i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_ITEMS;
    i64 v[SCAN_ITEMS];
    i64 s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? (i64)in[base + k] : 0; s += v[k]; }
    i64 tot;
    i64 e = block_excl_scan(s, &tot, smem) + partial[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = e; e += v[k]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) out[n] = e;   // out has n+1 entries
}
