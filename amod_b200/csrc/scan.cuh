// scan.cuh — device-wide exclusive prefix sum whose length lives in device memory.
// Two phases on a fixed grid: (a) every block reduces its contiguous segment, (b) every block sums the
// partials before it and scans its segment.  Phase (a) is a launch of its own only where no producer kernel
// can leave the block sums (scan_seg_len below).  The input is a functor, so flags derived
// from other arrays are scanned without being materialised; a finishing functor receives the
// total on one thread (it stores the count, checks the capacity of whatever the scan sizes).
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

__device__ __forceinline__ void scan_segment(i64 n, int vb, int vg, i64* lo, i64* hi) {   // block vb of vg
    i64 seg = (n + vg - 1) / vg;
    seg = (seg + SCAN_TILE - 1) / SCAN_TILE * SCAN_TILE;
    *lo = min(n, (i64)vb * seg);
    *hi = min(n, *lo + seg);
}

// A producer kernel can leave the per-block sums itself (atomicAdd into a zeroed array as it writes element i);
// then only phase (b) is launched.  `seg` = scan_seg_len(n, blocks of the scan).
__device__ __forceinline__ i64 scan_seg_len(i64 n, int vg) {
    i64 seg = (n + vg - 1) / vg;
    seg = (seg + SCAN_TILE - 1) / SCAN_TILE * SCAN_TILE;
    return seg > 0 ? seg : SCAN_TILE;
}

template <typename In>
__device__ __forceinline__ void scan_phase_a(const In& in, int vb, int vg, i64* __restrict__ partial, i64* smem) {
    const i64 n = in.size();
    i64 lo, hi;
    scan_segment(n, vb, vg, &lo, &hi);
    i64 s = 0;
    for (i64 i = lo + threadIdx.x; i < hi; i += SCAN_THREADS) s += in(i);
    i64 tot;
    block_excl_scan(s, &tot, smem);
    if (threadIdx.x == 0) partial[vb] = tot;
}

template <typename In, typename Fin>
__device__ __forceinline__ void scan_phase_b(const In& in, int vb, int vg, const i64* __restrict__ partial, i64* __restrict__ out,
                                             const Fin& fin, i64* smem) {
    const i64 n = in.size();
    i64 lo, hi;
    scan_segment(n, vb, vg, &lo, &hi);
    i64 p = 0;
    for (int b = threadIdx.x; b < vb; b += SCAN_THREADS) p += partial[b];
    i64 carry;
    block_excl_scan(p, &carry, smem);
    for (i64 base = lo; base < hi; base += SCAN_TILE) {
        const i64 i0 = base + (i64)threadIdx.x * SCAN_ITEMS;
        i64 v[SCAN_ITEMS];
        i64 s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = i0 + k < hi ? in(i0 + k) : 0; s += v[k]; }
        i64 tot;
        i64 e = block_excl_scan(s, &tot, smem) + carry;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { if (i0 + k < hi) out[i0 + k] = e; e += v[k]; }
        carry += tot;
    }
    if (vb == vg - 1 && threadIdx.x == 0) { out[n] = carry; fin(carry); }
}

template <typename In>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_a(In in, i64* __restrict__ partial) {
    pdl_enter();
    __shared__ i64 smem[SCAN_THREADS / 32 + 1];
    scan_phase_a(in, blockIdx.x, gridDim.x, partial, smem);
}

template <typename In, typename Fin>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_b(In in, const i64* __restrict__ partial, i64* __restrict__ out, Fin fin) {
    pdl_enter();
    __shared__ i64 smem[SCAN_THREADS / 32 + 1];
    scan_phase_b(in, blockIdx.x, gridDim.x, partial, out, fin, smem);
}

// two independent scans in one launch (their block sums come from the producer kernel): the first half of the
// grid does A, the second half B
template <typename InA, typename FinA, typename InB, typename FinB>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan2_b(InA a, FinA fa, i64* __restrict__ outa, InB b, FinB fb, i64* __restrict__ outb,
                                                          const i64* __restrict__ partial) {
    pdl_enter();
    __shared__ i64 smem[SCAN_THREADS / 32 + 1];
    const int half = gridDim.x >> 1;
    if ((int)blockIdx.x < half) scan_phase_b(a, blockIdx.x, half, partial, outa, fa, smem);
    else scan_phase_b(b, blockIdx.x - half, half, partial + half, outb, fb, smem);
}

// ---- input functors ---------------------------------------------------------------------------
template <typename T>
struct InArray {                     // in[i], length *n (int) or fixed
    const T* p; const int* n_ptr; i64 n_fixed;
    __device__ __forceinline__ i64 size() const { return n_ptr ? (i64)*n_ptr : n_fixed; }
    __device__ __forceinline__ i64 operator()(i64 i) const { return (i64)p[i]; }
};

// ---- finishing functors -----------------------------------------------------------------------
struct FinNone { __device__ __forceinline__ void operator()(i64) const {} };
// store min(total, cap) to *dst (int), remember the requirement, raise overflow
struct FinCountInt {
    int* dst; i64 cap; i64* need; unsigned long long* ovf; unsigned long long bit;
    __device__ __forceinline__ void operator()(i64 tot) const {
        if (tot > *need) *need = tot;
        if (tot > cap) { atomicOr(ovf, bit); tot = 0; }
        *dst = (int)tot;
    }
};
struct FinCountI64 {
    i64* dst; i64 cap; i64* need; unsigned long long* ovf; unsigned long long bit;
    __device__ __forceinline__ void operator()(i64 tot) const {
        if (tot > *need) *need = tot;
        if (tot > cap) { atomicOr(ovf, bit); tot = 0; }
        *dst = tot;
    }
};
