// npo.cuh — NPO rebalancing match (rebalancing_npo.py:26-59) without materialising or sorting
// the idle x pending candidate list.  The reference sorts all pairs by rebl_dt (stable, pairs
// enumerated request-major) and scans greedily; that equals repeatedly committing every edge that
// is the minimum, under the total order (dt, request rank, vehicle rank), among all remaining
// edges sharing its request or its vehicle ("locally dominant" edges), which can be done for all
// such edges at once.  dt values are compared as float64 exactly as the reference does.
#pragma once
#include "common.cuh"

#define NPO_THREADS 256

struct Key { double dt; int idx; };
__device__ __forceinline__ bool key_less(const Key& a, const Key& b) { return a.dt < b.dt || (a.dt == b.dt && a.idx < b.idx); }

__device__ __forceinline__ Key block_min_key(Key k, Key* smem) {
    for (int o = 16; o; o >>= 1) {
        Key t; t.dt = __shfl_xor_sync(FULL, k.dt, o); t.idx = __shfl_xor_sync(FULL, k.idx, o);
        if (key_less(t, k)) k = t;
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem[w] = k;
    __syncthreads();
    if (w == 0) {
        k = lane < NPO_THREADS / 32 ? smem[lane] : Key{DINF, 0x7fffffff};
        for (int o = 16; o; o >>= 1) {
            Key t; t.dt = __shfl_xor_sync(FULL, k.dt, o); t.idx = __shfl_xor_sync(FULL, k.idx, o);
            if (key_less(t, k)) k = t;
        }
        if (lane == 0) smem[0] = k;
    }
    __syncthreads();
    k = smem[0];
    __syncthreads();
    return k;
}

// side 0: one block per pending request -> its best free vehicle; side 1: per idle vehicle -> best free request
template <typename TT>
__global__ void __launch_bounds__(NPO_THREADS)
k_npo_best(Table<TT> tt, const int* __restrict__ idle_nid, int n_idle, const int* __restrict__ pend_onid, int n_pend,
           const int* __restrict__ veh_match, const int* __restrict__ req_match, int side, int* __restrict__ best,
           double* __restrict__ best_dt) {
    __shared__ Key smem[NPO_THREADS / 32];
    const int a = blockIdx.x;
    Key k{DINF, 0x7fffffff};
    if (side == 0) {
        if (req_match[a] >= 0) return;
        const int on = pend_onid[a];
        for (int v = threadIdx.x; v < n_idle; v += NPO_THREADS) {
            if (veh_match[v] >= 0) continue;
            Key c{tt(idle_nid[v], on), v};
            if (key_less(c, k)) k = c;
        }
    } else {
        if (veh_match[a] >= 0) return;
        const int nid = idle_nid[a];
        for (int r = threadIdx.x; r < n_pend; r += NPO_THREADS) {
            if (req_match[r] >= 0) continue;
            Key c{tt(nid, pend_onid[r]), r};
            if (key_less(c, k)) k = c;
        }
    }
    k = block_min_key(k, smem);
    if (threadIdx.x == 0) { best[a] = k.idx; if (best_dt) best_dt[a] = k.dt; }
}

__global__ void k_npo_commit(int n_pend, const int* __restrict__ rbest, const double* __restrict__ rbest_dt,
                             const int* __restrict__ vbest, int* veh_match, int* req_match, double* match_dt, int* n_matched) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pend || req_match[r] >= 0) return;
    const int v = rbest[r];
    if (v == 0x7fffffff || vbest[v] != r) return;
    req_match[r] = v; veh_match[v] = r; match_dt[r] = rbest_dt[r];
    atomicAdd(n_matched, 1);
}

template <typename TT>
__global__ void k_gather_bench(Table<TT> tt, const int* __restrict__ o, const int* __restrict__ d, float* __restrict__ out, i64 n) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; e < n; e += stride) out[e] = (float)tt(o[e], d[e]);
}
