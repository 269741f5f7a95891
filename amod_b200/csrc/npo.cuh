// npo.cuh — NPO rebalancing match (rebalancing_npo.py:26-59) without materialising or sorting
// the idle x pending candidate list.  The reference sorts all pairs by rebl_dt (stable, pairs
// enumerated request-major) and scans greedily; that equals repeatedly committing every edge that
// is the minimum, under the total order (dt, request rank, vehicle rank), among all remaining
// edges sharing its request or its vehicle ("locally dominant" edges), which can be done for all
// such edges at once.  dt values are compared as float64 exactly as the reference does.
#pragma once
#include "common.cuh"

#define NPO_THREADS 256
#define NPO_CACHE_BYTES 32768

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

// Idle vehicles standing at the same node (a station, platform.py:50-55) have the same travel time to every request and are
// taken in rank order, so they form a GROUP: a node, its vehicles in fleet order and a head pointer.  A request's best free
// vehicle is the head of the group minimising (dt, head's rank); every vehicle of a group has the same best free request
// (min (dt, request rank)); an edge is committed when the two agree.  The searches run over groups (tens to thousands)
// instead of vehicles, and a minimum found in an earlier round stays valid for as long as that partner is free (the free sets
// only shrink), so most rounds touch a handful of requests and groups.
struct NpoGroups {
    const int* node;      // [n_groups]
    const int* start;     // [n_groups+1] into veh[]
    const int* veh;       // vehicle ranks, grouped by node, ascending within a group
    int* head;            // [n_groups] vehicles of the group already matched
    int n_groups;
};

// One block per group.  The block repeatedly (1) finds the group's best free request r under (dt, request rank), (2) checks that
// r prefers this group: no other group with free vehicles offers r a smaller (dt, rank of its head vehicle), and (3) commits
// r to the group's head vehicle.  It stops when the group is exhausted, no request is free, or its best request prefers
// another group for now (that request will be settled elsewhere, or come back once the other group runs out: next launch).
// All groups run concurrently; the decisions are safe under that concurrency because everything they read is monotone:
// a request never becomes free again, a group never regains a vehicle, and heads only move to higher ranks, so "r prefers
// g" can only turn from false to true.  A request is claimed by compare-and-swap, so it is matched exactly once.
// This is the reference's stable sort + greedy scan (rebalancing_npo.py:48-59): the edges of one group are accepted in
// (dt, request rank) order and each takes the lowest free vehicle rank, exactly as the sorted scan meets them.
template <typename TT>
__global__ void __launch_bounds__(NPO_THREADS)
k_npo_group_run(Table<TT> tt, NpoGroups g, const int* __restrict__ pend_onid, int n_pend, int* veh_match, int* req_match,
                double* __restrict__ match_dt, int* n_matched) {
    __shared__ Key smem[NPO_THREADS / 32];
    __shared__ int s_flag;
    __shared__ TT s_dt[NPO_CACHE_BYTES / sizeof(TT)];     // the group's travel times to the first requests: gathered once, not per step
    constexpr int NC = NPO_CACHE_BYTES / sizeof(TT);
    const int x = blockIdx.x;
    const int nid = g.node[x];
    const int size = g.start[x + 1] - g.start[x];
    volatile int* vhead = g.head;
    volatile int* vreq = req_match;
    if (vhead[x] >= size) return;
    // a request known to be taken carries +inf here: the free set as this block knows it (requests taken by other groups in the
    // meantime are discovered when the compare-and-swap fails; everything read is monotone, so a stale view only costs a retry)
    for (int r = threadIdx.x; r < n_pend && r < NC; r += NPO_THREADS) s_dt[r] = req_match[r] >= 0 ? (TT)INFINITY : tt.raw(nid, pend_onid[r]);
    __syncthreads();
    for (int step = 0; step < 2 * n_pend + 64; step++) {      // (bounded: a launch can never spin; the host relaunches while matches are made)
        const int head = vhead[x];
        if (head >= size) return;
        // (1) the group's best free request
        Key k{DINF, 0x7fffffff};
        for (int r = threadIdx.x; r < n_pend && r < NC; r += NPO_THREADS) {
            Key c{(double)s_dt[r], r};
            if (key_less(c, k)) k = c;
        }
        for (int r = NC + threadIdx.x; r < n_pend; r += NPO_THREADS) {      // beyond the cache: straight from memory
            if (vreq[r] >= 0) continue;
            Key c{tt(nid, pend_onid[r]), r};
            if (key_less(c, k)) k = c;
        }
        k = block_min_key(k, smem);
        if (k.idx == 0x7fffffff || !(k.dt < DINF)) return;
        const int r = k.idx, v = g.veh[g.start[x] + head], on = pend_onid[r];
        // taken by another group since this block last looked?  (ONE thread reads: the answer must be the same for the whole block)
        if (threadIdx.x == 0) {
            s_flag = vreq[r] >= 0 ? 2 : 1;
            if (s_flag == 2 && r < NC) s_dt[r] = (TT)INFINITY;      // forget it and look again
        }
        __syncthreads();
        if (s_flag == 2) { __syncthreads(); continue; }
        // (2) does r prefer this group?
        const Key mine{k.dt, v};
        bool lose = false;
        for (int y = threadIdx.x; y < g.n_groups && !lose; y += NPO_THREADS) {
            if (y == x) continue;
            const int hy = vhead[y];
            if (hy >= g.start[y + 1] - g.start[y]) continue;
            Key c{tt(g.node[y], on), g.veh[g.start[y] + hy]};
            if (key_less(c, mine)) lose = true;
        }
        if (lose) s_flag = 0;
        __syncthreads();
        if (!s_flag) return;
        // (3) commit
        if (threadIdx.x == 0) {
            if (atomicCAS(&req_match[r], -1, v) == -1) {
                veh_match[v] = r; match_dt[r] = k.dt;
                vhead[x] = head + 1;
                atomicAdd(n_matched, 1);
            }
            if (r < NC) s_dt[r] = (TT)INFINITY;       // taken, by this group or (compare-and-swap failed) by another
            __threadfence();
        }
        __syncthreads();
    }
}

template <typename TT>
__global__ void k_gather_bench(Table<TT> tt, const int* __restrict__ o, const int* __restrict__ d, float* __restrict__ out, i64 n) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; e < n; e += stride) out[e] = (float)tt(o[e], d[e]);
}
