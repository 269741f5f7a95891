// npo.cuh — NPO rebalancing match (rebalancing_npo.py:26-59) without materialising or sorting
// the idle x pending candidate list.  The reference sorts all pairs by rebl_dt (stable, pairs
// enumerated request-major) and scans greedily; that equals repeatedly committing every edge that
// is the minimum, under the total order (dt, request rank, vehicle rank), among all remaining
// edges sharing its request or its vehicle ("locally dominant" edges), which can be done for all
// such edges at once.  dt values are compared as float64 exactly as the reference does.
#pragma once
#include "common.cuh"

#define NPO_THREADS 1024
#define NPO_SCAP 2048
#define NPO_UNROLL 4        // requests a thread has in flight per step of a pass

struct Key { double dt; int idx; };
__device__ __forceinline__ bool key_less(const Key& a, const Key& b) { return a.dt < b.dt || (a.dt == b.dt && a.idx < b.idx); }
__device__ __forceinline__ bool key_eq(const Key& a, const Key& b) { return a.dt == b.dt && a.idx == b.idx; }
#define KEY_INF Key{DINF, 0x7fffffff}
struct NpoPick { double dt; int r; int pad; };

__device__ __forceinline__ Key warp_min_key(Key k) {
    for (int o = 16; o; o >>= 1) {
        Key t; t.dt = __shfl_xor_sync(FULL, k.dt, o); t.idx = __shfl_xor_sync(FULL, k.idx, o);
        if (key_less(t, k)) k = t;
    }
    return k;
}
// Idle vehicles standing at the same node (a station, platform.py:50-55) have the same travel time to every request and are
// taken in rank order, so they form a GROUP: a node, its vehicles in fleet order and a head pointer (vehicles already matched).
// A request's best free vehicle is the head of the group minimising (dt, head's rank); every free vehicle of a group has the
// same order of preference over the free requests, (dt, request rank).
//
// Rounds.  k_npo_best gives every free request its best group.  k_npo_accept (one block per group) then accepts, all at once,
// the longest PREFIX of the group's preference order made of requests that name this group as their best: its i-th element
// goes to the group's i-th free vehicle.  Why that is the reference's stable sort + greedy scan (rebalancing_npo.py:48-59):
// by induction over the prefix, once elements 0..i-1 are matched the i-th free vehicle's smallest remaining edge is element i,
// and element i's smallest remaining edge leads into this group (to that very vehicle, the lowest free rank) - the edge is
// locally dominant, and a locally dominant edge is always taken by the sorted scan; matches made elsewhere in the same round
// only remove alternatives.  Ties: "best" was decided with the group's HEAD vehicle; a request whose dt ties with another
// group's may lose that tie against a later vehicle of this group, so such a request is only accepted as element 0 and ends
// the prefix anywhere else.  The globally smallest remaining edge is always element 0 of its group, so every round makes
// progress; a station with a thousand idle vehicles is served in a few rounds instead of a thousand dependent steps.
struct NpoGroups {
    const int* node;      // [n_groups]
    const int* start;     // [n_groups+1] into veh[]
    const int* veh;       // vehicle ranks, grouped by node, ascending within a group
    int* head;            // [n_groups] vehicles of the group already matched
    int* hrank;           // [n_groups] rank of the group's head vehicle, -1 when the group is exhausted
    int n_groups;
};

// D[g][r] = travel time from group g's node to request r's origin (the only table look-ups of the match), hrank
template <typename TT>
__global__ void k_npo_matrix(Table<TT> tt, NpoGroups g, const int* __restrict__ pend_onid, int n_pend, TT* __restrict__ D, TT* __restrict__ DT) {
    const i64 n = (i64)g.n_groups * n_pend, stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const int x = (int)(e / n_pend), r = (int)(e - (i64)x * n_pend);
        const TT d = tt.raw(g.node[x], pend_onid[r]);
        D[e] = d;                                   // [group][request]: a group's block walks its row (k_npo_accept)
        DT[(i64)r * g.n_groups + x] = d;            // [request][group]: a request's warp walks its row (k_npo_best)
    }
    for (i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x; x < g.n_groups; x += stride) g.hrank[x] = g.veh[g.start[x]];
}

// Everything the rounds change is read with ld.global.cg (L2): the persistent variant runs many rounds inside one launch, and a
// line cached in an SM's L1 in an earlier round would be stale.

// per free request (one warp each, lanes over the groups): the group offering the smallest (dt, rank of its head vehicle), and
// whether another group ties in dt
template <typename TT>
__device__ __forceinline__ void npo_best_request(const TT* __restrict__ DT, const NpoGroups& g, int r, int lane,
                                                 int* __restrict__ best_g, uint8_t* __restrict__ tie) {
    const TT* __restrict__ row = DT + (i64)r * g.n_groups;
    Key best = KEY_INF;
    int bg = -1, same = 0;                       // same = groups seen with best.dt besides the best one
    for (int x = lane; x < g.n_groups; x += 32) {
        const TT d = row[x];                     // (both loads in flight before the first is looked at)
        const int hr = ldcg_i(g.hrank + x);
        if (hr < 0) continue;
        const Key c{(double)d, hr};
        if (c.dt == best.dt) same++;
        else if (c.dt < best.dt) same = 0;
        if (key_less(c, best)) { best = c; bg = x; }
    }
    for (int o = 16; o; o >>= 1) {
        Key t; t.dt = __shfl_xor_sync(FULL, best.dt, o); t.idx = __shfl_xor_sync(FULL, best.idx, o);
        const int tg = __shfl_xor_sync(FULL, bg, o), ts = __shfl_xor_sync(FULL, same, o);
        if (tg < 0) continue;                    // (the partner saw no group at all)
        if (bg < 0) { best = t; bg = tg; same = ts; continue; }
        if (t.dt == best.dt) same += ts + 1;
        else if (t.dt < best.dt) same = ts;
        if (key_less(t, best)) { best = t; bg = tg; }
    }
    if (lane == 0) { best_g[r] = bg; tie[r] = same > 0; }
}

template <typename TT>
__global__ void k_npo_best(const TT* __restrict__ DT, NpoGroups g, int n_pend, const int* __restrict__ req_match,
                           int* __restrict__ best_g, uint8_t* __restrict__ tie, const int* __restrict__ n_matched, int target,
                           unsigned long long* __restrict__ cursor) {
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    if (gt == 0) *cursor = 0ull;                 // the accept kernel's scratch list starts empty every round
    if (*n_matched >= target) return;
    const int r = gt >> 5, lane = threadIdx.x & 31;
    if (r >= n_pend || req_match[r] >= 0) return;     // (warp-uniform)
    npo_best_request(DT, g, r, lane, best_g, tie);
}

// (first, A, t1 <= t2) of a set of requests, see npo_accept_group; two sets are merged element-wise (smallest two of the four ties)
struct NpoQuad { Key first, A, t1, t2; };
__device__ __forceinline__ Key key_shfl_xor(const Key& k, int o) {
    Key t; t.dt = __shfl_xor_sync(FULL, k.dt, o); t.idx = __shfl_xor_sync(FULL, k.idx, o);
    return t;
}
__device__ __forceinline__ NpoQuad quad_shfl_xor(const NpoQuad& q, int o) {
    return NpoQuad{key_shfl_xor(q.first, o), key_shfl_xor(q.A, o), key_shfl_xor(q.t1, o), key_shfl_xor(q.t2, o)};
}
__device__ __forceinline__ void quad_merge(NpoQuad& a, const NpoQuad& b) {
    if (key_less(b.first, a.first)) a.first = b.first;
    if (key_less(b.A, a.A)) a.A = b.A;
    if (key_less(b.t1, a.t1)) { a.t2 = key_less(a.t1, b.t2) ? a.t1 : b.t2; a.t1 = b.t1; }
    else if (key_less(b.t1, a.t2)) a.t2 = b.t1;
}

struct NpoShared {
    NpoQuad quad[NPO_THREADS / 32];
    int s_w;
    unsigned long long s_base;
    NpoPick s_pick[NPO_SCAP];                    // the accepted set of this round (beyond NPO_SCAP: a stretch of the global scratch list)
};

// one round of group x (the whole block); every exit is taken by all threads together
template <typename TT>
__device__ __forceinline__ void npo_accept_group(NpoShared& sh, const TT* __restrict__ D, const NpoGroups& g, int x, int n_pend, int* veh_match,
                                                 int* req_match, double* __restrict__ match_dt, const int* __restrict__ best_g,
                                                 const uint8_t* __restrict__ tie, NpoPick* __restrict__ scratch,
                                                 unsigned long long* __restrict__ cursor, int* n_matched) {
    int& s_w = sh.s_w;
    unsigned long long& s_base = sh.s_base;
    NpoPick* s_pick = sh.s_pick;
    const int size = g.start[x + 1] - g.start[x];
    const int head = ldcg_i(g.head + x);
    const int c = size - head;
    if (c <= 0) return;
    const TT* __restrict__ Dg = D + (i64)x * n_pend;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // pass 1: first = head of the group's preference order; A = its first request that names another group; t1 < t2 = its first
    // two requests that name this group but tie in dt with another
    Key first = KEY_INF, A = KEY_INF, t1 = KEY_INF, t2 = KEY_INF;
    for (int r0 = threadIdx.x; r0 < n_pend; r0 += NPO_UNROLL * NPO_THREADS) {
        int rm[NPO_UNROLL], bg[NPO_UNROLL]; TT d[NPO_UNROLL]; uint8_t ti[NPO_UNROLL];
#pragma unroll
        for (int u = 0; u < NPO_UNROLL; u++) {
            const int r = r0 + u * NPO_THREADS;
            rm[u] = 0; bg[u] = -1; d[u] = (TT)0; ti[u] = 0;
            if (r < n_pend) { rm[u] = ldcg_i(req_match + r); d[u] = Dg[r]; bg[u] = ldcg_i(best_g + r); ti[u] = (uint8_t)ldcg_b(tie + r); }
        }
#pragma unroll
        for (int u = 0; u < NPO_UNROLL; u++) {
            const int r = r0 + u * NPO_THREADS;
            if (r >= n_pend || rm[u] >= 0) continue;
            const Key k{(double)d[u], r};
            if (key_less(k, first)) first = k;
            if (bg[u] != x) { if (key_less(k, A)) A = k; }
            else if (ti[u]) { if (key_less(k, t1)) { t2 = t1; t1 = k; } else if (key_less(k, t2)) t2 = k; }
        }
    }
    {   // one block-wide reduction for all four
        NpoQuad q{first, A, t1, t2};
        for (int o = 16; o; o >>= 1) quad_merge(q, quad_shfl_xor(q, o));
        if (lane == 0) sh.quad[w] = q;
        __syncthreads();
        if (w == 0) {
            q = lane < NPO_THREADS / 32 ? sh.quad[lane] : NpoQuad{KEY_INF, KEY_INF, KEY_INF, KEY_INF};
            for (int o = 16; o; o >>= 1) quad_merge(q, quad_shfl_xor(q, o));
            if (lane == 0) sh.quad[0] = q;
        }
        __syncthreads();
        q = sh.quad[0];
        __syncthreads();
        first = q.first; A = q.A; t1 = q.t1; t2 = q.t2;
    }
    if (first.idx == 0x7fffffff) return;         // no free request
    const Key B = key_eq(t1, first) ? t2 : t1;   // a tying request may only be accepted as element 0
    const Key stop = key_less(A, B) ? A : B;
    // pass 2: the requests before `stop` (all name this group), into the shared list; counted as they go
    if (threadIdx.x == 0) s_w = 0;
    __syncthreads();
    NpoPick* S = s_pick;
    for (int attempt = 0; attempt < 2; attempt++) {
        for (int r0 = threadIdx.x; r0 < n_pend; r0 += NPO_UNROLL * NPO_THREADS) {
            int rm[NPO_UNROLL], bg[NPO_UNROLL]; TT d[NPO_UNROLL];
#pragma unroll
            for (int u = 0; u < NPO_UNROLL; u++) {
                const int r = r0 + u * NPO_THREADS;
                rm[u] = 0; bg[u] = -1; d[u] = (TT)0;
                if (r < n_pend) { rm[u] = ldcg_i(req_match + r); d[u] = Dg[r]; bg[u] = ldcg_i(best_g + r); }
            }
#pragma unroll
            for (int u = 0; u < NPO_UNROLL; u++) {
                const int r = r0 + u * NPO_THREADS;
                if (r >= n_pend || rm[u] >= 0 || bg[u] != x) continue;
                const Key k{(double)d[u], r};
                if (key_less(k, stop)) {
                    const int p = atomicAdd(&s_w, 1);
                    if (attempt || p < NPO_SCAP) { S[p].dt = k.dt; S[p].r = r; }
                }
            }
        }
        __syncthreads();
        if (attempt || s_w <= NPO_SCAP) break;
        // too many for shared memory: claim a stretch of the global list (the accepted sets of one round are disjoint: <= n_pend in total)
        const int need = s_w;
        __syncthreads();
        if (threadIdx.x == 0) { s_base = atomicAdd(cursor, (unsigned long long)need); s_w = 0; }
        __syncthreads();
        S = scratch + s_base;
    }
    const int m = s_w;
    if (m == 0) return;                          // the head of this group's order names another group: wait for it
    // pass 3: rank within the accepted set = which of the group's free vehicles takes the request
    const int* __restrict__ vv = g.veh + g.start[x] + head;
    for (int e = threadIdx.x; e < m; e += NPO_THREADS) {
        const Key k{S[e].dt, S[e].r};
        int rank = 0;
        for (int y = 0; y < m; y++) { const Key u{S[y].dt, S[y].r}; rank += key_less(u, k) ? 1 : 0; }
        if (rank < c) {
            const int v = vv[rank];
            req_match[k.idx] = v; veh_match[v] = k.idx; match_dt[k.idx] = k.dt;
        }
    }
    if (threadIdx.x == 0) {
        const int acc = min(m, c);
        g.head[x] = head + acc;
        g.hrank[x] = head + acc < size ? g.veh[g.start[x] + head + acc] : -1;
        atomicAdd(n_matched, acc);
    }
    __syncthreads();                             // (the shared list is reused by the next round of the persistent variant)
}

template <typename TT>
__global__ void __launch_bounds__(NPO_THREADS)
k_npo_accept(const TT* __restrict__ D, NpoGroups g, int n_pend, int* veh_match, int* req_match, double* __restrict__ match_dt,
             const int* __restrict__ best_g, const uint8_t* __restrict__ tie, NpoPick* __restrict__ scratch,
             unsigned long long* __restrict__ cursor, int* n_matched) {
    __shared__ NpoShared sh;
    npo_accept_group(sh, D, g, (int)blockIdx.x, n_pend, veh_match, req_match, match_dt, best_g, tie, scratch, cursor, n_matched);
}

// All rounds in ONE launch when every group's block is resident at once (cooperative launch: tens of groups - the stations - is
// the case that needs many rounds): the two phases of a round are separated by a device-wide barrier (arrive with a release,
// spin on an acquire load; bounded, so a fault cannot hang the device) instead of a kernel boundary, and the host synchronises once.
template <typename TT>
__global__ void __launch_bounds__(NPO_THREADS)
k_npo_rounds(const TT* __restrict__ D, const TT* __restrict__ DT, NpoGroups g, int n_pend, int* veh_match, int* req_match,
             double* __restrict__ match_dt, int* best_g, uint8_t* tie, NpoPick* __restrict__ scratch,
             unsigned long long* cursor, int* n_matched, int target, unsigned* bar, int max_rounds, int* rounds_out, int* err) {
    __shared__ NpoShared sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = gridDim.x * (NPO_THREADS / 32);
    unsigned phase = 0;
    int round = 0;
    while (round < max_rounds) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *cursor = 0ull;
        for (int r = blockIdx.x * (NPO_THREADS / 32) + warp; r < n_pend; r += nw)
            if (ldcg_i(req_match + r) < 0) npo_best_request(DT, g, r, lane, best_g, tie);
        if (!grid_barrier(bar, ++phase * gridDim.x, err)) break;
        if ((int)blockIdx.x < g.n_groups)       // (the grid fills the device: the blocks beyond the groups only serve the requests' phase)
            npo_accept_group(sh, D, g, (int)blockIdx.x, n_pend, veh_match, req_match, match_dt, best_g, tie, scratch, cursor, n_matched);
        round++;
        if (!grid_barrier(bar, ++phase * gridDim.x, err)) break;
        if (ldcg_i(n_matched) >= target) break;       // (read after the barrier: the same answer in every block)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *rounds_out = round;
}

// The reference's selection order: ascending (dt, request rank) over the matched requests.  A warp per matched request counts
// the matched requests before it and writes the triple to that place (used up to NPO_ORDER_MAX requests; beyond, the host sorts).
#define NPO_ORDER_MAX 16384
__global__ void k_npo_order(const int* __restrict__ req_match, const double* __restrict__ match_dt, int n_pend,
                            int* __restrict__ out_idle, int* __restrict__ out_pend, double* __restrict__ out_dt) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= n_pend) return;
    const int v = req_match[r];
    if (v < 0) return;
    const Key k{match_dt[r], r};
    int before = 0;
    for (int y = lane; y < n_pend; y += 32)
        if (req_match[y] >= 0) { const Key u{match_dt[y], y}; before += key_less(u, k) ? 1 : 0; }
    before = warp_sum(before);
    if (lane == 0) { out_idle[before] = v; out_pend[before] = r; out_dt[before] = k.dt; }
}

template <typename TT>
__global__ void k_gather_bench(Table<TT> tt, const int* __restrict__ o, const int* __restrict__ d, float* __restrict__ out, i64 n) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; e < n; e += stride) out[e] = (float)tt(o[e], d[e]);
}
