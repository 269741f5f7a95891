// place.cuh — from the insertion search's unordered records to the level's schedule store.
//
// The search (eval.cuh) leaves one 16 B record per feasible insertion in per-block regions, in no particular
// order.  A record's position in the reference's ENCOUNTER order (the loop order of scheduling.py:22-27) is
//     chunk_pos[c] + (feasible insertions of the chunk's lanes before its lane) + (ordinal within its lane),
// all known once the prefix sum over the chunks' counts is done.
//   k_place_cost   stores every record's cost at its encounter position (what list-order classification reads)
//   k_classify*    (levels.cuh) turns the encounter-order costs of each trip into the reference's list order
//                  (scheduling.py:31-36) and leaves final[encounter position] = row
//   k_place_rows   builds each schedule (base schedule + the inserted pick-up at i and drop-off at j) straight
//                  into its final row: 16 B stops (code, node, deadline), eight lanes per schedule
// Nothing here looks a travel time up: the search is not repeated.
#pragma once
#undef BFILE
#define BFILE 2
#include "common.cuh"
#include "eval.cuh"

#define PLACE_THREADS 256

// feasible insertions of lanes [0, lane) of chunk c (one byte per lane, 64 per chunk)
__device__ __forceinline__ int lanes_before(const uint8_t* __restrict__ lane_cnt, int c, int lane) {
    const unsigned long long* w = (const unsigned long long*)(lane_cnt + (i64)c * 64);
    int sum = 0;
    for (int x = 0; x * 8 < lane; x++) {
        unsigned long long v = w[x];
        const int keep = lane - x * 8;                 // bytes of this word that count
        if (keep < 8) v &= (1ull << (8 * keep)) - 1ull;
        // horizontal byte sum
        v = (v & 0x00ff00ff00ff00ffull) + ((v >> 8) & 0x00ff00ff00ff00ffull);
        v = (v & 0x0000ffff0000ffffull) + ((v >> 16) & 0x0000ffff0000ffffull);
        sum += (int)((v & 0xffffffffull) + (v >> 32));
    }
    return sum;
}

// one block per record region (= per block of the search launch)
__global__ void __launch_bounds__(PLACE_THREADS)
k_place_cost(const Control* __restrict__ ctl, const Tuple* __restrict__ tup, const int* __restrict__ tup_cnt, int region_cap,
             const uint8_t* __restrict__ lane_cnt, const i64* __restrict__ chunk_pos, Level cur, unsigned* __restrict__ tup_pos) {
    pdl_enter();
    if (ctl->overflow) return;
    const int n = tup_cnt[blockIdx.x];
    const i64 base = (i64)blockIdx.x * region_cap;
    for (int s = threadIdx.x; s < n; s += PLACE_THREADS) {
        const Tuple t = tup[base + s];
        const i64 pos = chunk_pos[t.c] + lanes_before(lane_cnt, t.c, TUP_LANE(t.info)) + TUP_ORD(t.info);
        if (BOK(pos >= 0 && pos < cur.cap_sch)) cur.sch_cost[pos] = t.cost;
        tup_pos[base + s] = (unsigned)pos;
    }
}

#define PLACE_LANES 8      // lanes that build one schedule row together
__global__ void __launch_bounds__(PLACE_THREADS)
k_place_rows(const DEpoch* __restrict__ epp, const Control* __restrict__ ctl, const Tuple* __restrict__ tup, const int* __restrict__ tup_cnt,
             int region_cap, const unsigned* __restrict__ tup_pos, const RowDesc* __restrict__ rows, int G, Level prev, Level cur) {
    pdl_enter();
    if (ctl->overflow) return;
    const DEpoch& ep = *epp;
    const int n = tup_cnt[blockIdx.x];
    const i64 base = (i64)blockIdx.x * region_cap;
    const int sub = threadIdx.x & (PLACE_LANES - 1), grp = threadIdx.x / PLACE_LANES;
    for (int s = grp; s < n; s += PLACE_THREADS / PLACE_LANES) {
        const Tuple t = tup[base + s];
        const RowDesc rd = rows[(i64)t.c * G + TUP_ROW(t.info)];
        const int i = TUP_I(t.info), j = TUP_J(t.info), l = rd.l, q = rd.q;
        const Stop* __restrict__ src = prev.sch + (rd.s0 + rd.s) * prev.stride;
        Stop* __restrict__ dst = cur.sch + (i64)cur.final[tup_pos[base + s]] * cur.stride;
        // new schedule: base[0..i) P base[i..j-1) D base[j-1..l)   (insert_req_into_sche, scheduling.py:56-57)
        for (int w = sub; w < l + 2; w += PLACE_LANES) {
            Stop o;
            if (w == i) { o.code = 2 * q; o.node = ep.onid[q]; o.ddl = ep.clp[q]; }
            else if (w == j) { o.code = 2 * q + 1; o.node = ep.dnid[q]; o.ddl = ep.cld[q]; }
            else o = src[w - (w > i ? 1 : 0) - (w > j ? 1 : 0)];
            if (BOK(w < cur.stride && cur.final[tup_pos[base + s]] >= 0 && (i64)cur.final[tup_pos[base + s]] < cur.cap_sch)) dst[w] = o;
        }
    }
}
