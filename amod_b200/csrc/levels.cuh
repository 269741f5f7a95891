// levels.cuh — everything around the insertion search: basic schedules (dispatch_osp.py:295-322),
// reachability screen (:160 / dispatch_sba.py:59), ordered compaction of feasible tasks into the
// next trip level, and candidate generation for trips of size k>=2 (:167-216).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// Level 0: basic schedules.  One warp per vehicle.
//   idle / rebalancing / no-reoptimisation: [veh.sche]                       (dispatch_osp.py:301-303)
//   working:  onboard drop-offs in route order, then every other permutation (itertools order)
//             that passes test_constraints_get_cost(..., 0, 0, 0, T)         (:308-320)
// MODE 0 counts, MODE 1 writes the code rows.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void unrank_perm(int p, int L, int* out /*positions*/) {
    // p-th permutation of 0..L-1 in lexicographic order (= itertools.permutations order)
    int fact[AMOD_MAX_CAP + 1];
    fact[0] = 1;
    for (int a = 1; a <= L; a++) fact[a] = fact[a - 1] * a;
    unsigned used = 0;
    for (int pos = 0; pos < L; pos++) {
        int f = fact[L - 1 - pos];
        int d = p / f; p -= d * f;
        int c = 0;
        for (int x = 0; x < L; x++) {
            if (used & (1u << x)) continue;
            if (c == d) { out[pos] = x; used |= 1u << x; break; }
            c++;
        }
    }
}

template <typename TT, int MODE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_basic(DEpoch ep, Table<TT> tt, int* __restrict__ len0, int* __restrict__ S0, int* __restrict__ words0,
        const i64* __restrict__ sch_off, code_t* __restrict__ sch, Stats* stats) {
    const int lane = threadIdx.x & 31;
    const int v = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (v >= ep.n_veh) return;
    const int s0 = ep.veh_sche_off[v], s1 = ep.veh_sche_off[v + 1];
    const int st = ep.veh_status[v];
    const bool perms = (ep.mode == AMOD_MODE_OSP) && st == 2;
    int base[AMOD_MAX_STOPS];
    int L = 0;
    for (int k = s0; k < s1 && L < AMOD_MAX_STOPS; k++)
        if (!perms || ep.sche_onboard[k]) base[L++] = ep.sche_code[k];
    code_t* out = MODE == 1 ? sch + sch_off[v] : nullptr;
    if (MODE == 1 && lane == 0) for (int m = 0; m < L; m++) out[m] = (code_t)base[m];
    int count = 1;
    unsigned long long evals = 0, lookups = 0;
    if (perms && L >= 2 && L <= AMOD_MAX_CAP) {
        int nperm = 1;
        for (int a = 2; a <= L; a++) nperm *= a;
        const double T = ep.T;
        for (int p0 = 1; p0 < nperm; p0 += 32) {
            const int p = p0 + lane;
            bool ok = false;
            int pos[AMOD_MAX_CAP];
            if (p < nperm) {
                unrank_perm(p, L, pos);
                evals++;
                int n = ep.veh_load[v];
                bool capok = true;
                for (int m = 0; m < L; m++) {
                    int node, pod; double ddl;
                    stop_info(ep, v, base[pos[m]], node, ddl, pod);
                    n += pod;
                    if (n > ep.cap) { capok = false; break; }
                }
                if (capok) {
                    ok = true;
                    double acc = ep.veh_t[v];
                    int nid = ep.veh_nid[v];
                    for (int m = 0; m < L; m++) {
                        int node, pod; double ddl;
                        stop_info(ep, v, base[pos[m]], node, ddl, pod);
                        acc += tt(nid, node);
                        lookups++;
                        if (T + acc > ddl) { ok = false; break; }
                        nid = node;
                    }
                }
            }
            const unsigned bal = __ballot_sync(FULL, ok);
            if (MODE == 1 && ok) {
                code_t* dst = out + (i64)(count + __popc(bal & ((1u << lane) - 1u))) * L;
                for (int m = 0; m < L; m++) dst[m] = (code_t)base[pos[m]];
            }
            count += __popc(bal);
        }
    }
    if (MODE == 0) {
        if (lane == 0) { len0[v] = L; S0[v] = count; words0[v] = count * L; }
        evals = warp_sum(evals); lookups = warp_sum(lookups);
        if (lane == 0 && evals) { atomicAdd(&stats->n_evals, evals); atomicAdd(&stats->n_lookups, lookups); }
        if (lane == 0 && (s1 - s0 > AMOD_MAX_STOPS || (perms && L > AMOD_MAX_CAP))) atomicOr(&stats->flags, 1ull);
    }
}

// level-0 trip records: trip v = vehicle v
__global__ void k_level0_finish(int n_veh, const int* __restrict__ S0, int* trip_veh, int* nfeas, int* rot, int* veh_start) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n_veh) { trip_veh[v] = v; nfeas[v] = S0[v]; rot[v] = 0; veh_start[v] = v; }
    if (v == n_veh) veh_start[v] = n_veh;
}

// ---------------------------------------------------------------------------------------------
// Screen: tt[veh.nid][req.onid] + veh.t_to_nid + T > req.Clp -> skip (dispatch_osp.py:160,
// dispatch_sba.py:59; note the association differs from the in-schedule check).
// Flat order = task order: vehicle-major for OSP, request-major for SBA.
// ---------------------------------------------------------------------------------------------
template <typename TT>
__global__ void k_screen(DEpoch ep, Table<TT> tt, uint8_t* __restrict__ flag, i64 n) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int v, s;
    if (ep.mode == AMOD_MODE_SBA) { s = (int)(e / ep.n_veh); v = (int)(e - (i64)s * ep.n_veh); }
    else { v = (int)(e / ep.n_search); s = (int)(e - (i64)v * ep.n_search); }
    const int q = ep.search[s];
    const double d = tt(ep.veh_nid[v], ep.onid[q]);
    flag[e] = !(d + ep.veh_t[v] + ep.T > ep.clp[q]);
}

__global__ void k_tasks_from_screen(DEpoch ep, const uint8_t* __restrict__ flag, const i64* __restrict__ pos, i64 n,
                                    int* task_veh, int* task_base, int* task_q) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n || !flag[e]) return;
    int v, s;
    if (ep.mode == AMOD_MODE_SBA) { s = (int)(e / ep.n_veh); v = (int)(e - (i64)s * ep.n_veh); }
    else { v = (int)(e / ep.n_search); s = (int)(e - (i64)v * ep.n_search); }
    const i64 t = pos[e];
    task_veh[t] = v; task_base[t] = v; task_q[t] = ep.search[s];
}

// ---------------------------------------------------------------------------------------------
// Feasible tasks -> trips of the new level, order preserved (dispatch_osp.py:163-164, :217-218).
// ---------------------------------------------------------------------------------------------
__global__ void k_task_flags(int n_tasks, const int* __restrict__ task_nfeas, const int* __restrict__ task_veh,
                             const int* __restrict__ len0, int k, uint8_t* flag, int* words, Stats* stats) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int nf = t < n_tasks ? task_nfeas[t] : 0;
    if (t < n_tasks) {
        flag[t] = nf > 0;
        words[t] = nf * (len0[task_veh[t]] + 2 * k);   // nf <= ~1e6, len <= 40: fits int32 per task
    }
    int tot = warp_sum(nf);
    if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&stats->n_sches, (unsigned long long)tot);
}

__global__ void k_compact_trips(int n_tasks, const int* __restrict__ task_nfeas, const double* __restrict__ task_cost,
                                const int* __restrict__ task_veh, const int* __restrict__ task_base,
                                const int* __restrict__ task_q, const i64* __restrict__ trip_pos,
                                const i64* __restrict__ word_pos, Level prev, Level cur) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks || task_nfeas[t] == 0) return;
    const int o = (int)trip_pos[t];
    cur.trip_veh[o] = task_veh[t];
    cur.nfeas[o] = task_nfeas[t];
    cur.cost[o] = task_cost[t];
    cur.sch_off[o] = word_pos[t];
    // trip = sorted(base trip + q)  (dispatch_osp.py:187)
    const int kb = prev.k, q = task_q[t];
    const int* b = prev.trip_req + (i64)task_base[t] * kb;
    int* d = cur.trip_req + (i64)o * (kb + 1);
    int w = 0; bool put = false;
    for (int m = 0; m < kb; m++) { if (!put && q < b[m]) { d[w++] = q; put = true; } d[w++] = b[m]; }
    if (!put) d[w++] = q;
}

// veh_start[v] = first trip of vehicle v (trips are vehicle-major)
__global__ void k_veh_start(int n_trips, int n_veh, const int* __restrict__ trip_veh, int* veh_start) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_trips) return;
    int vprev = t == 0 ? -1 : trip_veh[t - 1];
    int vcur = t == n_trips ? n_veh : trip_veh[t];
    for (int u = vprev + 1; u <= vcur; u++) veh_start[u] = t;
}

// ---------------------------------------------------------------------------------------------
// Trip lookup table of one level: (vehicle, ascending request tuple) -> trip index.
// Replaces the reference's `list in list` scans (dispatch_osp.py:194,202).
// ---------------------------------------------------------------------------------------------
struct TripHash { int* slots; unsigned mask; };

__device__ __forceinline__ unsigned hash_trip(int v, const int* r, int k, int skip, int extra) {
    // hash of the ascending tuple (r minus element `skip`, plus `extra` merged in order)
    unsigned long long h = 0x9E3779B97F4A7C15ull ^ (unsigned long long)(unsigned)v * 0xD6E8FEB86659FD93ull;
    bool put = extra < 0;
    for (int m = 0; m < k; m++) {
        if (m == skip) continue;
        if (!put && extra < r[m]) { h = (h ^ (unsigned)extra) * 0xFF51AFD7ED558CCDull; h ^= h >> 32; put = true; }
        h = (h ^ (unsigned)r[m]) * 0xFF51AFD7ED558CCDull; h ^= h >> 32;
    }
    if (!put) { h = (h ^ (unsigned)extra) * 0xFF51AFD7ED558CCDull; h ^= h >> 32; }
    return (unsigned)(h ^ (h >> 29));
}

__global__ void k_hash_build(Level lv, TripHash th) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= lv.n_trips) return;
    unsigned h = hash_trip(lv.trip_veh[t], lv.trip_req + (i64)t * lv.k, lv.k, -1, -1) & th.mask;
    while (atomicCAS(&th.slots[h], -1, t) != -1) h = (h + 1) & th.mask;
}

// index of the trip (v, r minus r[skip] plus extra), or -1
__device__ __forceinline__ int hash_find(const Level& lv, const TripHash& th, int v, const int* r, int skip, int extra) {
    const int k = lv.k;
    unsigned h = hash_trip(v, r, k, skip, extra) & th.mask;
    for (;;) {
        const int t = th.slots[h];
        if (t < 0) return -1;
        if (lv.trip_veh[t] == v) {
            const int* c = lv.trip_req + (i64)t * k;
            bool same = true, put = false;
            int w = 0;
            for (int m = 0; m < k && same; m++) {
                if (m == skip) continue;
                if (!put && extra < r[m]) { same = c[w++] == extra; put = true; if (!same) break; }
                same = c[w++] == r[m];
            }
            if (same && !put) same = c[w++] == extra;
            if (same) return t;
        }
        h = (h + 1) & th.mask;
    }
}

// ---------------------------------------------------------------------------------------------
// Candidates of size k+1 from level k (dispatch_osp.py:183-216) in closed form: a (k+1)-set U is
// searched iff all its k-subsets are feasible; it is searched from its lowest-index subset
// (trip_i), inserting the one request U - trip_i, and the searches of one vehicle run in
// (i, j) order, j = second-lowest subset index.  One warp per row (= trip_i); lanes scan the
// vehicle's size-1 trips for the request to add.
// MODE 0: count per row.  MODE 1: write (j, q) at row_off (ranked by j afterwards).
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_gen(Level l1, Level prev, TripHash th, int* __restrict__ row_cnt, const i64* __restrict__ row_off,
      int* __restrict__ tmp_j, int* __restrict__ tmp_q, int* __restrict__ tmp_row) {
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (t >= prev.n_trips) return;
    const int v = prev.trip_veh[t], k = prev.k;
    const int a0 = l1.veh_start[v], a1 = l1.veh_start[v + 1];
    const int* mine = prev.trip_req + (i64)t * k;
    int cnt = 0;
    const i64 off = MODE == 1 ? row_off[t] : 0;
    for (int c0 = a0; c0 < a1; c0 += 32) {
        const int c = c0 + lane;
        int j = -1, r = -1;
        if (c < a1) {
            r = l1.trip_req[c];
            if (k == 1) {
                if (c > t) j = c;                       // all pairs i<j (dispatch_osp.py:183-186)
            } else {
                bool in = false;
                for (int m = 0; m < k; m++) in |= mine[m] == r;
                if (!in) {
                    j = 0x7fffffff;
                    for (int d = 0; d < k; d++) {       // every other k-subset must exist with a larger index
                        const int u = hash_find(prev, th, v, mine, d, r);
                        if (u <= t) { j = -1; break; }
                        j = min(j, u);
                    }
                }
            }
        }
        const unsigned bal = __ballot_sync(FULL, j >= 0);
        if (MODE == 1 && j >= 0) {
            const i64 o = off + cnt + __popc(bal & ((1u << lane) - 1u));
            tmp_j[o] = j; tmp_q[o] = r; tmp_row[o] = t;
        }
        cnt += __popc(bal);
    }
    if (MODE == 0 && lane == 0) row_cnt[t] = cnt;
}

// rank the candidates of each row by j and emit the tasks
__global__ void k_gen_rank(i64 n, const i64* __restrict__ row_off, const int* __restrict__ tmp_j,
                           const int* __restrict__ tmp_q, const int* __restrict__ tmp_row, const int* __restrict__ trip_veh,
                           int sorted_already, int* task_veh, int* task_base, int* task_q) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int row = tmp_row[e];
    i64 o = e;
    if (!sorted_already) {
        const i64 b = row_off[row], en = row_off[row + 1];
        const int j = tmp_j[e];
        int rank = 0;
        for (i64 x = b; x < en; x++) rank += tmp_j[x] < j;
        o = b + rank;
    }
    task_veh[o] = trip_veh[row]; task_base[o] = row; task_q[o] = tmp_q[e];
}
