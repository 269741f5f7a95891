// levels.cuh — everything around the insertion search: basic schedules (dispatch_osp.py:295-322),
// reachability screen (:160 / dispatch_sba.py:59), ordered compaction of feasible tasks into the
// next trip level, list-order classification (scheduling.py:31-36) and candidate generation for
// trips of size k>=2 (:167-216).  All sizes come from the device-resident Control block.
#pragma once
#undef BFILE
#define BFILE 3
#include "common.cuh"
#include "eval.cuh"
#include "scan.cuh"

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

// The searched requests' origin and latest pick-up time gathered once per epoch into arrays indexed by search position, so that
// the reachability screen (one lookup per vehicle x request) reads them coalesced instead of through search[] -> onid[] / clp[].
struct SearchCache { int* onid; double* clp; };
__device__ __forceinline__ void fill_search_cache(const DEpoch& ep, const SearchCache& sc) {
    const int stride = gridDim.x * blockDim.x;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < ep.n_search; s += stride) {
        const int q = ep.search[s];
        sc.onid[s] = ep.onid[q]; sc.clp[s] = ep.clp[q];
    }
}

template <typename TT, int MODE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_basic(const DEpoch* __restrict__ epp, Table<TT> tt, Level lv0, int* __restrict__ len0, int* __restrict__ S0, const i64* __restrict__ s0pos,
        const Control* __restrict__ ctl, Stats* stats, unsigned long long* __restrict__ part, int part_vg, SearchCache sc) {
    pdl_enter();
    const DEpoch& ep = *epp;
    const int lane = threadIdx.x & 31;
    if (MODE == 0) fill_search_cache(ep, sc);
    if (MODE == 1 && ctl->overflow) return;          // level-0 store overflowed
    const int seg_veh = (int)scan_seg_len(ep.n_veh, part_vg);     // MODE 0 leaves the block sums of the scan over S0
    unsigned long long evals = 0, lookups = 0;
    bool bad = false;
    for (int v = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); v < ep.n_veh; v += gridDim.x * WARPS_PER_BLOCK) {
    const int s0 = ep.veh_sche_off[v], s1 = ep.veh_sche_off[v + 1];
    const int st = ep.veh_status[v];
    const bool perms = (ep.mode == AMOD_MODE_OSP) && st == 2;
    int base[AMOD_MAX_STOPS];
    int L = 0;
    for (int k = s0; k < s1 && L < AMOD_MAX_STOPS; k++)
        if (!perms || ep.sche_onboard[k]) base[L++] = ep.sche_code[k];
    const bool too_long = L > lv0.stride;
    if (too_long) L = 0;
    const i64 row0 = MODE == 1 ? s0pos[v] : 0;
    if (MODE == 1 && lane == 0) {
        Stop* out = lv0.sch + row0 * lv0.stride;
        for (int m = 0; m < L; m++) if (BOK(row0 < lv0.cap_sch && m < lv0.stride)) out[m] = make_stop(ep, v, base[m]);
        lv0.trip_veh[v] = v; lv0.nfeas[v] = S0[v]; lv0.s0[v] = row0; lv0.veh_start[v] = v;
        if (v == 0) lv0.veh_start[ep.n_veh] = ep.n_veh;
    }
    int count = 1;
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
                const int r = count + __popc(bal & ((1u << lane) - 1u));
                Stop* dst = lv0.sch + (row0 + r) * lv0.stride;       // itertools order = list order (dispatch_osp.py:315-320)
                for (int m = 0; m < L; m++) if (BOK(row0 + r < lv0.cap_sch && m < lv0.stride)) dst[m] = make_stop(ep, v, base[pos[m]]);
            }
            count += __popc(bal);
        }
    }
    if (MODE == 0) {
        if (lane == 0) { len0[v] = L; S0[v] = count; if (count) atomicAdd(&part[v / seg_veh], (unsigned long long)count); }
        bad |= too_long || s1 - s0 > AMOD_MAX_STOPS || (perms && L > AMOD_MAX_CAP);
    }
    }
    if (MODE == 0) {
        evals = warp_sum(evals); lookups = warp_sum(lookups);
        if (lane == 0 && evals) { atomicAdd(&stats->n_evals, evals); atomicAdd(&stats->n_lookups, lookups); }
        if (lane == 0 && bad) atomicOr(&stats->flags, 1ull);
    }
}

// ---------------------------------------------------------------------------------------------
// Level 0 for large capacities: the permutation search spread over the whole GPU.  A loaded vehicle has up to
// cap! orderings of its drop-offs (40,320 at capacity 8, dispatch_osp.py:315); one warp per vehicle would
// serialise them.  Work item = 32 consecutive permutation ranks of one vehicle; count -> prefix sum -> fill
// keeps the itertools order.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int factorial(int L) { int f = 1; for (int a = 2; a <= L; a++) f *= a; return f; }

// test_constraints_get_cost(veh_params, perm, 0, 0, 0, T) (dispatch_osp.py:318): flag only; counts the lookups done
template <typename TT>
__device__ __forceinline__ bool perm_feasible(const DEpoch& ep, const Table<TT>& tt, int v, const int* base, int L, const int* pos,
                                              unsigned long long& lookups) {
    int n = ep.veh_load[v];
    for (int m = 0; m < L; m++) {
        int node, pod; double ddl;
        stop_info(ep, v, base[pos[m]], node, ddl, pod);
        n += pod;
        if (n > ep.cap) return false;
    }
    double acc = ep.veh_t[v];
    int nid = ep.veh_nid[v];
    for (int m = 0; m < L; m++) {
        int node, pod; double ddl;
        stop_info(ep, v, base[pos[m]], node, ddl, pod);
        acc += tt(nid, node);
        lookups++;
        if (ep.T + acc > ddl) return false;
        nid = node;
    }
    return true;
}

__global__ void k_perm_prep(const DEpoch* __restrict__ epp, Level lv0, int* __restrict__ len0, int* __restrict__ perm_base,
                            int* __restrict__ n_items, Stats* stats, SearchCache sc) {
    pdl_enter();
    const DEpoch& ep = *epp;
    fill_search_cache(ep, sc);
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < ep.n_veh; v += gridDim.x * blockDim.x) {
        const int s0 = ep.veh_sche_off[v], s1 = ep.veh_sche_off[v + 1];
        const bool perms = (ep.mode == AMOD_MODE_OSP) && ep.veh_status[v] == 2;
        int L = 0;
        for (int k = s0; k < s1; k++)
            if (!perms || ep.sche_onboard[k]) { if (perms && L < AMOD_MAX_CAP) perm_base[v * AMOD_MAX_CAP + L] = ep.sche_code[k]; L++; }
        bool bad = L > lv0.stride || s1 - s0 > AMOD_MAX_STOPS || (perms && L > AMOD_MAX_CAP);
        if (bad) { atomicOr(&stats->flags, 1ull); L = 0; }
        len0[v] = L;
        n_items[v] = (perms && L >= 2) ? (factorial(L) - 1 + 31) / 32 : 0;
    }
}

struct FinPermItems {
    Control* ctl; i64 cap;
    __device__ __forceinline__ void operator()(i64 tot) const {
        if (tot > ctl->need_perm_items) ctl->need_perm_items = tot;
        if (tot > cap) { atomicOr(&ctl->overflow, OVF_ITEMS); tot = 0; }
        ctl->n_perm_items = (int)tot;
    }
};
struct FinBasicRows {       // level-0 rows = one identity row per vehicle + the feasible permutations
    Control* ctl; const DEpoch* ep; i64 cap;
    __device__ __forceinline__ void operator()(i64 tot) const {
        tot += ep->n_veh;
        if (tot > ctl->need_sch[0]) ctl->need_sch[0] = tot;
        if (tot > cap) { atomicOr(&ctl->overflow, OVF_SCH); tot = 0; }
        ctl->n_sch[0] = tot;
    }
};

template <typename TT, int MODE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_perm_items(const DEpoch* __restrict__ epp, Table<TT> tt, Level lv0, const Control* __restrict__ ctl, const int* __restrict__ len0,
             const int* __restrict__ perm_base, const i64* __restrict__ item_off, int* __restrict__ item_cnt,
             const i64* __restrict__ item_pos, Stats* stats) {
    pdl_enter();
    const DEpoch& ep = *epp;
    const int lane = threadIdx.x & 31;
    const int n = ctl->overflow ? 0 : ctl->n_perm_items;
    const int nwarps = gridDim.x * WARPS_PER_BLOCK;
    unsigned long long evals = 0, lookups = 0;
    for (int it = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); it < n; it += nwarps) {
        int lo = 0, hi = ep.n_veh;                      // vehicle of the item: last v with item_off[v] <= it
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (item_off[mid] <= it) lo = mid; else hi = mid; }
        const int v = lo, L = len0[v];
        const int* base = perm_base + v * AMOD_MAX_CAP;
        const int p = 1 + (int)(it - item_off[v]) * 32 + lane;
        bool ok = false;
        int pos[AMOD_MAX_CAP];
        if (p < factorial(L)) {
            unrank_perm(p, L, pos);
            evals++;
            ok = perm_feasible(ep, tt, v, base, L, pos, lookups);
        }
        const unsigned bal = __ballot_sync(FULL, ok);
        if (MODE == 0) { if (lane == 0) item_cnt[it] = __popc(bal); }
        else if (ok) {
            const i64 row = v + 1 + item_pos[it] + __popc(bal & ((1u << lane) - 1u));      // (row v + item_pos[item_off[v]] is the vehicle's identity row)
            Stop* dst = lv0.sch + row * lv0.stride;
            for (int m = 0; m < L; m++) dst[m] = make_stop(ep, v, base[pos[m]]);
        }
    }
    if (MODE == 0) {
        evals = warp_sum(evals); lookups = warp_sum(lookups);
        if (lane == 0 && evals) { atomicAdd(&stats->n_evals, evals); atomicAdd(&stats->n_lookups, lookups); }
    }
}

// identity rows and the level-0 trip records
__global__ void k_perm_rows(const DEpoch* __restrict__ epp, Level lv0, const Control* __restrict__ ctl, const int* __restrict__ len0,
                            int* __restrict__ S0, const i64* __restrict__ item_off, const i64* __restrict__ item_pos) {
    pdl_enter();
    const DEpoch& ep = *epp;
    if (ctl->overflow) return;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < ep.n_veh; v += gridDim.x * blockDim.x) {
        const i64 row0 = v + item_pos[item_off[v]];
        const int nf = 1 + (int)(item_pos[item_off[v + 1]] - item_pos[item_off[v]]);
        const bool perms = (ep.mode == AMOD_MODE_OSP) && ep.veh_status[v] == 2;
        Stop* out = lv0.sch + row0 * lv0.stride;
        int L = 0;
        for (int k = ep.veh_sche_off[v]; k < ep.veh_sche_off[v + 1] && L < len0[v]; k++)
            if (!perms || ep.sche_onboard[k]) out[L++] = make_stop(ep, v, ep.sche_code[k]);
        lv0.trip_veh[v] = v; lv0.nfeas[v] = nf; lv0.s0[v] = row0; lv0.veh_start[v] = v;
        S0[v] = nf;
        if (v == 0) lv0.veh_start[ep.n_veh] = ep.n_veh;
    }
}

// ---------------------------------------------------------------------------------------------
// Screen: tt[veh.nid][req.onid] + veh.t_to_nid + T > req.Clp -> skip (dispatch_osp.py:160,
// dispatch_sba.py:59; note the association differs from the in-schedule check).
// Task order is the reference's loop order: vehicle-major for OSP, request-major for SBA.  One warp per
// "major" (vehicle resp. request) walks the minors 32 at a time: pass 1 keeps a survivor bitmask and the
// (task, row) counts of the major, a prefix sum over the majors gives its offsets, pass 2 expands the
// bitmask into tasks in order.  (No V x Q intermediate except the bitmask.)
// ---------------------------------------------------------------------------------------------
// Packed two-component scans: low 32 bits count tasks, high 32 bits count their rows (base schedules),
// so one prefix sum yields both the task index and the task's first row (both totals < 2^31).
#define PACK_LO(x) ((i64)((x) & 0xffffffffll))
#define PACK_HI(x) ((i64)((x) >> 32))

#define SCREEN_WARPS 4      // warps sharing one major (vehicle / request)

template <typename TT>
__global__ void __launch_bounds__(SCREEN_WARPS * 32)
k_screen_count(const DEpoch* __restrict__ epp, Table<TT> tt, const int* __restrict__ S0, unsigned* __restrict__ bits,
               i64* __restrict__ major_cnt, unsigned long long* __restrict__ part, int part_vg, SearchCache sc) {
    pdl_enter();
    __shared__ int s_cnt[SCREEN_WARPS], s_rows[SCREEN_WARPS];
    const DEpoch& ep = *epp;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const bool sba = ep.mode == AMOD_MODE_SBA;
    const int n_major = sba ? ep.n_search : ep.n_veh, n_minor = sba ? ep.n_veh : ep.n_search;
    const int wpm = (n_minor + 31) >> 5;
    const int seg_major = (int)scan_seg_len(n_major, part_vg);      // block sums of the scan over the majors, left as we go
    for (int mj = blockIdx.x; mj < n_major; mj += gridDim.x) {      // one block per major, its warps interleave the words
        int cnt = 0, rows = 0;
        const int vq = sba ? ep.search[mj] : 0;
        const int m_node = sba ? ep.onid[vq] : ep.veh_nid[mj];
        const double m_a = sba ? ep.clp[vq] : ep.veh_t[mj];
        for (int w0 = wib; w0 < wpm; w0 += SCREEN_WARPS * 4) {     // four words per round: their lookups are in flight together
            bool ok[4];
            int S[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int mn = (w0 + u * SCREEN_WARPS) * 32 + lane;
                ok[u] = false; S[u] = 0;
                if (mn < n_minor) {
                    if (sba) {   // major = request, minor = vehicle
                        const double d = tt(ep.veh_nid[mn], m_node);
                        ok[u] = !(d + ep.veh_t[mn] + ep.T > m_a);
                        S[u] = S0[mn];
                    } else {     // major = vehicle, minor = searched request
                        const double d = tt(m_node, sc.onid[mn]);
                        ok[u] = !(d + m_a + ep.T > sc.clp[mn]);
                        S[u] = S0[mj];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int w = w0 + u * SCREEN_WARPS;
                const unsigned bal = __ballot_sync(FULL, ok[u]);
                if (lane == 0 && w < wpm) bits[(i64)mj * wpm + w] = bal;
                cnt += __popc(bal);
                rows += ok[u] ? S[u] : 0;
            }
        }
        rows = warp_sum(rows);
        if (lane == 0) { s_cnt[wib] = cnt; s_rows[wib] = rows; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int c = 0, r = 0;
            for (int x = 0; x < SCREEN_WARPS; x++) { c += s_cnt[x]; r += s_rows[x]; }
            major_cnt[mj] = (i64)c | ((i64)r << 32);
            if (c) atomicAdd(&part[mj / seg_major], (unsigned long long)c | ((unsigned long long)r << 32));
        }
        __syncthreads();
    }
}

struct InMajorPacked {       // scan input over the majors of the screen
    const i64* p; const DEpoch* ep;
    __device__ __forceinline__ i64 size() const { return ep->mode == AMOD_MODE_SBA ? ep->n_search : ep->n_veh; }
    __device__ __forceinline__ i64 operator()(i64 i) const { return p[i]; }
};
// finisher: task count of buffer `which`, row count and chunk count (G rows per chunk) of the next level
struct FinTasksRows {
    Control* ctl; int which; i64 cap_tasks, cap_rows, cap_chunks; int G;
    __device__ __forceinline__ void operator()(i64 packed) const {
        i64 nt = PACK_LO(packed), nr = PACK_HI(packed);
        if (packed < 0 || nr > 0x7ffffff0ll) nr = 0x7fffffffll;     // the packed row sum left 31 bits: reported as need_rows, the host fails with AMOD_E_LIMIT
        i64 nch = (nr + G - 1) / G;
        if (nt > ctl->need_tasks) ctl->need_tasks = nt;
        if (nr > ctl->need_rows) ctl->need_rows = nr;
        if (nch > ctl->need_chunks) ctl->need_chunks = nch;
        if (nt > cap_tasks) { atomicOr(&ctl->overflow, OVF_TASKS); nt = 0; }
        if (nr > cap_rows) { atomicOr(&ctl->overflow, OVF_ROWS); nt = 0; }
        if (nch > cap_chunks) { atomicOr(&ctl->overflow, OVF_CHUNKS); nt = 0; }
        if (nt == 0) { nr = 0; nch = 0; }
        ctl->n_tasks[which] = (int)nt; ctl->n_rows[which] = (int)nr; ctl->n_chunks[which] = (int)nch;
        ctl->stats.n_chunks += (unsigned long long)nch;
    }
};
struct InVehInt {            // scan input over a per-vehicle int array
    const int* p; const DEpoch* ep;
    __device__ __forceinline__ i64 size() const { return ep->n_veh; }
    __device__ __forceinline__ i64 operator()(i64 i) const { return (i64)p[i]; }
};

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_screen_fill(const DEpoch* __restrict__ epp, const Control* __restrict__ ctl, const int* __restrict__ S0,
              const unsigned* __restrict__ bits, const i64* __restrict__ major_off, int* task_veh, int* task_base, int* task_q,
              int* task_S, i64* task_row0) {
    pdl_enter();
    const DEpoch& ep = *epp;
    if (ctl->n_tasks[0] == 0 || ctl->overflow) return;
    const int lane = threadIdx.x & 31;
    const bool sba = ep.mode == AMOD_MODE_SBA;
    const int n_major = sba ? ep.n_search : ep.n_veh, n_minor = sba ? ep.n_veh : ep.n_search;
    const int wpm = (n_minor + 31) >> 5;
    const int nwarps = gridDim.x * WARPS_PER_BLOCK;
    for (int mj = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); mj < n_major; mj += nwarps) {
        i64 t0 = PACK_LO(major_off[mj]), r0 = PACK_HI(major_off[mj]);
        if (PACK_LO(major_off[mj + 1]) == t0) continue;
        // a lane per 32-pair word (the masks are sparse): count, scan over the lanes, then each lane writes its own tasks
        for (int w0 = 0; w0 < wpm; w0 += 32) {
            const int w = w0 + lane;
            unsigned word = w < wpm ? bits[(i64)mj * wpm + w] : 0u;
            const int S_major = sba ? 0 : S0[mj];
            i64 mine = 0;
            for (unsigned x = word; x; x &= x - 1) {
                const int S = sba ? S0[w * 32 + __ffs(x) - 1] : S_major;
                mine += 1 | ((i64)S << 32);
            }
            i64 incl = mine;
            for (int o = 1; o < 32; o <<= 1) { const i64 y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
            i64 t = t0 + PACK_LO(incl - mine), r = r0 + PACK_HI(incl - mine);
            for (; word; word &= word - 1) {
                const int mn = w * 32 + __ffs(word) - 1;
                const int v = sba ? mn : mj;
                const int S = sba ? S0[v] : S_major;
                if (!BOK(t < ctl->n_tasks[0] && r + S <= ctl->n_rows[0])) { t++; r += S; continue; }
                task_veh[t] = v; task_base[t] = v; task_q[t] = ep.search[sba ? mj : mn];
                task_S[t] = S; task_row0[t] = r;
                t++; r += S;
            }
            const i64 tot = __shfl_sync(FULL, incl, 31);
            t0 += PACK_LO(tot); r0 += PACK_HI(tot);
        }
    }
}

// row descriptors: one warp per task, lanes over the task's base schedules
__global__ void __launch_bounds__(256)
k_expand_rows(Control* __restrict__ ctl, int which, Level prev, const int* __restrict__ len0, const int* __restrict__ task_veh,
              const int* __restrict__ task_base, const int* __restrict__ task_q, const int* __restrict__ task_S,
              const i64* __restrict__ task_row0, RowDesc* __restrict__ rows, int* __restrict__ task_nfeas) {
    pdl_enter();
    const int n = (ctl->n_rows[which] && !ctl->overflow) ? ctl->n_tasks[which] : 0;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&ctl->stats.n_tasks, (unsigned long long)n);   // compute_schedule calls
    unsigned long long items = 0;
    // a lane per task (most tasks have a handful of base schedules); tasks with many rows are shared by the warp
    for (int t0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; t0 < n; t0 += nwarps * 32) {
        const int t = t0 + lane;
        int v = 0, q = 0, l = 0, S = 0;
        i64 s0 = 0, r0 = 0;
        if (t < n) {
            v = task_veh[t]; q = task_q[t];
            l = len0[v] + 2 * prev.k;
            s0 = prev.s0[task_base[t]];
            r0 = task_row0[t]; S = task_S[t];
            task_nfeas[t] = 0;
            items += (unsigned long long)S * (l + 1);
        }
        const bool big = S > 64;
        if (!big)
            for (int x = 0; x < S; x++) {
                RowDesc d;
                d.v = v; d.q = q; d.s = x; d.l = l; d.s0 = s0; d.task = t; d.pad = 0;
                rows[r0 + x] = d;
            }
        unsigned todo = __ballot_sync(FULL, big);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int bv = __shfl_sync(FULL, v, src), bq = __shfl_sync(FULL, q, src), bl = __shfl_sync(FULL, l, src);
            const int bS = __shfl_sync(FULL, S, src);
            const i64 bs0 = __shfl_sync(FULL, s0, src), br0 = __shfl_sync(FULL, r0, src);
            for (int x = lane; x < bS; x += 32) {
                RowDesc d;
                d.v = bv; d.q = bq; d.s = x; d.l = bl; d.s0 = bs0; d.task = t0 + src; d.pad = 0;
                rows[br0 + x] = d;
            }
        }
    }
    items = warp_sum(items);
    if (lane == 0 && items) atomicAdd(&ctl->stats.n_items, items);
}

// ---------------------------------------------------------------------------------------------
// Trip lookup table of one level: (vehicle, ascending request tuple) -> trip index.
// Replaces the reference's `list in list` scans (dispatch_osp.py:194,202).
// Open addressing, 8 B slots: (32-bit fingerprint of the key) << 32 | trip index.  A lookup is a chain of dependent
// L2 round trips and the candidate kernels are nothing but such chains (a level has fewer trips than the GPU has
// warps), so the chain is kept short: a miss ends at the first slot (empty, or another fingerprint), a hit costs one
// more round trip (vehicle and the whole tuple are loaded together and compared without an early exit).
// ---------------------------------------------------------------------------------------------
struct TripHash { unsigned long long* slots; unsigned mask; unsigned fp_mask; };   // fp_mask: bits of the fingerprint kept (all; tests keep two)
#define HASH_EMPTY 0xffffffffffffffffull
// Tuples live in registers, so everything below is unrolled to a compile-time bound KMAX >= k; the kernels are instantiated for
// KMAX = 1, 4, 8 and AMOD_MAX_TRIP and the host picks the smallest that holds the level's trips (registers = resident warps).
#define KM_OF(k) ((k) <= 1 ? 1 : (k) <= 4 ? 4 : (k) <= 8 ? 8 : AMOD_MAX_TRIP)
#define KM_IDX(k) ((k) <= 1 ? 0 : (k) <= 4 ? 1 : (k) <= 8 ? 2 : 3)

template <int KMAX> struct ReqTuple { int r[KMAX]; };      // an ascending request tuple in registers (entries >= k are padding)

template <int KMAX> __device__ __forceinline__ ReqTuple<KMAX> load_tuple(const int* __restrict__ p, int k) {
    ReqTuple<KMAX> t;
#pragma unroll
    for (int m = 0; m < KMAX; m++) t.r[m] = m < k ? p[m] : 0x7fffffff;
    return t;
}

// the ascending k-tuple (a minus a[skip]) plus `extra` (skip < 0: nothing removed, the result has k+1 entries)
template <int KMAX> __device__ __forceinline__ ReqTuple<KMAX> merged_tuple(const ReqTuple<KMAX>& a, int k, int skip, int extra) {
    if (skip < 0) skip = KMAX;
    int pe = 0;                           // entries of (a minus skip) below extra
#pragma unroll
    for (int m = 0; m < KMAX; m++) pe += (m < k && m != skip && a.r[m] < extra) ? 1 : 0;
    ReqTuple<KMAX> o;
#pragma unroll
    for (int w = 0; w < KMAX; w++) {      // rr = a minus skip: rr[i] = i < skip ? a[i] : a[i+1]
        const int a_w = a.r[w];
        const int a_n = w + 1 < KMAX ? a.r[(w + 1) % KMAX] : 0x7fffffff;
        const int a_p = w > 0 ? a.r[(w + KMAX - 1) % KMAX] : 0;
        const int rr_w = w < skip ? a_w : a_n;            // rr[w]
        const int rr_p = w - 1 < skip ? a_p : a_w;        // rr[w-1]
        o.r[w] = w < pe ? rr_w : (w == pe ? extra : rr_p);
    }
    return o;
}

template <int KMAX> __device__ __forceinline__ unsigned long long hash_tuple(int v, const ReqTuple<KMAX>& t, int k) {
    unsigned long long h = 0x9E3779B97F4A7C15ull ^ (unsigned long long)(unsigned)v * 0xD6E8FEB86659FD93ull;
#pragma unroll
    for (int m = 0; m < KMAX; m++) if (m < k) { h = (h ^ (unsigned)t.r[m]) * 0xFF51AFD7ED558CCDull; h ^= h >> 32; }
    return h;
}
__device__ __forceinline__ unsigned hash_slot(unsigned long long h) { return (unsigned)(h ^ (h >> 29)); }
__device__ __forceinline__ unsigned hash_fp(unsigned long long h) { return (unsigned)(h >> 32); }

template <int KMAX> __device__ __forceinline__ void hash_insert(const TripHash& th, int v, const ReqTuple<KMAX>& key, int k, int trip) {
    const unsigned long long hv = hash_tuple(v, key, k);
    const unsigned long long val = ((unsigned long long)(hash_fp(hv) & th.fp_mask) << 32) | (unsigned)trip;
    unsigned h = hash_slot(hv) & th.mask;
    while (atomicCAS(&th.slots[h], HASH_EMPTY, val) != HASH_EMPTY) h = (h + 1) & th.mask;
}

// A lookup in two halves, so that a lane can have the first probes of several lookups in flight at once.
struct HashProbe { unsigned long long slot; unsigned h, fp; };
template <int KMAX> __device__ __forceinline__ HashProbe hash_begin(const TripHash& th, int v, const ReqTuple<KMAX>& want, int k) {
    const unsigned long long hv = hash_tuple(v, want, k);
    HashProbe p; p.h = hash_slot(hv) & th.mask; p.fp = hash_fp(hv) & th.fp_mask;
    p.slot = th.slots[p.h];
    return p;
}
// index of the trip (v, want), or -1
template <int KMAX> __device__ __forceinline__ int hash_finish(const Level& lv, const TripHash& th, int v, const ReqTuple<KMAX>& want, HashProbe p) {
    const int k = lv.k;
    for (;;) {
        if (p.slot == HASH_EMPTY) return -1;
        if ((unsigned)(p.slot >> 32) == p.fp) {
            const int t = (int)(unsigned)p.slot;
            const int* __restrict__ c = lv.trip_req + (i64)t * k;
            const int tv = lv.trip_veh[t];
            int cv[KMAX];
#pragma unroll
            for (int m = 0; m < KMAX; m++) cv[m] = m < k ? c[m] : 0;
            bool same = tv == v;
#pragma unroll
            for (int m = 0; m < KMAX; m++) same &= (m >= k) | (cv[m] == want.r[m]);
            if (same) return t;
        }
        p.h = (p.h + 1) & th.mask;
        p.slot = th.slots[p.h];
    }
}
template <int KMAX> __device__ __forceinline__ int hash_find(const Level& lv, const TripHash& th, int v, const ReqTuple<KMAX>& want) {
    return hash_finish(lv, th, v, want, hash_begin(th, v, want, lv.k));
}

// ---------------------------------------------------------------------------------------------
// Feasible tasks -> trips of the new level, order preserved (dispatch_osp.py:163-164, :217-218).
// ---------------------------------------------------------------------------------------------
struct InTaskFeasible {       // scan input: (the task produced a feasible schedule) | (its feasible schedules) << 32
    const Control* ctl; int which; const int* task_nfeas;
    __device__ __forceinline__ i64 size() const { return (ctl->n_chunks[which] && !ctl->overflow) ? ctl->n_tasks[which] : 0; }
    __device__ __forceinline__ i64 operator()(i64 t) const { const i64 nf = task_nfeas[t]; return (nf > 0 ? 1 : 0) | (nf << 32); }
};
// finisher of that scan: the trip count of the level (low half; the high half repeats the level's schedule count)
struct FinTrips {
    int* dst; i64 cap; i64* need; unsigned long long* ovf;
    __device__ __forceinline__ void operator()(i64 packed) const {
        i64 tot = PACK_LO(packed);
        if (tot > *need) *need = tot;
        if (tot > cap) { atomicOr(ovf, OVF_TRIPS); tot = 0; }
        *dst = (int)tot;
    }
};

template <int KMAX>
__global__ void k_compact_trips(const Control* __restrict__ ctl, int which, const int* __restrict__ task_veh,
                                const int* __restrict__ task_base, const int* __restrict__ task_q,
                                const int* __restrict__ task_nfeas,
                                const i64* __restrict__ trip_pos /* exclusive (trips | schedules << 32) before each task */, Level prev, Level cur,
                                TripHash th, Stats* stats, const DEpoch* __restrict__ epp, int sba,
                                const i64* __restrict__ major_off, const i64* __restrict__ task_first,
                                unsigned long long* __restrict__ gen_partial, int n_gen_partial) {
    pdl_enter();
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < n_gen_partial; x += gridDim.x * blockDim.x) gen_partial[x] = 0ull;   // segment sums of the candidate count that follows
    const int n = (ctl->n_trips[cur.k] && !ctl->overflow) ? ctl->n_tasks[which] : 0;
    const int stride = gridDim.x * blockDim.x;
    if (!sba && !ctl->overflow) {
        // veh_start[v] = first trip of vehicle v (trips are vehicle-major, like the tasks they come from): the
        // trip position of the vehicle's first task.  Level 1 tasks start at the screen's per-vehicle offsets;
        // level k tasks of v start where the candidates of v's first level-(k-1) trip start.
        const int n_veh = epp->n_veh;
        for (int v = blockIdx.x * blockDim.x + threadIdx.x; v <= n_veh; v += stride) {
            const i64 ft = cur.k == 1 ? PACK_LO(major_off[v]) : task_first[prev.veh_start[v]];
            cur.veh_start[v] = (int)PACK_LO(trip_pos[ft]);
        }
    }
    unsigned long long nsch = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        const int nf = task_nfeas[t];
        if (nf == 0) continue;
        nsch += (unsigned long long)nf;
        const i64 tp = trip_pos[t];
        const int o = (int)PACK_LO(tp);
        if (!BOK(o >= 0 && o < cur.cap_trips && PACK_HI(tp) + nf <= cur.cap_sch)) continue;
        cur.trip_veh[o] = task_veh[t];
        cur.nfeas[o] = nf;
        cur.s0[o] = PACK_HI(tp);            // rows are task-major: the trip's first schedule follows those of all earlier tasks
        // trip = sorted(base trip + q)  (dispatch_osp.py:187)
        const int kb = prev.k, q = task_q[t];
        const ReqTuple<KMAX> trip = merged_tuple(load_tuple<KMAX>(prev.trip_req + (i64)task_base[t] * kb, kb), kb, -1, q);
        int* d = cur.trip_req + (i64)o * (kb + 1);
#pragma unroll
        for (int m = 0; m < KMAX; m++) if (m <= kb) d[m] = trip.r[m];
        if (th.mask) hash_insert(th, task_veh[t], trip, kb + 1, o);   // the next level looks this one's trips up (table cleared two kernels back, by k_gen_pairs)
    }
    nsch = warp_sum(nsch);
    if ((threadIdx.x & 31) == 0 && nsch) atomicAdd(&stats->n_sches, nsch);
}

// ---------------------------------------------------------------------------------------------
// List order of feasible_sches (scheduling.py:31-36): a schedule whose cost is strictly below every
// earlier one is inserted at the front, the others are appended.  From the encounter-order costs:
// order = [records, newest first] + [non-records in encounter order]; order[0] is the best schedule
// (first-encountered minimum).  One warp per trip.
// ---------------------------------------------------------------------------------------------
// one trip by the whole warp (more than CLASSIFY_SMALL schedules)
__device__ __forceinline__ void classify_warp(const Level& cur, int t, i64 s0, int nf, int lane) {
    const double* __restrict__ cost = cur.sch_cost + s0;
    int* __restrict__ fin = cur.final + s0;       // encounter position -> row (list order)
    const int r0 = (int)s0;
    if (nf <= 32) {     // one warp step decides everything
        const bool in = lane < nf;
        const double c = in ? cost[lane] : DINF;
        const double before = warp_excl_min(c, lane);      // (outside the &&: every lane takes part in the shuffles)
        const bool isrec = in && c < before;
        const unsigned br = __ballot_sync(FULL, isrec), bn = __ballot_sync(FULL, in && !isrec);
        const unsigned lt = (1u << lane) - 1u;
        const int nrec1 = __popc(br);
        if (!BOK(s0 + nf <= cur.cap_sch)) return;
        if (isrec) fin[lane] = r0 + nrec1 - 1 - __popc(br & lt);
        else if (in) fin[lane] = r0 + nrec1 + __popc(bn & lt);
        const double best = warp_min(c);
        if (lane == 0) cur.cost[t] = best;
        return;
    }
    // pass 1: number of records
    double running = DINF;
    int nrec = 0;
    for (int b = 0; b < nf; b += 32) {
        const double c = b + lane < nf ? cost[b + lane] : DINF;
        const double incoming = fmin(running, warp_excl_min(c, lane));
        nrec += __popc(__ballot_sync(FULL, c < incoming));
        running = fmin(running, warp_min(c));
    }
    // pass 2: positions
    double run2 = DINF;
    int rec_before = 0, non_before = 0;
    for (int b = 0; b < nf; b += 32) {
        const bool in = b + lane < nf;
        const double c = in ? cost[b + lane] : DINF;
        const double incoming = fmin(run2, warp_excl_min(c, lane));
        const bool isrec = in && c < incoming;
        const unsigned br = __ballot_sync(FULL, isrec), bn = __ballot_sync(FULL, in && !isrec);
        const unsigned lt = (1u << lane) - 1u;
        if (isrec) fin[b + lane] = r0 + nrec - 1 - (rec_before + __popc(br & lt));
        else if (in) fin[b + lane] = r0 + nrec + non_before + __popc(bn & lt);
        rec_before += __popc(br); non_before += __popc(bn);
        run2 = fmin(run2, warp_min(c));
    }
    if (lane == 0) cur.cost[t] = running;
}

// With many more trips than warps (the SBA peak epoch: a million trips with one to five feasible schedules each) a warp takes 32
// consecutive trips: those with a handful of schedules are decided by ONE lane each, walking the trip's costs as the reference's loop
// does; the others go through the whole warp one after the other.  Otherwise (the OSP levels: ~10^4 trips) a warp per trip.
#define CLASSIFY_SMALL 8
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_classify(Control* __restrict__ ctl, Level cur, int heavy_min, int* __restrict__ heavy_list) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int n = ctl->overflow ? 0 : ctl->n_trips[cur.k];
    const int nwarps = gridDim.x * WARPS_PER_BLOCK;
    if (n < nwarps * 4) {        // fewer trips than lanes to spread them over (the OSP levels): a warp per trip
        for (int t = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); t < n; t += nwarps) {
            const int nf = cur.nfeas[t];
            if (heavy_min > 0 && nf >= heavy_min) {      // left to k_classify_heavy (one block per such trip)
                if (lane == 0) heavy_list[atomicAdd(&ctl->n_heavy[cur.k], 1)] = t;
                continue;
            }
            classify_warp(cur, t, cur.s0[t], nf, lane);
        }
        return;
    }
    for (int t0 = (blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5)) * 32; t0 < n; t0 += nwarps * 32) {
        const int t = t0 + lane;
        i64 s0 = 0;
        int nf = 0;
        if (t < n) { s0 = cur.s0[t]; nf = cur.nfeas[t]; }
        const bool heavy = t < n && heavy_min > 0 && nf >= heavy_min;      // left to k_classify_heavy (one block per such trip)
        if (heavy) heavy_list[atomicAdd(&ctl->n_heavy[cur.k], 1)] = t;
        const bool small = t < n && !heavy && nf <= CLASSIFY_SMALL;
        if (small && BOK(s0 + nf <= cur.cap_sch)) {
            const double* __restrict__ cost = cur.sch_cost + s0;
            int* __restrict__ fin = cur.final + s0;
            double c[CLASSIFY_SMALL];
#pragma unroll
            for (int x = 0; x < CLASSIFY_SMALL; x++) c[x] = x < nf ? cost[x] : DINF;
            double run = DINF;
            int nrec = 0;
#pragma unroll
            for (int x = 0; x < CLASSIFY_SMALL; x++) if (c[x] < run) { run = c[x]; nrec++; }
            double run2 = DINF;
            int rb = 0, nb = 0;
#pragma unroll
            for (int x = 0; x < CLASSIFY_SMALL; x++)
                if (x < nf) {
                    if (c[x] < run2) { run2 = c[x]; fin[x] = (int)s0 + nrec - 1 - rb; rb++; }
                    else { fin[x] = (int)s0 + nrec + nb; nb++; }
                }
            cur.cost[t] = run;
        }
        unsigned todo = __ballot_sync(FULL, t < n && !heavy && !small);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            classify_warp(cur, t0 + src, __shfl_sync(FULL, s0, src), __shfl_sync(FULL, nf, src), lane);
        }
    }
}

// Trips with thousands of feasible schedules (large capacities): one block per trip, every thread owns a
// contiguous segment of the encounter-order costs and walks it sequentially; the segments are stitched with
// two block-wide scans (running minimum entering each segment, record / non-record counts before it).
#define HEAVY_THREADS 512
__global__ void __launch_bounds__(HEAVY_THREADS)
k_classify_heavy(const Control* __restrict__ ctl, Level cur, const int* __restrict__ heavy_list) {
    pdl_enter();
    __shared__ double s_min[HEAVY_THREADS];
    __shared__ int s_rec[HEAVY_THREADS], s_non[HEAVY_THREADS];
    const int n = ctl->overflow ? 0 : ctl->n_heavy[cur.k];
    const int tid = threadIdx.x;
    for (int h = blockIdx.x; h < n; h += gridDim.x) {
        const int t = heavy_list[h];
        const i64 s0 = cur.s0[t];
        const int nf = cur.nfeas[t];
        const double* __restrict__ cost = cur.sch_cost + s0;
        int* __restrict__ fin = cur.final + s0;
        const int r0 = (int)s0;
        const int seg = (nf + HEAVY_THREADS - 1) / HEAVY_THREADS;
        const int a = min(nf, tid * seg), b = min(nf, a + seg);
        double m = DINF;
        for (int x = a; x < b; x++) m = fmin(m, cost[x]);
        s_min[tid] = m;
        __syncthreads();
        if (tid == 0) {      // exclusive prefix-min over the segments (512 values: one thread is enough)
            double run = DINF;
            for (int x = 0; x < HEAVY_THREADS; x++) { const double v = s_min[x]; s_min[x] = run; run = fmin(run, v); }
            cur.cost[t] = run;
        }
        __syncthreads();
        const double incoming = s_min[tid];
        double run = incoming;
        int nrec = 0;
        for (int x = a; x < b; x++) { const double c = cost[x]; if (c < run) { run = c; nrec++; } }
        s_rec[tid] = nrec; s_non[tid] = (b - a) - nrec;
        __syncthreads();
        if (tid == 0) {      // exclusive prefix sums of both counts; the total record count goes to s_rec[0] slot after use
            int r = 0, q = 0;
            for (int x = 0; x < HEAVY_THREADS; x++) { const int vr = s_rec[x], vq = s_non[x]; s_rec[x] = r; s_non[x] = q; r += vr; q += vq; }
            s_min[0] = (double)r;      // total number of records
        }
        __syncthreads();
        const int total_rec = (int)s_min[0];
        int rb = s_rec[tid], nb = s_non[tid];
        run = incoming;
        for (int x = a; x < b; x++) {
            const double c = cost[x];
            if (c < run) { run = c; fin[x] = r0 + total_rec - 1 - rb++; }
            else fin[x] = r0 + total_rec + nb++;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Candidates of size k+1 from level k (dispatch_osp.py:183-216) in closed form: a (k+1)-set U is
// searched iff all its k-subsets are feasible; it is searched from its lowest-index subset
// (trip_i), inserting the one request U - trip_i, and the searches of one vehicle run in
// (i, j) order, j = second-lowest subset index.  One warp per row (= trip_i); lanes scan the
// vehicle's size-1 trips for the request to add.
//
// Two launches, no scan kernel in between.
// k_gen_pairs finds every row's (j, q) pairs ONCE and stores them in a region of the pair buffer claimed with one
// atomicAdd per row (regions are in no particular order; the row keeps its base), and leaves one packed
// (tasks | rows << 32) sum per contiguous segment of rows.
// k_gen_emit (every block owns one such segment) adds up the sums of the segments before it, scans its own in shared
// memory, ranks each row's pairs by j and emits the next level's tasks AND their row descriptors (one per base
// schedule the task walks); the last block publishes the level's task / row / chunk counts.
// A level has fewer rows than the GPU has warps, so both kernels last as long as the longest chain of dependent L2
// round trips of one row: 4 to reach the vehicle's size-1 trips, 2 per lookup phase (see TripHash).
// ---------------------------------------------------------------------------------------------
struct GenOut {
    int* task_veh; int* task_base; int* task_q; i64* task_row0;
    RowDesc* rows; int* task_nfeas;
    i64* task_first;          // [trips of the level + 1] first task spawned by each trip (-> veh_start of the next level)
    i64 cap_tasks, cap_rows, cap_chunks; int G_next;
};

__device__ __forceinline__ void gen_segment(int n, int* lo, int* hi) {
    const int seg = (n + (int)gridDim.x - 1) / (int)gridDim.x;
    *lo = min(n, (int)blockIdx.x * seg);
    *hi = min(n, *lo + seg);
}

#define GEN_NS 2          // 32-wide slices of a vehicle's size-1 trips a warp handles in one go (first probes in flight together)
struct GenWarp { int j[GEN_NS * 32]; int r[GEN_NS * 32]; unsigned char list[GEN_NS * 32]; };

// Candidates of row t among the vehicle's size-1 trips [c0, c0 + 32*GEN_NS) below a1: per slice s, lane `lane` holds
// the request rv[s] of size-1 trip c0 + 32 s + lane and jv[s] = the second-lowest subset index, or -1 (no candidate).
// The k lookups of a candidate are not walked one after the other: every lane first tries the subset without
// mine[0]; the (few) survivors' remaining k-1 lookups are then spread over the lanes, one lookup deep.
template <int KMAX>
__device__ __forceinline__ void gen_block(const Level& l1, const Level& prev, const TripHash& th, int t, int v, const ReqTuple<KMAX>& mine,
                                          int c0, int a1, int lane, GenWarp& sh, int (&jv)[GEN_NS], int (&rv)[GEN_NS]) {
    const int k = prev.k;
#pragma unroll
    for (int s = 0; s < GEN_NS; s++) { const int c = c0 + 32 * s + lane; rv[s] = c < a1 ? l1.trip_req[c] : -1; jv[s] = -1; }
    if (KMAX == 1 || k == 1) {                      // all pairs i<j (dispatch_osp.py:183-186)
#pragma unroll
        for (int s = 0; s < GEN_NS; s++) { const int c = c0 + 32 * s + lane; if (c < a1 && c > t) jv[s] = c; }
        return;
    }
    // every other k-subset must exist with a larger index; first the one without mine[0]
    HashProbe pb[GEN_NS];
    bool go[GEN_NS];
#pragma unroll
    for (int s = 0; s < GEN_NS; s++) {
        go[s] = false;
        if (c0 + 32 * s < a1) {                     // (warp-uniform)
            bool in = rv[s] < 0;
#pragma unroll
            for (int m = 0; m < KMAX; m++) in |= (m < k) & (mine.r[m] == rv[s]);
            go[s] = !in;
            if (go[s]) pb[s] = hash_begin(th, v, merged_tuple(mine, k, 0, rv[s]), k);
        }
    }
    int nsurv = 0;
    const unsigned lt = (1u << lane) - 1u;
    __syncwarp();
#pragma unroll
    for (int s = 0; s < GEN_NS; s++) {
        if (c0 + 32 * s < a1) {
            if (go[s]) { const int u = hash_finish(prev, th, v, merged_tuple(mine, k, 0, rv[s]), pb[s]); if (u > t) jv[s] = u; }
            const unsigned bal = __ballot_sync(FULL, jv[s] >= 0);
            const int slot = 32 * s + lane;
            sh.j[slot] = jv[s]; sh.r[slot] = rv[s];
            if (jv[s] >= 0) sh.list[nsurv + __popc(bal & lt)] = (unsigned char)slot;
            nsurv += __popc(bal);
        }
    }
    __syncwarp();
    const int items = nsurv * (k - 1);
    for (int it0 = 0; it0 < items; it0 += 32) {
        const int it = it0 + lane;
        if (it < items) {
            const int si = it / (k - 1), d = 1 + it % (k - 1);
            const int slot = sh.list[si];
            const int u = hash_find(prev, th, v, merged_tuple(mine, k, d, sh.r[slot]));
            atomicMin(&sh.j[slot], u > t ? u : -1);      // j = lowest subset index; any subset missing or not above t: rejected
        }
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < GEN_NS; s++) if (c0 + 32 * s < a1) jv[s] = sh.j[32 * s + lane];
    __syncwarp();
}

template <int KMAX>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_gen_pairs(const Control* __restrict__ ctl, Level l1, Level prev, TripHash th,
            int* __restrict__ row_cnt, int* __restrict__ row_base, int2* __restrict__ pairs, i64 cap_pairs,
            unsigned long long* __restrict__ blk_partial, int n_seg /* blocks of the emit pass */, unsigned long long* __restrict__ cursor,
            unsigned long long* __restrict__ next_hash, unsigned next_hash_size) {
    pdl_enter();
    __shared__ GenWarp s_w[WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    GenWarp& sh = s_w[wib];
    const int n = ctl->overflow ? 0 : ctl->n_trips[prev.k];
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    // the lookup table the NEXT level will build into (the other buffer; its last readers are two kernels back)
    for (unsigned x = gtid; x < next_hash_size; x += gsz) next_hash[x] = HASH_EMPTY;
    // Rows are dealt out cyclically over ALL warps of the grid (a vehicle's trips are consecutive rows); the emit pass owns
    // contiguous segments, so every row adds its (tasks | rows << 32) to the sum of the segment it belongs to
    // (blk_partial and the cursor are zeroed by k_compact_trips, the kernel before this one on the branch).
    const int seg = max(1, (n + n_seg - 1) / n_seg);
    const unsigned lt = (1u << lane) - 1u;
    const int k = prev.k;
    for (int t = blockIdx.x * WARPS_PER_BLOCK + wib; t < n; t += gridDim.x * WARPS_PER_BLOCK) {
        const int v = prev.trip_veh[t];
        const ReqTuple<KMAX> mine = load_tuple<KMAX>(prev.trip_req + (i64)t * k, k);
        const int a0 = l1.veh_start[v], a1 = l1.veh_start[v + 1];
        int jv[GEN_NS], rv[GEN_NS];
        int cnt = 0;
        i64 base = 0;
        if (a1 - a0 <= 32 * GEN_NS) {
            gen_block(l1, prev, th, t, v, mine, a0, a1, lane, sh, jv, rv);
            int pos[GEN_NS];
#pragma unroll
            for (int s = 0; s < GEN_NS; s++) { const unsigned bal = __ballot_sync(FULL, jv[s] >= 0); pos[s] = cnt + __popc(bal & lt); cnt += __popc(bal); }
            if (cnt) {
                if (lane == 0) base = (i64)atomicAdd(cursor, (unsigned long long)cnt);
                base = __shfl_sync(FULL, base, 0);
                if (base + cnt <= cap_pairs) {       // (else the emit pass reports the task overflow and the epoch is rerun)
#pragma unroll
                    for (int s = 0; s < GEN_NS; s++) if (jv[s] >= 0) pairs[base + pos[s]] = make_int2(jv[s], rv[s]);
                }
            }
        } else {                                     // a vehicle with more size-1 trips: count, claim, find again
            for (int c0 = a0; c0 < a1; c0 += 32 * GEN_NS) {
                gen_block(l1, prev, th, t, v, mine, c0, a1, lane, sh, jv, rv);
#pragma unroll
                for (int s = 0; s < GEN_NS; s++) cnt += __popc(__ballot_sync(FULL, jv[s] >= 0));
            }
            if (cnt) {
                if (lane == 0) base = (i64)atomicAdd(cursor, (unsigned long long)cnt);
                base = __shfl_sync(FULL, base, 0);
                int w = 0;
                for (int c0 = a0; c0 < a1; c0 += 32 * GEN_NS) {
                    gen_block(l1, prev, th, t, v, mine, c0, a1, lane, sh, jv, rv);
#pragma unroll
                    for (int s = 0; s < GEN_NS; s++) {
                        const unsigned bal = __ballot_sync(FULL, jv[s] >= 0);
                        if (jv[s] >= 0 && base + cnt <= cap_pairs) pairs[base + w + __popc(bal & lt)] = make_int2(jv[s], rv[s]);
                        w += __popc(bal);
                    }
                }
            }
        }
        if (lane == 0) {
            row_cnt[t] = cnt;
            row_base[t] = (int)min(base, (i64)0x7fffffff);
            if (cnt) atomicAdd(&blk_partial[t / seg], (unsigned long long)cnt | ((unsigned long long)cnt * (unsigned long long)prev.nfeas[t]) << 32);
        }
    }
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_gen_emit(Control* __restrict__ ctl, int which_next, Level prev, const int* __restrict__ len0,
           const int* __restrict__ row_cnt, const int* __restrict__ row_base, const int2* __restrict__ pairs,
           const unsigned long long* __restrict__ blk_partial, GenOut o, int wide /* != 0: every row ranks through memory (tests) */) {
    pdl_enter();
    __shared__ i64 s_scan[SCAN_THREADS / 32 + 1];
    __shared__ i64 s_off[WARPS_PER_BLOCK * 32];
    __shared__ int s_q[WARPS_PER_BLOCK][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n = ctl->overflow ? 0 : ctl->n_trips[prev.k];
    int lo, hi;
    gen_segment(n, &lo, &hi);
    const bool last = blockIdx.x == gridDim.x - 1;
    if (lo >= hi && !last) return;
    // packed (tasks | rows << 32) total of the blocks before this one
    i64 p = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += blockDim.x) p += (i64)blk_partial[b];
    i64 carry;
    block_excl_scan(p, &carry, s_scan);
    unsigned long long n_tasks = 0, n_items = 0;
    const int k = prev.k;
    for (int base = lo; base < hi; base += WARPS_PER_BLOCK * 32) {
        const int tt = base + threadIdx.x;
        i64 x = 0;
        if (tt < hi) { const i64 c = row_cnt[tt]; x = c | ((c * (i64)prev.nfeas[tt]) << 32); }
        i64 tile_tot;
        s_off[threadIdx.x] = block_excl_scan(x, &tile_tot, s_scan) + carry;
        __syncthreads();
        for (int i = wib; i < WARPS_PER_BLOCK * 32 && base + i < hi; i += WARPS_PER_BLOCK) {
            const int t = base + i;
            const i64 off = PACK_LO(s_off[i]), rbase = PACK_HI(s_off[i]);
            const int cnt = row_cnt[t];
            const int S = prev.nfeas[t];          // every task inserting into trip t walks all of its schedules
            if (lane == 0) o.task_first[t] = off;
            if (cnt == 0) continue;
            if (off + cnt > o.cap_tasks || rbase + (i64)cnt * S > o.cap_rows) continue;   // reported by the last block; the epoch is rerun
            if ((i64)row_base[t] + cnt > o.cap_tasks) continue;                           // (its pairs did not fit the pair buffer: same overflow)
            const int v = prev.trip_veh[t];
            const int2* __restrict__ pp = pairs + row_base[t];
            const i64 s0 = prev.s0[t];
            const int l = len0[v] + 2 * k;
            // rank the pairs by j (k == 1 rows are already in j order) and emit the next level's tasks
            const bool narrow = cnt <= 32 && !wide;
            if (narrow) {
                int2 me = make_int2(0x7fffffff, 0);
                if (lane < cnt) me = pp[lane];
                int rank = lane;
                if (k > 1) { rank = 0; for (int y = 0; y < cnt; y++) rank += __shfl_sync(FULL, me.x, y) < me.x; }
                if (lane < cnt) {
                    const i64 w = off + rank;
                    if (BOK(w >= 0 && w < o.cap_tasks && rank < cnt)) {
                        o.task_veh[w] = v; o.task_base[w] = t; o.task_q[w] = me.y;
                        o.task_row0[w] = rbase + (i64)rank * S;
                        o.task_nfeas[w] = 0;
                    }
                    s_q[wib][rank & 31] = me.y;
                }
            } else {
                for (int e = lane; e < cnt; e += 32) {
                    const int2 me = pp[e];
                    int rank = e;
                    if (k > 1) { rank = 0; for (int y = 0; y < cnt; y++) rank += pp[y].x < me.x; }
                    const i64 w = off + rank;
                    if (BOK(w >= 0 && w < o.cap_tasks && rank < cnt)) {
                        o.task_veh[w] = v; o.task_base[w] = t; o.task_q[w] = me.y;
                        o.task_row0[w] = rbase + (i64)rank * S;
                        o.task_nfeas[w] = 0;
                    }
                }
            }
            __syncwarp();
            // ... and one row descriptor per (task, base schedule): task `off + e` owns rows rbase + e*S ..
            const int n_rows = cnt * S;            // (< 2^31: bounded by cap_rows above)
            for (int idx = lane; idx < n_rows; idx += 32) {
                const int e = idx / S, xs = idx - e * S;
                RowDesc d;
                d.v = v; d.q = narrow ? s_q[wib][e] : o.task_q[off + e]; d.s = xs; d.l = l; d.s0 = s0; d.task = (int)(off + e); d.pad = 0;
                if (BOK(rbase + idx < o.cap_rows)) o.rows[rbase + idx] = d;
            }
            __syncwarp();
            if (lane == 0) { n_tasks += (unsigned long long)cnt; n_items += (unsigned long long)n_rows * (unsigned long long)(l + 1); }
        }
        carry += tile_tot;
        __syncthreads();
    }
    if (lane == 0 && n_tasks) { atomicAdd(&ctl->stats.n_tasks, n_tasks); atomicAdd(&ctl->stats.n_items, n_items); }
    if (last && threadIdx.x == 0) {
        o.task_first[n] = PACK_LO(carry);
        FinTasksRows fin{ctl, which_next, o.cap_tasks, o.cap_rows, o.cap_chunks, o.G_next};
        fin(carry);
    }
}
