// eval.cuh — the insertion search (scheduling.py:13-98) as a warp-cooperative kernel.
//
// A *task* is one compute_schedule() call (vehicle v, base trip, inserted request q); it is cut
// into *chunks* of whole base schedules so that one warp step covers one chunk: work items are
// (sub_sche s, pick-up index i) pairs in the reference's loop order, one per lane, and a lane walks
// the drop-off index j sequentially with incremental partial sums.  That is exactly the reference's
// inner loop, so its early exits (scheduling.py:41-44) apply per lane (j-loop breaks) and by ballot
// (viol==3 ends the i-loop of that sub_sche).  Chunks of one task run on different warps: heavy
// tasks (thousands of base schedules) no longer serialise on one warp.
// All arithmetic is float64 accumulated left to right in the reference's order: costs and every
// comparison are bit-identical to the reference.
//
// ONE pass: every feasible insertion is appended, as a 16 B record (cost, chunk, lane, ordinal within the lane,
// i, j, row), to the launching block's region of a record buffer (a block-local cursor in shared memory: no
// global atomics, no ordering between blocks).  The per-lane counts give, by one prefix sum over the chunks, the
// record's position in the reference's encounter order; place.cuh then stores cost and schedule there.  The
// search itself is never repeated.
#pragma once
#undef BFILE
#define BFILE 1
#include "common.cuh"
#include "scan.cuh"

#ifndef EVAL_BLOCKS_PER_SM
#define EVAL_BLOCKS_PER_SM 4
#endif
#ifndef EVAL_WARPS
#define EVAL_WARPS WARPS_PER_BLOCK     // warps per block of the insertion search
#endif
#define STAGE_SLOTS 96    // stops staged per warp step: nsub*l <= 32 (whole schedules) or l <= 38
#define LEG_SLOTS 200     // travel times staged per warp step: nsub*(5l+2) <= 192

// Per base schedule of length l the insertion search only ever needs 5l+2 distinct table entries
// (+ tt[P][D]): the base legs, and the legs into / out of the inserted pick-up P and drop-off D at
// every position.  They are gathered once per warp step, all lanes in parallel (one L2 round trip),
// and the sequential walks then read shared memory.  Values are the table's own elements, converted
// to float64 at use, so the arithmetic is unchanged.
template <typename TT>
struct WarpStage {
    double ddl[STAGE_SLOTS];
    int node[STAGE_SLOTS];
    TT leg[LEG_SLOTS];        // per schedule: g[l] | toP[l+1] | fromP[l] | toD[l+1] | fromD[l]
    signed char pod[STAGE_SLOTS];
};

// One feasible insertion found by the search.  info = lane (6 bits: 0..63, the work item within the chunk)
// | ordinal among the lane's feasible insertions << 6 | j << 12 | i << 18 | row within the chunk << 24.
struct __align__(16) Tuple { double cost; int c; unsigned info; };
#define TUP_LANE(x) ((int)((x) & 63u))
#define TUP_ORD(x) ((int)(((x) >> 6) & 63u))
#define TUP_J(x) ((int)(((x) >> 12) & 63u))
#define TUP_I(x) ((int)(((x) >> 18) & 63u))
#define TUP_ROW(x) ((int)(((x) >> 24) & 31u))

struct TupleOut {
    Tuple* region;      // this block's region of the record buffer
    int cap;            // records per region
    int* cursor;        // shared-memory cursor of the block
    int c;              // chunk
    unsigned tag;       // lane | i << 18 | row << 24
};

struct LaneCount { int n, evals, lookups; };

// x / d for 32-bit x by one multiplication (d fixed per launch): q = hi(x * floor(2^32 / d)) is the quotient or one below it
struct FastDiv {
    unsigned d, m;
    __device__ __forceinline__ explicit FastDiv(unsigned dd) : d(dd), m(dd > 1 ? (unsigned)(0x100000000ull / dd) : 0u) {}
    __device__ __forceinline__ unsigned div(unsigned x) const {
        if (d <= 1) return x;
        unsigned q = __umulhi(x, m);
        return q + ((x - q * d) >= d ? 1u : 0u);
    }
};

// Walk j = i+1 .. for one (sub_sche, i).  `A` = accumulated time at the inserted pick-up.
// Every feasible (i, j) is appended to the block's record region with its cost.
template <typename TT>
__device__ __forceinline__ void walk_j(const TT* __restrict__ lg, double PD, const double* __restrict__ ddl,
                                       int l, int i, double A, double cld, double T, unsigned long long fullmask,
                                       LaneCount& o, const TupleOut& out) {
    const TT* __restrict__ g = lg;                       // g[m]     = tt[stop m-1 (or vehicle)][stop m]
    const TT* __restrict__ fromP = lg + 2 * l + 1;       // fromP[m] = tt[P][stop m]
    const TT* __restrict__ toD = lg + 3 * l + 1;         // toD[m]   = tt[stop m-1 (or vehicle)][D]
    const TT* __restrict__ fromD = lg + 4 * l + 2;       // fromD[m] = tt[D][stop m]
    // capacity (scheduling.py:74-77): (i,j) is over capacity iff fullmask has a bit in [i, j-1]
    unsigned long long hi = fullmask >> (i + 1);
    const int b = hi ? (i + __ffsll((long long)hi)) : 64;     // lowest set bit above i; j <= b is capacity-feasible
    double B = A;
    for (int j = i + 1;; j++) {
        o.evals++;
        double C = B + (j == i + 1 ? PD : (double)toD[j - 1]);
        if (T + C > cld) { o.lookups += j + 1; break; }              // viol 2 -> break (scheduling.py:41)
        bool ok = true;
        for (int m = j - 1; m < l; m++) {
            C += (double)(m == j - 1 ? fromD[m] : g[m]);
            if (T + C > ddl[m]) { ok = false; o.lookups += m + 3; break; }   // viol 0: infeasible, keep going
        }
        if (ok) {
            o.lookups += l + 2;
            const int slot = atomicAdd(out.cursor, 1);
            if (slot < out.cap && BOK(slot >= 0)) {
                Tuple t; t.cost = C; t.c = out.c; t.info = out.tag | ((unsigned)o.n << 6) | ((unsigned)j << 12);
                out.region[slot] = t;
            }
            o.n++;
        }
        if (j == l + 1) break;
        if (j == b) { o.evals++; break; }                           // next j is over capacity: viol 1 -> break
        B += (double)(j == i + 1 ? fromP[j - 1] : g[j - 1]);        // extend the pick-up side over base stop j-1
        if (T + B > ddl[j - 1]) { o.evals++; o.lookups += j + 1; break; }   // viol 4 at (i, j+1) -> break
    }
}

// A chunk = G consecutive rows (base schedules) of the level's global row list, whatever task or
// vehicle they belong to; G is chosen per level so that G * (longest schedule + 1) <= 32 lanes.
template <typename TT>
__global__ void __launch_bounds__(EVAL_WARPS * 32, EVAL_BLOCKS_PER_SM)
k_eval(const DEpoch* __restrict__ epp, Table<TT> tt, Level prev, Level cur, Control* __restrict__ ctl, int which, int G,
       const RowDesc* __restrict__ rows, int* __restrict__ chunk_cnt, uint8_t* __restrict__ lane_cnt,
       int* __restrict__ task_nfeas, Stats* stats, unsigned long long* __restrict__ part, int part_vg,
       Tuple* __restrict__ tup, int* __restrict__ tup_cnt, int region_cap) {
    pdl_enter();
    __shared__ WarpStage<TT> stage_all[EVAL_WARPS];
    __shared__ DEpoch s_ep;       // the epoch header once per block: array pointers without a second indirection
    __shared__ int s_cursor;      // feasible-insertion records appended by this block
    if (threadIdx.x < sizeof(DEpoch) / 4) ((int*)&s_ep)[threadIdx.x] = ((const int*)epp)[threadIdx.x];
    if (threadIdx.x == 0) s_cursor = 0;
    __syncthreads();
    const DEpoch& ep = s_ep;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpStage<TT>& st = stage_all[wib];
    const int warp0 = blockIdx.x * EVAL_WARPS + wib, nwarps = gridDim.x * EVAL_WARPS;
    unsigned long long acc_evals = 0, acc_lookups = 0;
    const double T = ep.T;
    const int n_rows = ctl->overflow ? 0 : ctl->n_rows[which];
    const int n_chunks = ctl->overflow ? 0 : ctl->n_chunks[which];
    const int cap = ep.cap;
    // the block sums of the two scans that follow (schedules per chunk; feasible tasks | their schedules << 32) are left as we go
    const FastDiv seg_chunks((unsigned)scan_seg_len(n_chunks, part_vg)), seg_tasks((unsigned)scan_seg_len(ctl->n_tasks[which], part_vg));
    TupleOut out;
    out.region = tup + (i64)blockIdx.x * region_cap; out.cap = region_cap; out.cursor = &s_cursor; out.c = 0; out.tag = 0;

    RowDesc nxt; nxt.l = 0; nxt.v = 0; nxt.q = 0; nxt.s = 0; nxt.s0 = 0; nxt.task = 0;
    if (warp0 < n_chunks && lane < G && warp0 * G + lane < n_rows) nxt = rows[(i64)warp0 * G + lane];
    for (int c = warp0; c < n_chunks; c += nwarps) {
        const RowDesc rd = nxt;                                   // lane j < nr owns row j of the chunk
        const int nr = min(G, n_rows - c * G);
        {
            const i64 cn = (i64)c + nwarps;
            if (cn < n_chunks && lane < G && cn * G + lane < n_rows) nxt = rows[cn * G + lane];   // prefetch
        }
        const bool own = lane < nr;
        // per-row parameters live in the owner lane's registers and are read by shuffle
        int o_l = own ? rd.l : 0;
        const bool too_long = own && (o_l + 2 > MAXS || o_l + 2 > cur.stride);
        if (__any_sync(FULL, too_long)) {   // compiled stop limit: reported as AMOD_E_LIMIT by the host
            if (lane == 0) { atomicOr(&stats->flags, 1ull); chunk_cnt[c] = 0; }
            lane_cnt[(i64)c * 64 + lane] = 0; lane_cnt[(i64)c * 64 + 32 + lane] = 0;
            continue;
        }
        int o_P = 1, o_D = 1, o_nid = 1, o_load = 0;
        double o_clp = 0, o_cld = 0, o_t0 = 0, o_PD = 0;
        i64 o_src = 0;
        if (own) {
            o_P = ep.onid[rd.q]; o_D = ep.dnid[rd.q]; o_clp = ep.clp[rd.q]; o_cld = ep.cld[rd.q];
            o_nid = ep.veh_nid[rd.v]; o_load = ep.veh_load[rd.v]; o_t0 = ep.veh_t[rd.v];
            o_src = rd.s0 + rd.s;                                 // rows are stored in list order: row s IS feasible_sches[s]
            o_PD = tt(o_P, o_D);
        }
        // One prefix sum over the rows' lengths places everything: row j (owner lane j) starts at lane L + j, its staged stops at
        // slot L and its staged legs (5l+2 per row) at slot 5L + 2j, L = stops of the rows before it.
        const int o_len = own ? o_l : 0;
        const int o_L = warp_excl_sum(o_len, lane);
        const int o_j = min(lane, nr);
        const int o_woff = o_L + o_j;                             // first lane of the row
        const int items = __shfl_sync(FULL, o_L + o_len, 31) + nr;
        const int o_soff = o_L;                                   // first staged stop of the row
        const int o_goff = 5 * o_L + 2 * o_j;                     // first staged leg of the row

        // lane -> (row, pick-up index): the rows' first lanes as a bit mask (a single row of more than 31 stops starts at lane 0)
        const unsigned starts = __reduce_or_sync(FULL, own ? (1u << (o_woff & 31)) : 0u);
        const int r = __popc(starts & (0xffffffffu >> (31 - lane))) - 1;
        const int i_first = lane - __shfl_sync(FULL, o_woff, r);
        const int l = __shfl_sync(FULL, o_l, r);
        const int soff = __shfl_sync(FULL, o_soff, r), goff = __shfl_sync(FULL, o_goff, r);
        const int task = __shfl_sync(FULL, rd.task, r);
        const int load = __shfl_sync(FULL, o_load, r);
        const double clp = __shfl_sync(FULL, o_clp, r), cld = __shfl_sync(FULL, o_cld, r);
        const double t0 = __shfl_sync(FULL, o_t0, r), PD = __shfl_sync(FULL, o_PD, r);
        {   // every row is staged by its own lanes: stops first, then the 5l+2 travel times it can ever need
            const int P = __shfl_sync(FULL, o_P, r), D = __shfl_sync(FULL, o_D, r), nid = __shfl_sync(FULL, o_nid, r);
            const i64 src = __shfl_sync(FULL, o_src, r);
            const bool stage = lane < items;
            const int wl = min(l + 1, 32);                       // lanes this row has in the first step
            __syncwarp();
            if (stage)
                for (int m = i_first; m < l; m += wl) {
                    const Stop sp = prev.sch[src * prev.stride + m];      // one 16 B load: code, node, deadline
                    st.node[soff + m] = sp.node; st.ddl[soff + m] = sp.ddl; st.pod[soff + m] = (signed char)pod_of(sp.code);
                }
            __syncwarp();
            if (stage) {
                // the lane at position m fetches the five legs that touch position m, branch-free and
                // back to back (one L2 round trip): g[m], toP[m], fromP[m], toD[m], fromD[m]
                const int* n = st.node + soff;
                TT* lg = st.leg + goff;
                for (int m = i_first; m <= l; m += wl) {
                    const bool in = m < l;
                    const int a = m ? n[m - 1] : nid;
                    const int b = in ? n[m] : P;
                    const TT g = tt.raw(a, b), toP = tt.raw(a, P), frP = tt.raw(P, b), toD = tt.raw(a, D), frD = tt.raw(D, b);
                    lg[l + m] = toP; lg[3 * l + 1 + m] = toD;
                    if (in) { lg[m] = g; lg[2 * l + 1 + m] = frP; lg[4 * l + 2 + m] = frD; }
                }
            }
            __syncwarp();
        }

        int dead_s = -1, cnt_total = 0;
        out.c = c;
        for (int it0 = 0; it0 < items; it0 += 32) {               // one step unless a single schedule has > 31 stops
            const int it = it0 + lane;
            const bool valid = it < items;
            const int i = it0 + i_first;                          // (a second step only exists for a single-row chunk)
            const double* ddl = st.ddl + soff;
            const TT* lg = st.leg + goff;

            // capacity profile of the base schedule and accumulated time at the inserted pick-up
            unsigned long long fullmask = 0;
            bool basebad = false, stop3 = false, cap1 = false;
            double A = 0.0;
            if (items <= 32) {
                // occupancy after every base stop by a prefix sum over the row's own lanes, shared by ballot
                const bool at = valid && i < l;
                int run = at ? (int)st.pod[soff + i] : 0;
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, run, o); if (i >= o) run += t; }
                run += load;
                const unsigned bal_full = __ballot_sync(FULL, at && run >= cap);
                const unsigned bal_over = __ballot_sync(FULL, at && run > cap);
                const unsigned rowbits = l >= 31 ? FULL : ((2u << l) - 1u);
                const int woff = lane - i;
                fullmask = ((unsigned long long)((bal_full >> woff) & rowbits) << 1) | (load >= cap ? 1ull : 0ull);
                basebad = ((bal_over >> woff) & rowbits) != 0u;
            }
            if (valid) {
                if (items > 32) {
                    int run = load;
                    if (run >= cap) fullmask |= 1ull;
                    for (int m = 0; m < l; m++) {
                        run += st.pod[soff + m];
                        if (run > cap) basebad = true;
                        if (run >= cap) fullmask |= 1ull << (m + 1);
                    }
                }
                double acc = t0;
                for (int m = 0; m < i; m++) acc += (double)lg[m];                        // trusted prefix (no checks)
                A = acc + (double)lg[l + i];                                             // toP[i]
                cap1 = basebad || ((fullmask >> i) & 1ull);                              // viol 1 at j = i+1
                stop3 = !cap1 && (T + A > clp);                                          // viol 3 (scheduling.py:43)
            }
            LaneCount o; o.n = 0; o.evals = 0; o.lookups = 0;
            // viol==3 ends the i-loop of its base schedule: later i of the same row are never evaluated
            const unsigned bal3 = __ballot_sync(FULL, stop3);
            const int lo = max(lane - i, 0);
            const unsigned same_s_lower = (lane > lo) ? (((1u << lane) - 1u) & ~((1u << lo) - 1u)) : 0u;
            const bool dead = valid && ((bal3 & same_s_lower) != 0u || dead_s == 0);
            {   // carry to the second step of a > 31-stop schedule (then the chunk is that single row)
                const unsigned killed = __ballot_sync(FULL, valid && (stop3 || dead));
                dead_s = (items > 32 && killed) ? 0 : -1;
            }
            const bool live = valid && !dead;
            if (live) {
                if (cap1) o.evals = 1;
                else if (stop3) { o.evals = 1; o.lookups = i + 1; }
                else {
                    out.tag = (unsigned)it | ((unsigned)i << 18) | ((unsigned)r << 24);
                    walk_j<TT>(lg, PD, ddl, l, i, A, cld, T, fullmask, o, out);
                }
            }
            acc_evals += o.evals; acc_lookups += o.lookups;
            lane_cnt[(i64)c * 64 + it0 + lane] = (uint8_t)o.n;     // <= 64 items per chunk
            if (o.n) {      // block sums of the scan over tasks: (task has a feasible schedule) | (its schedules) << 32
                const int before = atomicAdd(&task_nfeas[task], o.n);
                atomicAdd(&part[part_vg + seg_tasks.div((unsigned)task)], (before == 0 ? 1ull : 0ull) | ((unsigned long long)o.n << 32));
            }
            cnt_total += o.n;
        }
        if (items <= 32) lane_cnt[(i64)c * 64 + 32 + lane] = 0;
        cnt_total = warp_sum(cnt_total);
        if (lane == 0) { chunk_cnt[c] = cnt_total; if (cnt_total) atomicAdd(&part[seg_chunks.div((unsigned)c)], (unsigned long long)cnt_total); }
    }
    acc_evals = warp_sum(acc_evals);
    acc_lookups = warp_sum(acc_lookups);
    if (lane == 0 && (acc_evals | acc_lookups)) {
        atomicAdd(&stats->n_evals, acc_evals);
        atomicAdd(&stats->n_lookups, acc_lookups);
        atomicAdd(&stats->n_lookups_eval, acc_lookups);
        atomicAdd(&stats->lvl_evals[cur.k], acc_evals);
        atomicAdd(&stats->lvl_lookups[cur.k], acc_lookups);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n = s_cursor;
        tup_cnt[blockIdx.x] = min(n, region_cap);
        if (n > region_cap) {      // the region overflowed: the host regrows the record buffer and reruns the epoch
            atomicMax((unsigned long long*)&ctl->need_tuples, (unsigned long long)n * gridDim.x);
            atomicOr(&ctl->overflow, OVF_TUPLES);
        }
    }
}
