// eval.cuh — the insertion search (scheduling.py:13-98) as a warp-cooperative kernel.
//
// One warp owns one compute_schedule() call = one task (vehicle v, base trip, inserted request q).
// Work items are (sub_sche s, pick-up index i) pairs in the reference's loop order, 32 per warp
// step; a lane walks the drop-off index j sequentially with incremental partial sums, which is
// exactly the reference's inner loop and lets its early exits (scheduling.py:41-44) be applied
// per lane (j-loop breaks) and by ballot (viol==3 ends the i-loop of that sub_sche).
// All arithmetic is float64, accumulated left to right in the reference's order, so costs and
// every comparison are bit-identical to the reference.
#pragma once
#include "common.cuh"

#define STAGE_SLOTS 128   // stops staged per warp step: nsub*l <= 32 + 2*l <= 108

struct WarpStage {
    double ddl[STAGE_SLOTS];
    int node[STAGE_SLOTS];
    code_t code[STAGE_SLOTS];
    signed char pod[STAGE_SLOTS];
};

struct LaneOut {
    unsigned long long feas;   // bit j set: insertion (i, j) feasible
    unsigned long long rec;    // bit j set: it was a new running minimum (scheduling.py:31-34)
    double lmin;               // min cost over this lane's feasible insertions
    int evals, lookups;        // as the reference would have executed them
};

// Walk j = i+1 .. for one (sub_sche, i).  `A` = accumulated time at the inserted pick-up.
// CLASSIFY: also mark running-minimum records given the minimum seen before this lane.
template <typename TT, bool CLASSIFY>
__device__ __forceinline__ void walk_j(const Table<TT>& tt, const int* __restrict__ nd, const double* __restrict__ ddl, int l,
                                       int i, double A, int P, int D, double cld, double T, unsigned long long fullmask,
                                       double running, LaneOut& o) {
    // capacity (scheduling.py:74-77): (i,j) is over capacity iff fullmask has a bit in [i, j-1]
    unsigned long long hi = fullmask >> (i + 1);
    int b = hi ? (i + __ffsll((long long)hi)) : 64;    // lowest set bit above i; j <= b is capacity-feasible
    double B = A;
    int nodeB = P;
    for (int j = i + 1;; j++) {
        o.evals++;
        double C = B + tt(nodeB, D);
        if (T + C > cld) { o.lookups += j + 1; break; }            // viol 2 -> break (scheduling.py:41)
        int node = D;
        bool ok = true;
        for (int m = j - 1; m < l; m++) {
            C += tt(node, nd[m]);
            if (T + C > ddl[m]) { ok = false; o.lookups += m + 3; break; }   // viol 0: infeasible, keep going
            node = nd[m];
        }
        if (ok) {
            o.lookups += l + 2;
            o.feas |= 1ull << j;
            if (C < o.lmin) o.lmin = C;
            if (CLASSIFY) { if (C < running) { running = C; o.rec |= 1ull << j; } }
        }
        if (j == l + 1) break;
        if (j == b) { o.evals++; break; }                         // next j is over capacity: viol 1 -> break
        B += tt(nodeB, nd[j - 1]);                                // extend the pick-up side over base stop j-1
        nodeB = nd[j - 1];
        if (T + B > ddl[j - 1]) { o.evals++; o.lookups += j + 1; break; }   // viol 4 at (i, j+1) -> break
    }
}

// MODE 0: count feasible schedules + min cost per task.  MODE 1: materialise them in list order.
template <typename TT, int MODE>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_eval(DEpoch ep, Table<TT> tt, Level prev, const int* __restrict__ len0, int n_tasks,
       const int* __restrict__ task_veh, const int* __restrict__ task_base, const int* __restrict__ task_q,
       int* __restrict__ task_nfeas, double* __restrict__ task_cost,
       const i64* __restrict__ task_trip /* exclusive scan of feasible flags */, Level cur, Stats* stats) {
    __shared__ WarpStage stage_all[WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpStage& st = stage_all[wib];
    const int warp0 = blockIdx.x * WARPS_PER_BLOCK + wib, nwarps = gridDim.x * WARPS_PER_BLOCK;
    unsigned long long acc_evals = 0, acc_lookups = 0;
    const double T = ep.T;

    for (int task = warp0; task < n_tasks; task += nwarps) {
        if (MODE == 1 && task_nfeas[task] == 0) continue;
        const int v = task_veh[task], base = task_base[task], q = task_q[task];
        const int l = len0[v] + 2 * prev.k;
        if (l + 2 > MAXS) {   // compiled stop limit: reported as AMOD_E_LIMIT by the host
            if (lane == 0) { atomicOr(&stats->flags, 1ull); if (MODE == 0) { task_nfeas[task] = 0; task_cost[task] = DINF; } }
            continue;
        }
        const int S = prev.nfeas[base], rot = prev.rot[base];
        const code_t* __restrict__ sch = prev.sch + prev.sch_off[base];
        const int P = ep.onid[q], D = ep.dnid[q];
        const double clp = ep.clp[q], cld = ep.cld[q];
        const int nid = ep.veh_nid[v], load = ep.veh_load[v], cap = ep.cap;
        const double t0 = ep.veh_t[v];
        const int items = S * (l + 1);

        int trip = 0, nfeas_total = 0;
        code_t* __restrict__ out = nullptr;
        if (MODE == 1) {
            trip = (int)task_trip[task];
            nfeas_total = task_nfeas[task];
            out = cur.sch + cur.sch_off[trip];
        }
        double run_min = DINF;
        int n_rec = 0, n_non = 0, n_feas = 0, dead_s = -1;

        for (int it0 = 0; it0 < items; it0 += 32) {
            const int it = it0 + lane;
            const bool valid = it < items;
            const int s = valid ? it / (l + 1) : -1;
            const int i = valid ? it - s * (l + 1) : 0;
            const int s_first = it0 / (l + 1);
            const int last_it = min(items, it0 + 32) - 1;
            const int s_last = last_it / (l + 1);
            __syncwarp();
            if (l > 0) {   // stage the stops of sub_sches s_first..s_last (logical order -> rotated rows)
                const int ne = (s_last - s_first + 1) * l;
                for (int e = lane; e < ne; e += 32) {
                    int ss = s_first + e / l, m = e - (e / l) * l;
                    int row = ss + rot; if (row >= S) row -= S;
                    int code = sch[(i64)row * l + m];
                    int node, pod; double ddl;
                    stop_info(ep, v, code, node, ddl, pod);
                    st.node[e] = node; st.ddl[e] = ddl; st.pod[e] = (signed char)pod; st.code[e] = (code_t)code;
                }
            }
            __syncwarp();
            const int eb = valid ? (s - s_first) * l : 0;
            const int* nd = st.node + eb;
            const double* ddl = st.ddl + eb;

            // capacity profile of the base schedule and accumulated time at the inserted pick-up
            unsigned long long fullmask = 0;
            bool basebad = false;
            double A = 0.0;
            bool stop3 = false, cap1 = false;
            if (valid) {
                int run = load;
                if (run >= cap) fullmask |= 1ull;
                for (int m = 0; m < l; m++) {
                    run += st.pod[eb + m];
                    if (run > cap) basebad = true;
                    if (run >= cap) fullmask |= 1ull << (m + 1);
                }
                double acc = t0;
                int node = nid;
                for (int m = 0; m < i; m++) { acc += tt(node, nd[m]); node = nd[m]; }   // trusted prefix (no checks)
                A = acc + tt(node, P);
                cap1 = basebad || ((fullmask >> i) & 1ull);                              // viol 1 at j = i+1
                stop3 = !cap1 && (T + A > clp);                                          // viol 3 (scheduling.py:43)
            }
            // viol==3 ends the i-loop of its sub_sche: later i of the same s are never evaluated
            const unsigned bal3 = __ballot_sync(FULL, stop3);
            const int lo = max(lane - i, 0);
            const unsigned same_s_lower = (lane > lo) ? (((1u << lane) - 1u) & ~((1u << lo) - 1u)) : 0u;
            const bool dead = valid && ((bal3 & same_s_lower) != 0u || s == dead_s);
            {   // carry to the next warp step if the last sub_sche continues there
                const unsigned killed = __ballot_sync(FULL, valid && s == s_last && (stop3 || dead));
                dead_s = killed ? s_last : -1;
            }

            LaneOut o; o.feas = 0; o.rec = 0; o.lmin = DINF; o.evals = 0; o.lookups = 0;
            const bool live = valid && !dead;
            if (live) {
                if (cap1) { o.evals = 1; }
                else if (stop3) { o.evals = 1; o.lookups = i + 1; }
                else walk_j<TT, false>(tt, nd, ddl, l, i, A, P, D, cld, T, fullmask, 0.0, o);
            }
            if (MODE == 0) {
                acc_evals += o.evals; acc_lookups += o.lookups;
                n_feas += __popcll(o.feas);
                run_min = fmin(run_min, o.lmin);
            } else {
                // records need the minimum over everything evaluated before this lane, in loop order
                const double incoming = fmin(run_min, warp_excl_min(o.lmin, lane));
                LaneOut c; c.feas = 0; c.rec = 0; c.lmin = DINF; c.evals = 0; c.lookups = 0;
                if (live && o.feas) walk_j<TT, true>(tt, nd, ddl, l, i, A, P, D, cld, T, fullmask, incoming, c);
                const int nr = __popcll(c.rec), nn = __popcll(c.feas & ~c.rec);
                int rec_before = n_rec + warp_excl_sum(nr, lane);
                int non_before = n_non + warp_excl_sum(nn, lane);
                unsigned long long f = c.feas;
                while (f) {
                    const int j = __ffsll((long long)f) - 1;
                    f &= f - 1;
                    // memory layout: non-records ascending from row 0, records descending from the last row;
                    // the reference's list order is this rotated by n_non (see Level)
                    const int row = ((c.rec >> j) & 1ull) ? (nfeas_total - 1 - rec_before++) : non_before++;
                    code_t* dst = out + (i64)row * (l + 2);
                    int w = 0;
                    for (int m = 0; m < i; m++) dst[w++] = st.code[eb + m];
                    dst[w++] = (code_t)(2 * q);
                    for (int m = i; m < j - 1; m++) dst[w++] = st.code[eb + m];
                    dst[w++] = (code_t)(2 * q + 1);
                    for (int m = j - 1; m < l; m++) dst[w++] = st.code[eb + m];
                }
                n_rec += warp_sum(nr);
                n_non += warp_sum(nn);
                run_min = fmin(run_min, warp_min(o.lmin));
            }
        }
        if (MODE == 0) {
            n_feas = warp_sum(n_feas);
            run_min = warp_min(run_min);
            if (lane == 0) { task_nfeas[task] = n_feas; task_cost[task] = run_min; }
        } else if (lane == 0) {
            cur.rot[trip] = n_non;
        }
    }
    if (MODE == 0) {
        acc_evals = warp_sum(acc_evals);
        acc_lookups = warp_sum(acc_lookups);
        if (lane == 0 && (acc_evals | acc_lookups)) {
            atomicAdd(&stats->n_evals, acc_evals);
            atomicAdd(&stats->n_lookups, acc_lookups);
        }
    }
}
