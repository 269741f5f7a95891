// pool.cuh — assemble the candidate pool in the reference's order and compute each entry's
// delay (scheduling.py:126-137) in the same pass.
//   OSP  (dispatch_osp.py:107-125): per vehicle: basic pair, trips level by level, current schedule
//   SBA  (dispatch_sba.py:56-69):   feasible (request, vehicle) pairs request-major, then one
//                                   empty pair per vehicle
#pragma once
#undef BFILE
#define BFILE 4
#include "common.cuh"
#include "scan.cuh"

struct LevelSet { Level lv[MAX_LEVELS]; int n; };   // lv[0] = basic schedules, lv[k] = trips of size k

#define SRC_INPUT 100   // entry's schedule is the vehicle's current schedule (input arrays)

struct PoolBufs {
    i64 cap_ent, cap_ptr, cap_pst;
    int* ent_veh; int* ent_kind; int* ent_nfeas; int* ent_tlen; int* ent_slen; int* ent_src; int* ent_trip;
    double* ent_cost; double* ent_delay; i64* trip_off; i64* sche_off; int* trip_req; int* sche_code;
};

// First pool entry of vehicle v.  The per-level veh_start arrays already are prefix sums over vehicles, so the
// vehicle-major entry offsets need no scan: (fixed entries per vehicle) * v + sum over levels of veh_start_k[v].
__device__ __forceinline__ i64 pool_veh_off(const LevelSet& ls, int mode, int v) {
    if (mode == AMOD_MODE_SBA) return v;
    i64 e = (i64)v * (1 + (mode == AMOD_MODE_OSP ? 1 : 0));
    for (int k = 1; k < ls.n; k++) e += ls.lv[k].veh_start[v];
    return e;
}

// adds (tlen, slen) of one entry per lane to the block sums of the two scans over the entries; one atomic pair
// per warp when all its entries fall into the same scan segment (the common case: lanes hold consecutive entries)
__device__ __forceinline__ void pool_len_partials(unsigned long long* __restrict__ part, int vg, i64 seg, bool valid, i64 e, int tlen, int slen,
                                                  int lane) {
    const unsigned vm = __ballot_sync(FULL, valid);
    if (!vm) return;
    const int sg = valid ? (int)(e / seg) : -1;
    const int lead = __ffs(vm) - 1;
    const int sg0 = __shfl_sync(FULL, sg, lead);
    const bool uni = __all_sync(FULL, !valid || sg == sg0);
    if (uni) {
        const int t = warp_sum(valid ? tlen : 0), sl = warp_sum(valid ? slen : 0);
        if (lane == lead) { atomicAdd(&part[sg0], (unsigned long long)t); atomicAdd(&part[vg + sg0], (unsigned long long)sl); }
    } else if (valid) {
        atomicAdd(&part[sg], (unsigned long long)tlen); atomicAdd(&part[vg + sg], (unsigned long long)slen);
    }
}

// Entry metadata: the per-vehicle fixed entries (OSP: basic + current; SBA: the empty pair) and the trip entries
// of every level; publishes the entry count, the per-vehicle offsets (multi-GPU assembly reads them) and the block
// sums of the two scans that follow (trip lengths, schedule lengths).
__global__ void __launch_bounds__(256)
k_pool_meta(const DEpoch* __restrict__ epp, LevelSet ls, const int* __restrict__ len0, i64* __restrict__ veh_off, Control* ctl, PoolBufs pb,
            unsigned long long* __restrict__ part, int part_vg) {
    pdl_enter();
    const DEpoch& ep = *epp;
    const int mode = ep.mode, n_veh = ep.n_veh;
    const bool sba = mode == AMOD_MODE_SBA;
    const int lane = threadIdx.x & 31;
    i64 n_ent = 0;
    if (!ctl->overflow) n_ent = sba ? (i64)ctl->n_trips[1] + n_veh : pool_veh_off(ls, mode, n_veh);
    const bool over = n_ent > pb.cap_ent;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (n_ent > ctl->need_ent) ctl->need_ent = n_ent;
        if (over) atomicOr(&ctl->overflow, OVF_POOL);
        ctl->n_ent = over ? 0 : n_ent;
    }
    if (over || n_ent == 0) return;
    const i64 seg = scan_seg_len(n_ent, part_vg);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    unsigned long long look = 0;      // compute_sche_cost walks of the fixed pairs (dispatch_osp.py:125,148, dispatch_sba.py:69)
    int max_np = 0;
    for (int base = blockIdx.x * blockDim.x; base <= n_veh; base += stride) {      // (whole warps iterate together: one atomic per warp and sum)
        const int v = base + threadIdx.x;
        const bool valid = v < n_veh;
        i64 e0 = 0, e1 = 0;
        int cur_len = 0, l0 = 0, np = 0;
        if (v <= n_veh) { e0 = pool_veh_off(ls, mode, v); veh_off[v] = e0; }
        if (valid) {
            const int c0 = ep.veh_sche_off[v], c1 = ep.veh_sche_off[v + 1];
            cur_len = c1 - c0; l0 = len0[v];
            look += (unsigned long long)((sba ? 0 : l0) + (mode == AMOD_MODE_OSP_NR ? 0 : cur_len));
            if (sba) {
                const i64 e = (i64)ctl->n_trips[1] + v;
                e0 = e;
                pb.ent_veh[e] = v; pb.ent_kind[e] = AMOD_KIND_SBA_EMPTY; pb.ent_nfeas[e] = 1; pb.ent_tlen[e] = 0; pb.ent_slen[e] = cur_len;
                pb.ent_src[e] = SRC_INPUT; pb.ent_trip[e] = v;
            } else {
                const i64 e = e0;
                pb.ent_veh[e] = v; pb.ent_kind[e] = AMOD_KIND_BASIC; pb.ent_nfeas[e] = ls.lv[0].nfeas[v]; pb.ent_tlen[e] = 0;
                pb.ent_slen[e] = l0; pb.ent_src[e] = 0; pb.ent_trip[e] = v;
                if (mode == AMOD_MODE_OSP) {
                    e1 = pool_veh_off(ls, mode, v + 1) - 1;
                    for (int k = c0; k < c1; k++) np += ep.sche_code[k] >= 0 && !(ep.sche_code[k] & 1);
                    pb.ent_veh[e1] = v; pb.ent_kind[e1] = AMOD_KIND_CURRENT; pb.ent_nfeas[e1] = 1; pb.ent_tlen[e1] = np; pb.ent_slen[e1] = cur_len;
                    pb.ent_src[e1] = SRC_INPUT; pb.ent_trip[e1] = v;
                }
            }
        }
        max_np = max(max_np, np);
        pool_len_partials(part, part_vg, seg, valid, e0, 0, sba ? cur_len : l0, lane);                  // the SBA empty pair / the basic pair
        pool_len_partials(part, part_vg, seg, valid && mode == AMOD_MODE_OSP, e1, np, cur_len, lane);   // the current schedule's pair
    }
    look = warp_sum(look);
    for (int o = 16; o; o >>= 1) max_np = max(max_np, __shfl_xor_sync(FULL, max_np, o));
    if (lane == 0) {
        if (look) atomicAdd(&ctl->stats.n_lookups, look);
        if (max_np > 0) atomicMax(&ctl->max_tlen, max_np);
    }
    for (int k = 1; k < ls.n; k++) {
        const Level& lv = ls.lv[k];
        const int n = ctl->n_trips[k];
        if (n > 0 && gtid == 0) atomicMax(&ctl->max_tlen, k);
        for (int base = blockIdx.x * blockDim.x; base < n; base += stride) {      // (whole warps iterate together)
            const int t = base + threadIdx.x;
            const bool valid = t < n;
            i64 e = 0;
            int slen = 0;
            if (valid) {
                const int v = lv.trip_veh[t];
                if (sba) e = t;
                else {
                    e = pool_veh_off(ls, mode, v) + 1 + (t - lv.veh_start[v]);
                    for (int kk = 1; kk < k; kk++) e += ls.lv[kk].veh_start[v + 1] - ls.lv[kk].veh_start[v];
                }
                slen = len0[v] + 2 * k;
                if (!BOK(e >= 0 && e < pb.cap_ent)) { e = 0; slen = 0; }
                else {
                pb.ent_veh[e] = v; pb.ent_kind[e] = sba ? AMOD_KIND_SBA_PAIR : AMOD_KIND_TRIP;
                pb.ent_nfeas[e] = lv.nfeas[t]; pb.ent_tlen[e] = k; pb.ent_slen[e] = slen;
                pb.ent_src[e] = k; pb.ent_trip[e] = t;
                }
            }
            pool_len_partials(part, part_vg, seg, valid, e, k, slen, lane);
        }
    }
}

struct InEntLen {   // scan input over pool entries
    const int* p; const Control* ctl;
    __device__ __forceinline__ i64 size() const { return ctl->n_ent; }
    __device__ __forceinline__ i64 operator()(i64 i) const { return (i64)p[i]; }
};

#define POOL_BATCH 8      // stops of one entry whose look-ups are in flight together

// One thread per entry.  An entry's walk is a chain of dependent round trips (stop -> node -> travel time), so the stops are
// taken POOL_BATCH at a time: all their codes / nodes first (a stored schedule carries its nodes), then all their travel
// times and request times, and only then the left-to-right float64 sums (scheduling.py:101-107, :126-137) from registers.
template <typename TT>
__global__ void k_pool_payload(const DEpoch* __restrict__ epp, Table<TT> tt, LevelSet ls, const Control* __restrict__ ctl, PoolBufs pb) {
    pdl_enter();
    const DEpoch& ep = *epp;
    const i64 n_ent = ctl->overflow ? 0 : ctl->n_ent;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n_ent; e += stride) {
        const int v = pb.ent_veh[e], src = pb.ent_src[e], t = pb.ent_trip[e];
        const i64 so0 = pb.sche_off[e], so1 = pb.sche_off[e + 1], to0 = pb.trip_off[e], to1 = pb.trip_off[e + 1];
        const int slen = (int)(so1 - so0);
        int* sc = pb.sche_code + so0;
        int* tr = pb.trip_req + to0;
        if (!BOK(e < pb.cap_ent && so1 <= pb.cap_pst && to1 <= pb.cap_ptr && so0 >= 0 && to0 >= 0)) continue;
        const bool input = src == SRC_INPUT;
        const Level& lv = ls.lv[input ? 0 : src];
        const Stop* __restrict__ in = nullptr;       // rows are in list order: element 0 = best (scheduling.py:32-34)
        const int* __restrict__ inc = nullptr;
        if (input) inc = ep.sche_code + ep.veh_sche_off[v];
        else in = lv.sch + lv.s0[t] * lv.stride;
        double acc = ep.veh_t[v], delay = 0.0;
        int nid = ep.veh_nid[v];
        const double cost_stored = (input || src == 0) ? 0.0 : lv.cost[t];
        if (!input)
            for (int m0 = 0; m0 < src; m0 += 4) {
                int q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) if (m0 + u < src) q[u] = lv.trip_req[(i64)t * src + m0 + u];
#pragma unroll
                for (int u = 0; u < 4; u++) if (m0 + u < src) tr[m0 + u] = q[u];
            }
        int w = 0;
        for (int m0 = 0; m0 < slen; m0 += POOL_BATCH) {
            int code[POOL_BATCH], node[POOL_BATCH];
#pragma unroll
            for (int u = 0; u < POOL_BATCH; u++) {
                code[u] = -1; node[u] = 1;
                if (m0 + u < slen) {
                    if (input) code[u] = inc[m0 + u];
                    else { const Stop sp = in[m0 + u]; code[u] = sp.code; node[u] = sp.node; }
                }
            }
            if (input) {
#pragma unroll
                for (int u = 0; u < POOL_BATCH; u++)
                    if (m0 + u < slen) node[u] = code[u] < 0 ? ep.reb_node[v] : ((code[u] & 1) ? ep.dnid[code[u] >> 1] : ep.onid[code[u] >> 1]);
            }
            TT leg[POOL_BATCH];
            double rtr[POOL_BATCH], rts[POOL_BATCH];
#pragma unroll
            for (int u = 0; u < POOL_BATCH; u++) {
                leg[u] = (TT)0; rtr[u] = 0.0; rts[u] = 0.0;
                if (m0 + u < slen) {
                    leg[u] = tt.raw(u ? node[u - 1] : nid, node[u]);
                    if (code[u] >= 0 && (code[u] & 1)) { rtr[u] = ep.tr[code[u] >> 1]; rts[u] = ep.ts[code[u] >> 1]; }
                }
            }
#pragma unroll
            for (int u = 0; u < POOL_BATCH; u++) {
                if (m0 + u < slen) {
                    sc[m0 + u] = code[u];
                    acc += (double)leg[u];
                    if (code[u] >= 0 && (code[u] & 1)) delay += (ep.T + acc - rtr[u] - rts[u]);
                    nid = node[u];
                    if (input && to1 > to0 && code[u] >= 0 && !(code[u] & 1)) {     // picking requests, ascending (dispatch_osp.py:121-124)
                        const int r = code[u] >> 1;
                        int x = w++;
                        while (x > 0 && tr[x - 1] > r) { tr[x] = tr[x - 1]; x--; }
                        tr[x] = r;
                    }
                }
            }
        }
        pb.ent_cost[e] = (input || src == 0) ? acc : cost_stored;
        pb.ent_delay[e] = delay;
    }
}
