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
    for (int v = gtid; v <= n_veh; v += stride) {
        const i64 e0 = pool_veh_off(ls, mode, v);
        veh_off[v] = e0;
        if (v == n_veh) break;
        const int cur_len = ep.veh_sche_off[v + 1] - ep.veh_sche_off[v];
        // compute_sche_cost walks of the fixed pairs (dispatch_osp.py:125,148, dispatch_sba.py:69)
        atomicAdd(&ctl->stats.n_lookups, (unsigned long long)((sba ? 0 : len0[v]) + (mode == AMOD_MODE_OSP_NR ? 0 : cur_len)));
        if (sba) {
            const i64 e = (i64)ctl->n_trips[1] + v;
            pb.ent_veh[e] = v; pb.ent_kind[e] = AMOD_KIND_SBA_EMPTY; pb.ent_nfeas[e] = 1; pb.ent_tlen[e] = 0; pb.ent_slen[e] = cur_len;
            pb.ent_src[e] = SRC_INPUT; pb.ent_trip[e] = v;
            atomicAdd(&part[part_vg + e / seg], (unsigned long long)cur_len);
            continue;
        }
        i64 e = e0;
        pb.ent_veh[e] = v; pb.ent_kind[e] = AMOD_KIND_BASIC; pb.ent_nfeas[e] = ls.lv[0].nfeas[v]; pb.ent_tlen[e] = 0;
        pb.ent_slen[e] = len0[v]; pb.ent_src[e] = 0; pb.ent_trip[e] = v;
        atomicAdd(&part[part_vg + e / seg], (unsigned long long)len0[v]);
        if (mode == AMOD_MODE_OSP) {
            e = pool_veh_off(ls, mode, v + 1) - 1;
            int np = 0;
            for (int k = ep.veh_sche_off[v]; k < ep.veh_sche_off[v + 1]; k++) np += ep.sche_code[k] >= 0 && !(ep.sche_code[k] & 1);
            pb.ent_veh[e] = v; pb.ent_kind[e] = AMOD_KIND_CURRENT; pb.ent_nfeas[e] = 1; pb.ent_tlen[e] = np; pb.ent_slen[e] = cur_len;
            if (np > ctl->max_tlen) atomicMax(&ctl->max_tlen, np);
            pb.ent_src[e] = SRC_INPUT; pb.ent_trip[e] = v;
            atomicAdd(&part[e / seg], (unsigned long long)np); atomicAdd(&part[part_vg + e / seg], (unsigned long long)cur_len);
        }
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

template <typename TT>
__global__ void k_pool_payload(const DEpoch* __restrict__ epp, Table<TT> tt, LevelSet ls, const Control* __restrict__ ctl, PoolBufs pb) {
    pdl_enter();
    const DEpoch& ep = *epp;
    const i64 n_ent = ctl->overflow ? 0 : ctl->n_ent;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n_ent; e += stride) {
        const int v = pb.ent_veh[e], src = pb.ent_src[e], t = pb.ent_trip[e];
        const int slen = (int)(pb.sche_off[e + 1] - pb.sche_off[e]);
        int* sc = pb.sche_code + pb.sche_off[e];
        int* tr = pb.trip_req + pb.trip_off[e];
        if (!BOK(e < pb.cap_ent && pb.sche_off[e + 1] <= pb.cap_pst && pb.trip_off[e + 1] <= pb.cap_ptr && pb.sche_off[e] >= 0 && pb.trip_off[e] >= 0)) continue;
        if (src == SRC_INPUT) {
            const int* in = ep.sche_code + ep.veh_sche_off[v];
            int w = 0;
            for (int m = 0; m < slen; m++) sc[m] = in[m];
            if (pb.trip_off[e + 1] > pb.trip_off[e]) {     // picking requests, ascending (dispatch_osp.py:121-124)
                for (int m = 0; m < slen; m++) if (in[m] >= 0 && !(in[m] & 1)) {
                    int r = in[m] >> 1, x = w++;
                    while (x > 0 && tr[x - 1] > r) { tr[x] = tr[x - 1]; x--; }
                    tr[x] = r;
                }
            }
        } else {
            const Level& lv = ls.lv[src];
            const i64 s0 = lv.s0[t];
            const Stop* in = lv.sch + s0 * lv.stride;   // rows are in list order: element 0 = best (scheduling.py:32-34)
            for (int m = 0; m < slen; m++) sc[m] = in[m].code;
            for (int m = 0; m < src; m++) tr[m] = lv.trip_req[(i64)t * src + m];
        }
        // cost walk (scheduling.py:101-107) and delay (:126-137)
        double acc = ep.veh_t[v], delay = 0.0;
        int nid = ep.veh_nid[v];
        for (int m = 0; m < slen; m++) {
            int node, pod; double ddl;
            stop_info(ep, v, sc[m], node, ddl, pod);
            acc += tt(nid, node);
            if (pod == -1) { const int r = sc[m] >> 1; delay += (ep.T + acc - ep.tr[r] - ep.ts[r]); }
            nid = node;
        }
        pb.ent_cost[e] = (src == SRC_INPUT || src == 0) ? acc : ls.lv[src].cost[t];
        pb.ent_delay[e] = delay;
    }
}
