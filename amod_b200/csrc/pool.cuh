// pool.cuh — assemble the candidate pool in the reference's order and compute each entry's
// delay (scheduling.py:126-137) in the same pass.
//   OSP  (dispatch_osp.py:107-125): per vehicle: basic pair, trips level by level, current schedule
//   SBA  (dispatch_sba.py:56-69):   feasible (request, vehicle) pairs request-major, then one
//                                   empty pair per vehicle
#pragma once
#include "common.cuh"

struct LevelSet { Level lv[MAX_LEVELS]; int n; };   // lv[0] = basic schedules, lv[k] = trips of size k

#define SRC_INPUT 100   // entry's schedule is the vehicle's current schedule (input arrays)

struct PoolBufs {
    i64 cap_ent, cap_ptr, cap_pst;
    int* ent_veh; int* ent_kind; int* ent_nfeas; int* ent_tlen; int* ent_slen; int* ent_src; int* ent_trip;
    double* ent_cost; double* ent_delay; i64* trip_off; i64* sche_off; int* trip_req; int* sche_code;
};

__global__ void k_pool_veh_count(const DEpoch* __restrict__ epp, LevelSet ls, const Control* __restrict__ ctl, int* __restrict__ cnt) {
    const int n_veh = epp->n_veh, mode = epp->mode;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n_veh; v += gridDim.x * blockDim.x) {
        int c = 1 + (mode == AMOD_MODE_OSP ? 1 : 0);
        if (ctl->overflow) c = 0;
        else if (mode == AMOD_MODE_SBA) c = 1;
        else for (int k = 1; k < ls.n; k++) c += ls.lv[k].veh_start[v + 1] - ls.lv[k].veh_start[v];
        cnt[v] = c;
    }
}

struct InVehEntries {       // scan input: pool entries of vehicle v (dispatch_osp.py:113-125) / one empty pair (dispatch_sba.py:68)
    const DEpoch* ep; LevelSet ls; const Control* ctl;
    __device__ __forceinline__ i64 size() const { return ep->n_veh; }
    __device__ __forceinline__ i64 operator()(i64 v) const {
        if (ctl->overflow) return 0;
        const int mode = ep->mode;
        if (mode == AMOD_MODE_SBA) return 1;
        i64 c = 1 + (mode == AMOD_MODE_OSP ? 1 : 0);
        for (int k = 1; k < ls.n; k++) c += ls.lv[k].veh_start[v + 1] - ls.lv[k].veh_start[v];
        return c;
    }
};

struct FinPoolEntries {   // n_ent = scan total (+ the SBA pairs, which precede the per-vehicle empties)
    Control* ctl; i64 cap; int sba;
    __device__ __forceinline__ void operator()(i64 tot) const {
        if (sba) tot += ctl->n_trips[1];
        if (tot > ctl->need_ent) ctl->need_ent = tot;
        if (tot > cap) { atomicOr(&ctl->overflow, OVF_POOL); tot = 0; }
        ctl->n_ent = tot;
    }
};

// per-vehicle fixed entries (OSP: basic + current; SBA: the empty pair)
__global__ void k_pool_meta_veh(const DEpoch* __restrict__ epp, LevelSet ls, const int* __restrict__ len0, const i64* __restrict__ veh_off,
                                Control* ctl, PoolBufs pb) {
    const DEpoch& ep = *epp;
    if (ctl->n_ent == 0 || ctl->overflow) return;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < ep.n_veh; v += gridDim.x * blockDim.x) {
    const int cur_len = ep.veh_sche_off[v + 1] - ep.veh_sche_off[v];
    // compute_sche_cost walks of the fixed pairs (dispatch_osp.py:125,148, dispatch_sba.py:69)
    atomicAdd(&ctl->stats.n_lookups, (unsigned long long)((ep.mode == AMOD_MODE_SBA ? 0 : len0[v]) +
                                                          (ep.mode == AMOD_MODE_OSP_NR ? 0 : cur_len)));
    if (ep.mode == AMOD_MODE_SBA) {
        i64 e = (i64)ctl->n_trips[1] + v;
        pb.ent_veh[e] = v; pb.ent_kind[e] = AMOD_KIND_SBA_EMPTY; pb.ent_nfeas[e] = 1; pb.ent_tlen[e] = 0; pb.ent_slen[e] = cur_len;
        pb.ent_src[e] = SRC_INPUT; pb.ent_trip[e] = v;
        continue;
    }
    i64 e = veh_off[v];
    pb.ent_veh[e] = v; pb.ent_kind[e] = AMOD_KIND_BASIC; pb.ent_nfeas[e] = ls.lv[0].nfeas[v]; pb.ent_tlen[e] = 0;
    pb.ent_slen[e] = len0[v]; pb.ent_src[e] = 0; pb.ent_trip[e] = v;
    if (ep.mode == AMOD_MODE_OSP) {
        e = veh_off[v + 1] - 1;
        int np = 0;
        for (int k = ep.veh_sche_off[v]; k < ep.veh_sche_off[v + 1]; k++) np += ep.sche_code[k] >= 0 && !(ep.sche_code[k] & 1);
        pb.ent_veh[e] = v; pb.ent_kind[e] = AMOD_KIND_CURRENT; pb.ent_nfeas[e] = 1; pb.ent_tlen[e] = np; pb.ent_slen[e] = cur_len;
        pb.ent_src[e] = SRC_INPUT; pb.ent_trip[e] = v;
    }
    }
}

// trip entries of every level
__global__ void k_pool_meta_levels(const DEpoch* __restrict__ epp, LevelSet ls, const int* __restrict__ len0, const i64* __restrict__ veh_off,
                                   const Control* __restrict__ ctl, PoolBufs pb) {
    const DEpoch& ep = *epp;
    if (ctl->n_ent == 0 || ctl->overflow) return;
    const int stride = gridDim.x * blockDim.x;
    for (int k = 1; k < ls.n; k++) {
        const Level& lv = ls.lv[k];
        const int n = ctl->n_trips[k];
        for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
            const int v = lv.trip_veh[t];
            i64 e;
            if (ep.mode == AMOD_MODE_SBA) e = t;
            else {
                e = veh_off[v] + 1 + (t - lv.veh_start[v]);
                for (int kk = 1; kk < k; kk++) e += ls.lv[kk].veh_start[v + 1] - ls.lv[kk].veh_start[v];
            }
            pb.ent_veh[e] = v; pb.ent_kind[e] = ep.mode == AMOD_MODE_SBA ? AMOD_KIND_SBA_PAIR : AMOD_KIND_TRIP;
            pb.ent_nfeas[e] = lv.nfeas[t]; pb.ent_tlen[e] = k; pb.ent_slen[e] = len0[v] + 2 * k;
            pb.ent_src[e] = k; pb.ent_trip[e] = t;
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
    const DEpoch& ep = *epp;
    const i64 n_ent = ctl->overflow ? 0 : ctl->n_ent;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n_ent; e += stride) {
        const int v = pb.ent_veh[e], src = pb.ent_src[e], t = pb.ent_trip[e];
        const int slen = (int)(pb.sche_off[e + 1] - pb.sche_off[e]);
        int* sc = pb.sche_code + pb.sche_off[e];
        int* tr = pb.trip_req + pb.trip_off[e];
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
            const code_t* in = lv.sch + (s0 + lv.order[s0]) * lv.stride;   // list element 0 = best (scheduling.py:32-34)
            for (int m = 0; m < slen; m++) sc[m] = in[m];
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
