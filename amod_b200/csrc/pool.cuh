// pool.cuh — assemble the candidate pool in the reference's order and compute each entry's
// delay (scheduling.py:126-137) in the same pass.
//   OSP  (dispatch_osp.py:107-125): per vehicle: basic pair, trips level by level, current schedule
//   SBA  (dispatch_sba.py:56-69):   feasible (request, vehicle) pairs request-major, then one
//                                   empty pair per vehicle
#pragma once
#include "common.cuh"

#define MAX_LEVELS (AMOD_MAX_TRIP + 1)

struct LevelSet { Level lv[MAX_LEVELS]; int n; };   // lv[0] = basic schedules, lv[k] = trips of size k

// source of an entry's schedule
#define SRC_INPUT 100   // the vehicle's current schedule (input arrays)

__global__ void k_pool_veh_count(int n_veh, int mode, LevelSet ls, int* __restrict__ cnt) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_veh) return;
    int c = 1 + (mode == AMOD_MODE_OSP ? 1 : 0);
    for (int k = 1; k < ls.n; k++) c += ls.lv[k].veh_start[v + 1] - ls.lv[k].veh_start[v];
    cnt[v] = c;
}

// per-vehicle fixed entries (OSP: basic + current; SBA: the empty pair)
__global__ void k_pool_meta_veh(DEpoch ep, LevelSet ls, const int* __restrict__ len0, const i64* __restrict__ veh_off,
                                i64 sba_base, int* ent_veh, int* ent_kind, int* ent_nfeas, int* ent_tlen, int* ent_slen,
                                int* ent_src, int* ent_trip, Stats* stats) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= ep.n_veh) return;
    const int cur_len = ep.veh_sche_off[v + 1] - ep.veh_sche_off[v];
    // compute_sche_cost walks of the fixed pairs (dispatch_osp.py:125,148, dispatch_sba.py:69)
    atomicAdd(&stats->n_lookups, (unsigned long long)((ep.mode == AMOD_MODE_SBA ? 0 : len0[v]) +
                                                      (ep.mode == AMOD_MODE_OSP_NR ? 0 : cur_len)));
    if (ep.mode == AMOD_MODE_SBA) {
        i64 e = sba_base + v;
        ent_veh[e] = v; ent_kind[e] = AMOD_KIND_SBA_EMPTY; ent_nfeas[e] = 1; ent_tlen[e] = 0; ent_slen[e] = cur_len;
        ent_src[e] = SRC_INPUT; ent_trip[e] = v;
        return;
    }
    i64 e = veh_off[v];
    ent_veh[e] = v; ent_kind[e] = AMOD_KIND_BASIC; ent_nfeas[e] = ls.lv[0].nfeas[v]; ent_tlen[e] = 0; ent_slen[e] = len0[v];
    ent_src[e] = 0; ent_trip[e] = v;
    if (ep.mode == AMOD_MODE_OSP) {
        e = veh_off[v + 1] - 1;
        int np = 0;
        for (int k = ep.veh_sche_off[v]; k < ep.veh_sche_off[v + 1]; k++) np += ep.sche_code[k] >= 0 && !(ep.sche_code[k] & 1);
        ent_veh[e] = v; ent_kind[e] = AMOD_KIND_CURRENT; ent_nfeas[e] = 1; ent_tlen[e] = np; ent_slen[e] = cur_len;
        ent_src[e] = SRC_INPUT; ent_trip[e] = v;
    }
}

// trip entries of one level
__global__ void k_pool_meta_level(DEpoch ep, LevelSet ls, int k, const int* __restrict__ len0, const i64* __restrict__ veh_off,
                                  int* ent_veh, int* ent_kind, int* ent_nfeas, int* ent_tlen, int* ent_slen, int* ent_src,
                                  int* ent_trip) {
    const Level& lv = ls.lv[k];
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= lv.n_trips) return;
    const int v = lv.trip_veh[t];
    i64 e;
    if (ep.mode == AMOD_MODE_SBA) e = t;
    else {
        e = veh_off[v] + 1 + (t - lv.veh_start[v]);
        for (int kk = 1; kk < k; kk++) e += ls.lv[kk].veh_start[v + 1] - ls.lv[kk].veh_start[v];
    }
    ent_veh[e] = v; ent_kind[e] = ep.mode == AMOD_MODE_SBA ? AMOD_KIND_SBA_PAIR : AMOD_KIND_TRIP;
    ent_nfeas[e] = lv.nfeas[t]; ent_tlen[e] = k; ent_slen[e] = len0[v] + 2 * k;
    ent_src[e] = k; ent_trip[e] = t;
}

template <typename TT>
__global__ void k_pool_payload(DEpoch ep, Table<TT> tt, LevelSet ls, i64 n_ent, const int* __restrict__ ent_veh,
                               const int* __restrict__ ent_src, const int* __restrict__ ent_trip,
                               const i64* __restrict__ trip_off, const i64* __restrict__ sche_off, int* __restrict__ trip_req,
                               int* __restrict__ sche_code, double* __restrict__ ent_cost, double* __restrict__ ent_delay) {
    i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_ent) return;
    const int v = ent_veh[e], src = ent_src[e], t = ent_trip[e];
    const int slen = (int)(sche_off[e + 1] - sche_off[e]);
    int* sc = sche_code + sche_off[e];
    int* tr = trip_req + trip_off[e];
    if (src == SRC_INPUT) {
        const int* in = ep.sche_code + ep.veh_sche_off[v];
        int w = 0;
        for (int m = 0; m < slen; m++) { sc[m] = in[m]; }
        if (trip_off[e + 1] > trip_off[e]) {       // picking requests, ascending (dispatch_osp.py:121-124)
            for (int m = 0; m < slen; m++) if (in[m] >= 0 && !(in[m] & 1)) {
                int r = in[m] >> 1, x = w++;
                while (x > 0 && tr[x - 1] > r) { tr[x] = tr[x - 1]; x--; }
                tr[x] = r;
            }
        }
    } else {
        const Level& lv = ls.lv[src];
        const code_t* in = lv.sch + lv.sch_off[t] + (i64)lv.rot[t] * slen;   // list element 0 = best (scheduling.py:32-34)
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
    ent_cost[e] = (src == SRC_INPUT || src == 0) ? acc : ls.lv[src].cost[t];
    ent_delay[e] = delay;
}
