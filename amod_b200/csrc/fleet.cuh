// fleet.cuh — SURVEY.md 8f-3: the vehicles' state advance, the request life cycle and the apply step on the device, so
// that an epoch's state never leaves HBM.
//
//   k_fleet_advance      Veh.move_to_time (vehicle.py:73-162) with update_service_status_after_veh_finish_a_leg, pop_leg,
//                        pop_step, cut_step, jump_to_location (:164-226), and the pick / drop bookkeeping of
//                        Platform.upd_vehs_and_reqs_stat_to_time (platform.py:218-225, request.py:43-57).  One thread per
//                        vehicle: the walk over legs and steps is a sequential float64 recurrence (dT -= leg.t; ...), kept
//                        in the reference's statement order with explicitly rounded products (__dmul_rn: nvcc must not
//                        contract `a - b * c` into an FMA, the reference rounds twice).
//   k_fleet_walkaway     platform.py:227-232
//   k_fleet_apply_*      scheduling.py:110-119 + Veh.build_route / add_leg / clear_route (vehicle.py:231-310): the new
//                        schedule row, the legs (one thread per leg walks the predecessor table, one warp per leg writes the
//                        steps and forms the ordered sums, route_functions.py:36-70), totals and status.
//   k_fleet_considered   dispatch_osp.py:33-37: ids of the PICKING / PENDING requests, ascending.
//
// The vehicle row the pool build reads (FleetState: nid, t_to_nid, load, status, schedule codes, onboard flags,
// rebalancing stop) is updated in place, so amod_fleet_build runs the unchanged pipeline on it.
#pragma once
#undef BFILE
#define BFILE 7
#include "common.cuh"
#include "routes.cuh"
#include "state.cuh"
#include "pool.cuh"

#define FL_IDLE 1          // types.py:18-21
#define FL_WORKING 2
#define FL_REBALANCING 3
#define RQ_PENDING 1       // types.py:11-16
#define RQ_PICKING 2
#define RQ_ONBOARD 3
#define RQ_COMPLETE 4
#define RQ_WALKAWAY 5

#define FLE_SCHE 1ull          // a finished leg does not match the head of the schedule (vehicle.py:170)
#define FLE_LOAD 2ull          // load != number of onboard drop-offs (vehicle.py:182)
#define FLE_END 4ull           // route exhausted with stops / load left (vehicle.py:155-156)
#define FLE_PICK 8ull          // picked a request that was not PICKING (request.py:49)
#define FLE_DROP 16ull         // dropped a request that was not ONBOARD (request.py:55)
#define FLE_PATH 32ull         // the path table does not lead back to a leg's origin
#define FLE_REB 64ull          // vehicle.py:236 / :285
#define FLE_CAPSUM 128ull      // vehicle.py:281
#define FLE_ARG 256ull         // vehicle index / schedule length / stop code out of range
#define FLE_STEPS 512ull       // a route needs more than max_steps steps
#define FLE_EVENTS 1024ull

struct FleetSim {
    int n_veh, max_stops, max_legs, max_steps;
    double *lng, *lat, *d_to_nid, *state_time, *tot_t, *tot_d;
    double* stats;                       // [8][n_veh]: Ds Ts Ds_empty Ts_empty Dr Tr Lt Ld (vehicle.py:57-64)
    int* tnid;
    uint8_t* has_step; double *cut_t, *cut_d; int *cut_u, *cut_v;      // step_to_nid (vehicle.py:216)
    int *leg_head, *leg_n;
    int *leg_rid, *leg_pod, *leg_tnid, *leg_o, *leg_sb, *leg_se; double *leg_ddl, *leg_t, *leg_d;      // [n_veh * max_legs]
    double *st_t, *st_d; int *st_u, *st_v;                             // [n_veh * max_steps]
    int* rq_status; double *rq_tp, *rq_td;
    const int *rq_onid, *rq_dnid; const double *rq_clp, *rq_cld, *rq_tr;
    int* ev_n; int *ev_veh, *ev_seq, *ev_rid, *ev_pod; double* ev_time; int ev_cap;
    const double *node_lng, *node_lat;
    unsigned long long* err;             // [0] error mask, [1] steps a route needed, [2] longest schedule applied
};

__device__ __forceinline__ double f_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double f_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double f_sub(double a, double b) { return __dsub_rn(a, b); }

struct VehRegs {          // the fields move_to_time rewrites, in registers for the walk
    int nid, load, status; double lng, lat, state_time, t, d;
    double Ds, Ts, Ds_empty, Ts_empty, Dr, Tr, Lt, Ld;
};

// vehicle.py:89-98 (and :107-116): statistics of a finished leg / step
__device__ __forceinline__ void fl_stat(VehRegs& r, int upd, double t, double d) {
    if (!upd) return;
    r.Ts = f_add(r.Ts, t); r.Ds = f_add(r.Ds, d);
    r.Lt = f_add(r.Lt, f_mul(t, (double)r.load)); r.Ld = f_add(r.Ld, f_mul(d, (double)r.load));
    if (r.status == FL_WORKING && r.load == 0) { r.Ts_empty = f_add(r.Ts_empty, t); r.Ds_empty = f_add(r.Ds_empty, d); }
    if (r.status == FL_REBALANCING) { r.Tr = f_add(r.Tr, t); r.Dr = f_add(r.Dr, d); }
}

// vehicle.py:164-191: update_service_status_after_veh_finish_a_leg + pop_leg
__device__ __forceinline__ void fl_finish_leg(const FleetSim& s, const FleetState& fs, int v, VehRegs& r, int lg, int& seq) {
    const i64 L = (i64)v * s.max_legs + lg;
    const int tn = s.leg_tnid[L], rid = s.leg_rid[L], pod = s.leg_pod[L];
    r.nid = tn; r.lng = s.node_lng[tn - 1]; r.lat = s.node_lat[tn - 1];
    r.load += pod;
    if (rid != -2) {
        const int slot = atomicAdd(s.ev_n, 1);
        if (slot < s.ev_cap) { s.ev_veh[slot] = v; s.ev_seq[slot] = seq; s.ev_rid[slot] = rid; s.ev_pod[slot] = pod; s.ev_time[slot] = r.state_time; }
        else atomicOr(s.err, FLE_EVENTS);
        seq++;
        int* code = fs.code + (i64)v * fs.max_stops;
        uint8_t* onb = fs.onboard + (i64)v * fs.max_stops;
        const int n = fs.len[v];
        const int want = rid < 0 ? -1 : 2 * rid + (pod == -1 ? 1 : 0);
        if (n <= 0 || code[0] != want) atomicOr(s.err, FLE_SCHE);
        for (int m = 0; m + 1 < n; m++) { code[m] = code[m + 1]; onb[m] = onb[m + 1]; }       // self.sche.pop(0)
        if (n > 0) { code[n - 1] = 0; onb[n - 1] = 0; fs.len[v] = n - 1; }
        if (pod == 1) {                  // picked up: its drop-off is now an onboard stop; request.py:43-50
            for (int m = 0; m + 1 < n; m++) if (code[m] == 2 * rid + 1) onb[m] = 1;
            s.rq_tp[rid] = r.state_time;
            if (s.rq_status[rid] != RQ_PICKING) atomicOr(s.err, FLE_PICK);
            s.rq_status[rid] = RQ_ONBOARD;
        } else if (pod == -1) {          // request.py:52-57
            s.rq_td[rid] = r.state_time;
            if (s.rq_status[rid] != RQ_ONBOARD) atomicOr(s.err, FLE_DROP);
            s.rq_status[rid] = RQ_COMPLETE;
        }
    }
    r.d = f_sub(r.d, s.leg_d[L]); r.t = f_sub(r.t, s.leg_t[L]);        // pop_leg
}

__global__ void __launch_bounds__(128) k_fleet_advance(FleetSim s, FleetState fs, double T, int upd) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < s.n_veh; v += gridDim.x * blockDim.x) {
        double dT = f_sub(T, s.state_time[v]);
        if (dT <= 0) continue;
        VehRegs r;
        r.nid = fs.nid[v]; r.load = fs.load[v]; r.status = fs.status[v]; r.lng = s.lng[v]; r.lat = s.lat[v];
        r.state_time = s.state_time[v]; r.t = s.tot_t[v]; r.d = s.tot_d[v];
        double* S = s.stats;
        const i64 V = s.n_veh;
        r.Ds = S[0 * V + v]; r.Ts = S[1 * V + v]; r.Ds_empty = S[2 * V + v]; r.Ts_empty = S[3 * V + v];
        r.Dr = S[4 * V + v]; r.Tr = S[5 * V + v]; r.Lt = S[6 * V + v]; r.Ld = S[7 * V + v];
        double t_to = 0.0, d_to = 0.0;
        int has_step = 0, seq = 0;
        int head = s.leg_head[v];
        const int nl = s.leg_n[v];
        bool stopped = false;
        while (dT > 0 && head < nl && !stopped) {
            const i64 L = (i64)v * s.max_legs + head;
            const double lt = s.leg_t[L];
            if (lt < dT) {
                dT = f_sub(dT, lt); r.state_time = f_add(r.state_time, lt);
                fl_stat(r, upd, lt, s.leg_d[L]);
                fl_finish_leg(s, fs, v, r, head, seq);
                head++;
            } else {
                int sb = s.leg_sb[L];
                const int se = s.leg_se[L];
                double leg_t = lt, leg_d = s.leg_d[L];
                bool leg_done = false;
                while (dT > 0 && sb < se) {
                    const i64 P = (i64)v * s.max_steps + sb;
                    const double st = s.st_t[P], sd = s.st_d[P];
                    if (st < dT) {
                        dT = f_sub(dT, st); r.state_time = f_add(r.state_time, st);
                        fl_stat(r, upd, st, sd);
                        const int to = s.st_v[P];
                        r.nid = to; r.lng = s.node_lng[to - 1]; r.lat = s.node_lat[to - 1];
                        r.t = f_sub(r.t, st); r.d = f_sub(r.d, sd); leg_t = f_sub(leg_t, st); leg_d = f_sub(leg_d, sd);     // pop_step
                        sb++;
                        if (sb == se) {          // corner case vehicle.py:121-125
                            s.leg_t[L] = leg_t; s.leg_d[L] = leg_d; s.leg_sb[L] = sb;
                            fl_finish_leg(s, fs, v, r, head, seq);
                            head++;
                            leg_done = true;
                            break;
                        }
                    } else {
                        const double pct = __ddiv_rn(dT, st);
                        if (upd) {               // vehicle.py:129-140
                            const double dp = f_mul(sd, pct);
                            r.Ts = f_add(r.Ts, dT); r.Ds = f_add(r.Ds, dp);
                            r.Lt = f_add(r.Lt, f_mul(dT, (double)r.load)); r.Ld = f_add(r.Ld, f_mul(dp, (double)r.load));
                            if (r.status == FL_WORKING && r.load == 0) { r.Ts_empty = f_add(r.Ts_empty, dT); r.Ds_empty = f_add(r.Ds_empty, dp); }
                            if (r.status == FL_REBALANCING) { r.Tr = f_add(r.Tr, dT); r.Dr = f_add(r.Dr, dp); }
                        }
                        // cut_step (vehicle.py:202-217); geo_pair[0] of the head step is the vehicle's own position
                        const int to = s.st_v[P];
                        r.lng = f_add(r.lng, f_mul(pct, f_sub(s.node_lng[to - 1], r.lng)));
                        r.lat = f_add(r.lat, f_mul(pct, f_sub(s.node_lat[to - 1], r.lat)));
                        const double rest = f_sub(1.0, pct);
                        t_to = f_mul(st, rest); d_to = f_mul(sd, rest);
                        const double tp = f_mul(st, pct), dp = f_mul(sd, pct);
                        r.t = f_sub(r.t, tp); r.d = f_sub(r.d, dp);
                        leg_t = f_sub(leg_t, tp); leg_d = f_sub(leg_d, dp);
                        const double nst = f_sub(st, tp), nsd = f_sub(sd, dp);
                        s.st_t[P] = nst; s.st_d[P] = nsd; s.st_u[P] = to;
                        s.cut_t[v] = nst; s.cut_d[v] = nsd; s.cut_u[v] = to; s.cut_v[v] = to;
                        has_step = 1;
                        r.nid = to;
                        r.state_time = T;
                        stopped = true;
                        break;
                    }
                }
                if (!leg_done) { s.leg_t[L] = leg_t; s.leg_d[L] = leg_d; s.leg_sb[L] = sb; }
            }
        }
        if (!stopped) {          // vehicle.py:152-162: the whole schedule is finished
            if (head != nl || fs.len[v] != 0 || r.load != 0) atomicOr(s.err, FLE_END);
            head = 0; s.leg_n[v] = 0;
            r.state_time = T; r.d = 0.0; r.t = 0.0; r.status = FL_IDLE;
        }
        s.leg_head[v] = head;
        fs.nid[v] = r.nid; fs.load[v] = r.load; fs.status[v] = r.status; fs.t[v] = t_to;
        s.d_to_nid[v] = d_to; s.has_step[v] = (uint8_t)has_step;
        if (!has_step) { s.cut_t[v] = 0.0; s.cut_d[v] = 0.0; s.cut_u[v] = 0; s.cut_v[v] = 0; }
        s.lng[v] = r.lng; s.lat[v] = r.lat; s.state_time[v] = r.state_time; s.tot_t[v] = r.t; s.tot_d[v] = r.d;
        S[0 * V + v] = r.Ds; S[1 * V + v] = r.Ts; S[2 * V + v] = r.Ds_empty; S[3 * V + v] = r.Ts_empty;
        S[4 * V + v] = r.Dr; S[5 * V + v] = r.Tr; S[6 * V + v] = r.Lt; S[7 * V + v] = r.Ld;
        // load == len(onboard_rids) (vehicle.py:182): the onboard drop-offs left in the schedule
        int onb = 0;
        for (int m = 0; m < fs.len[v]; m++) onb += fs.onboard[(i64)v * fs.max_stops + m];
        if (onb != r.load) atomicOr(s.err, FLE_LOAD);
    }
}

// platform.py:227-232
__global__ void k_fleet_walkaway(FleetSim s, int n_req, double T) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_req; r += gridDim.x * blockDim.x) {
        if (s.rq_status[r] != RQ_PENDING) continue;
        if (T >= f_add(s.rq_tr[r], 150.0) || T >= s.rq_clp[r]) s.rq_status[r] = RQ_WALKAWAY;
    }
}

__global__ void k_fleet_fill_requests(FleetSim s, int first, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        s.rq_status[first + i] = RQ_PENDING; s.rq_tp[first + i] = -1.0; s.rq_td[first + i] = -1.0;       // request.py:27,38-39
    }
}

__global__ void k_fleet_set_status(FleetSim s, const int* __restrict__ rids, int n, int status, int n_req) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = rids[i];
        if (r < 0 || r >= n_req) { atomicOr(s.err, FLE_ARG); continue; }
        s.rq_status[r] = status;
    }
}

// dispatch_osp.py:77-78: every considered (PICKING or PENDING) request goes back to PENDING before the selection is applied
__global__ void k_fleet_reset_picking(FleetSim s, int n_req) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_req; r += gridDim.x * blockDim.x)
        if (s.rq_status[r] == RQ_PICKING) s.rq_status[r] = RQ_PENDING;
}

// dispatch_osp.py:33-37
struct InConsidered {
    const int* status; i64 n;
    __device__ __forceinline__ i64 size() const { return n; }
    __device__ __forceinline__ i64 operator()(i64 r) const { const int st = status[r]; return (st == RQ_PICKING || st == RQ_PENDING) ? 1 : 0; }
};
__global__ void k_fleet_considered(FleetSim s, int n_req, const i64* __restrict__ pos, int* __restrict__ out) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_req; r += gridDim.x * blockDim.x) {
        const int st = s.rq_status[r];
        if (st == RQ_PICKING || st == RQ_PENDING) out[pos[r]] = r;
    }
}

// ---- apply step ---------------------------------------------------------------------------------------------------
// Where the new schedules come from: arrays the caller staged (NPO matches, an external assigner), or entries of the pool
// the last build left on the device (the greedy selection: entry = order[sel[i]]).
struct ApplySrc {
    int n, from_pool, mark_picking;
    const int* veh; const int* off; const int* code; const int* reb_node; const double* reb_ddl;
    const int* entries; const int* sel; const int* order;
    PoolBufs pb;
};

// one thread per call: the schedule row, the legs' (rid, pod, tnid, ddl, origin), the requests of the trip -> PICKING
__global__ void k_fleet_apply_rows(FleetSim s, FleetState fs, ApplySrc a, int n_req, int* __restrict__ call_veh) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        int v, n; const int* codes; int rn = 0; double rd = 0.0; i64 e = -1;
        if (a.from_pool) {
            e = a.entries ? a.entries[i] : a.order[a.sel[i]];
            v = a.pb.ent_veh[e];
            const i64 b = a.pb.sche_off[e];
            n = (int)(a.pb.sche_off[e + 1] - b); codes = a.pb.sche_code + b;
        } else {
            v = a.veh[i]; n = a.off[i + 1] - a.off[i]; codes = a.code + a.off[i]; rn = a.reb_node[i]; rd = a.reb_ddl[i];
        }
        call_veh[i] = -1;
        if (v < 0 || v >= s.n_veh || n < 0 || n > fs.max_stops) { atomicOr(s.err, FLE_ARG); continue; }
        if (a.from_pool) { rn = fs.reb_node[v]; rd = fs.reb_ddl[v]; }      // a pool schedule can only carry the vehicle's own rebalancing stop
        // vehicle.py:234-241: a rebalancing vehicle that receives a trip drops its rebalancing stop
        int skip = -1;
        if (fs.status[v] == FL_REBALANCING && n > 1) {
            if (fs.len[v] != 1) atomicOr(s.err, FLE_REB);
            for (int m = 0; m < n; m++) if (codes[m] == -1) { skip = m; break; }
        }
        int* row = fs.code + (i64)v * fs.max_stops;
        uint8_t* onb = fs.onboard + (i64)v * fs.max_stops;
        int len = 0, has_reb = 0;
        int tmp[AMOD_MAX_STOPS];
        for (int m = 0; m < n; m++) {
            if (m == skip) continue;
            const int c = codes[m];
            if (c < -1 || (c >= 0 && (c >> 1) >= n_req)) { atomicOr(s.err, FLE_ARG); continue; }
            tmp[len++] = c;
            has_reb |= c == -1;
        }
        for (int m = 0; m < fs.max_stops; m++) {
            int c = 0, ob = 0;
            if (m < len) {
                c = tmp[m];
                if (c >= 0 && (c & 1)) {         // a drop-off whose pick-up is not in the schedule: the request is on board
                    ob = 1;
                    for (int q = 0; q < len; q++) if (tmp[q] == c - 1) ob = 0;
                }
            }
            row[m] = c; onb[m] = (uint8_t)ob;
        }
        fs.len[v] = len;
        if (has_reb) { fs.reb_node[v] = rn; fs.reb_ddl[v] = rd; }
        atomicMax(s.err + 2, (unsigned long long)len);
        // clear_route (vehicle.py:304-310) and the legs of build_route (:246-270)
        const int nid = fs.nid[v];
        int o = nid, L = 0;
        const i64 base = (i64)v * s.max_legs;
        if (s.has_step[v]) {             // the unfinished step of the last move becomes a leg of its own (rid -2)
            s.leg_rid[base] = -2; s.leg_pod[base] = 0; s.leg_tnid[base] = s.cut_v[v]; s.leg_ddl[base] = f_add(s.state_time[v], s.cut_t[v]);
            s.leg_o[base] = 0; L = 1; o = s.cut_v[v];
        }
        int cap_sum = fs.load[v];
        for (int m = 0; m < len; m++, L++) {
            const int c = tmp[m];
            int rid, pod, tn; double ddl;
            if (c < 0) { rid = -1; pod = 0; tn = rn; ddl = rd; }
            else if (c & 1) { rid = c >> 1; pod = -1; tn = s.rq_dnid[rid]; ddl = s.rq_cld[rid]; }
            else { rid = c >> 1; pod = 1; tn = s.rq_onid[rid]; ddl = s.rq_clp[rid]; }
            s.leg_rid[base + L] = rid; s.leg_pod[base + L] = pod; s.leg_tnid[base + L] = tn; s.leg_ddl[base + L] = ddl; s.leg_o[base + L] = o;
            o = tn; cap_sum += pod;
        }
        s.leg_head[v] = 0; s.leg_n[v] = L; s.tnid[v] = o;
        // status (vehicle.py:274-288)
        int status = FL_IDLE;
        if (len > 0) {
            if (tmp[0] >= 0) { status = FL_WORKING; if (cap_sum != 0) atomicOr(s.err, FLE_CAPSUM); }
            else { status = FL_REBALANCING; if (len != 1) atomicOr(s.err, FLE_REB); }
        }
        fs.status[v] = status;
        if (a.from_pool && a.mark_picking) {     // scheduling.py:116-117
            for (i64 q = a.pb.trip_off[e]; q < a.pb.trip_off[e + 1]; q++) s.rq_status[a.pb.trip_req[q]] = RQ_PICKING;
        }
        call_veh[i] = v;
    }
}

// one thread per (call, leg): steps of the leg = nodes on the path (the closing zero-length step included)
__global__ void k_fleet_route_count(FleetSim s, RouteTables rt, const int* __restrict__ call_veh, int n_calls) {
    const i64 total = (i64)n_calls * s.max_legs;
    for (i64 g = blockIdx.x * (i64)blockDim.x + threadIdx.x; g < total; g += (i64)gridDim.x * blockDim.x) {
        const int v = call_veh[g / s.max_legs], k = (int)(g % s.max_legs);
        if (v < 0 || k >= s.leg_n[v]) continue;
        const i64 L = (i64)v * s.max_legs + k;
        int len = 2;
        if (s.leg_rid[L] != -2) {
            const int o = s.leg_o[L], d = s.leg_tnid[L];
            len = (o >= 1 && o <= rt.n && d >= 1 && d <= rt.n) ? route_len(rt, o, d) : 0;
            if (len == 0) { atomicOr(s.err, FLE_PATH); len = 1; }
        }
        s.leg_se[L] = len;               // (count, turned into a range by k_fleet_route_offsets)
    }
}

__global__ void k_fleet_route_offsets(FleetSim s, int* __restrict__ call_veh, int n_calls) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_calls; i += gridDim.x * blockDim.x) {
        const int v = call_veh[i];
        if (v < 0) continue;
        const i64 base = (i64)v * s.max_legs;
        const int nl = s.leg_n[v];
        i64 at = 0;
        for (int k = 0; k < nl; k++) at += s.leg_se[base + k];
        atomicMax(s.err + 1, (unsigned long long)at);
        if (at > s.max_steps) {          // the route does not fit: the vehicle keeps no route; the host reports AMOD_E_CAPACITY
            atomicOr(s.err, FLE_STEPS);
            s.leg_n[v] = 0; call_veh[i] = -1;
            continue;
        }
        int p = 0;
        for (int k = 0; k < nl; k++) { const int c = s.leg_se[base + k]; s.leg_sb[base + k] = p; p += c; s.leg_se[base + k] = p; }
    }
}

// one warp per (call, leg): route_functions.py:36-70 into the vehicle's step region
template <typename TT>
__global__ void __launch_bounds__(256) k_fleet_route_fill(FleetSim s, RouteTables rt, Table<TT> tt, const int* __restrict__ call_veh, int n_calls) {
    const int lane = threadIdx.x & 31;
    const i64 nwarps = (i64)gridDim.x * (blockDim.x >> 5);
    const i64 total = (i64)n_calls * s.max_legs;
    const TT* dist = (const TT*)rt.dist;
    for (i64 g = blockIdx.x * (i64)(blockDim.x >> 5) + (threadIdx.x >> 5); g < total; g += nwarps) {
        const int v = call_veh[g / s.max_legs], k = (int)(g % s.max_legs);
        if (v < 0 || k >= s.leg_n[v]) continue;
        const i64 L = (i64)v * s.max_legs + k;
        const i64 a = (i64)v * s.max_steps + s.leg_sb[L];
        const int len = s.leg_se[L] - s.leg_sb[L];
        if (s.leg_rid[L] == -2) {        // [copy of step_to_nid, closing step] (vehicle.py:255-257)
            if (lane == 0) {
                const int to = s.cut_v[v];
                s.st_t[a] = s.cut_t[v]; s.st_d[a] = s.cut_d[v]; s.st_u[a] = s.cut_u[v]; s.st_v[a] = to;
                s.st_t[a + 1] = 0.0; s.st_d[a + 1] = 0.0; s.st_u[a + 1] = to; s.st_v[a + 1] = to;
                s.leg_t[L] = s.cut_t[v]; s.leg_d[L] = s.cut_d[v];
            }
            continue;
        }
        const int o = s.leg_o[L], d = s.leg_tnid[L];
        if (lane == 0) {
            const int* row = rt.pred + (size_t)(o - 1) * (size_t)rt.n;
            int x = d;
            for (int p = len - 1; p >= 0; p--) { s.st_u[a + p] = x; x = p > 0 ? row[x - 1] : x; }      // path.reverse()
        }
        __syncwarp();
        for (int p = lane; p < len; p += 32) {
            const int u = s.st_u[a + p];
            const int w = p + 1 < len ? s.st_u[a + p + 1] : u;
            double t = 0.0, dd = 0.0;
            if (p + 1 < len) { t = tt(u, w); dd = (double)__ldg(dist + (size_t)(u - 1) * (size_t)rt.n + (size_t)(w - 1)); }
            s.st_t[a + p] = t; s.st_d[a + p] = dd; s.st_v[a + p] = w;
        }
        __syncwarp();
        if (lane == 0) {
            double st = 0.0, sd = 0.0;   // duration += t; distance += d (route_functions.py:56-57)
            for (int p = 0; p + 1 < len; p++) { st = f_add(st, s.st_t[a + p]); sd = f_add(sd, s.st_d[a + p]); }
            s.leg_t[L] = st; s.leg_d[L] = sd;
        }
        __syncwarp();
    }
}

// one thread per call: self.d += leg.d; self.t += leg.t in leg order (vehicle.py:265-266, 301-302)
__global__ void k_fleet_apply_totals(FleetSim s, const int* __restrict__ call_veh, int n_calls) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_calls; i += gridDim.x * blockDim.x) {
        const int v = call_veh[i];
        if (v < 0) continue;
        const i64 base = (i64)v * s.max_legs;
        double t = 0.0, d = 0.0;
        for (int k = 0; k < s.leg_n[v]; k++) { d = f_add(d, s.leg_d[base + k]); t = f_add(t, s.leg_t[base + k]); }
        s.tot_t[v] = t; s.tot_d[v] = d;
    }
}
