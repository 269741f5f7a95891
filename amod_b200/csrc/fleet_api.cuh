// fleet_api.cuh — C-ABI of the device-resident fleet (include/amod_b200.h amod_fleet_*; kernels in fleet.cuh).
// Included by amod_b200.cu after the resident-state functions (it builds on amod_state_* and the routes tables).
#pragma once

static int fleet_check(amod_ctx* ctx, const char* what) {
    unsigned long long h[3] = {0, 0, 0};
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(h, ctx->fl.err, 24, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    ctx->fs_max_len = std::max(ctx->fs_max_len, (int)h[2]);          // (monotone: keeps the SBA / OSP-NR graph and strides put)
    if (h[0]) CK(cudaMemsetAsync(ctx->fl.err, 0, 8, st));
    if (h[0] & FLE_STEPS)
        return ctx->fail(AMOD_E_CAPACITY, "%s: a route needs %llu steps, the fleet was created with max_route_steps = %d (that vehicle's route was dropped: recreate the fleet)",
                         what, h[1], ctx->fl.max_steps);
    if (h[0] & FLE_ARG) return ctx->fail(AMOD_E_BADARG, "%s: a vehicle index, schedule length, stop code or request id is out of range", what);
    if (h[0] & FLE_PATH) return ctx->fail(AMOD_E_BADARG, "%s: the path table does not lead from a leg's destination back to its origin", what);
    if (h[0]) return ctx->fail(AMOD_E_BADARG, "%s: the fleet state violates an invariant the reference asserts (mask 0x%llx: 1 leg/schedule mismatch, "
                               "2 load, 4 route exhausted with stops left, 8 pick-up of a request not PICKING, 16 drop-off of a request not ONBOARD, "
                               "64 rebalancing schedule, 128 capacity sum, 1024 event buffer)", what, h[0]);
    return AMOD_OK;
}

extern "C" int amod_fleet_create(amod_ctx* ctx, int32_t n_veh, int32_t max_stops, int32_t max_route_steps, int32_t request_capacity,
                                 const int32_t* start_nid, const double* node_lng, const double* node_lat) {
    if (!ctx || n_veh < 1 || max_route_steps < 4 || !start_nid || !node_lng || !node_lat) return AMOD_E_BADARG;
    if (!ctx->d_pred) return ctx->fail(AMOD_E_BADARG, "amod_routes_upload first (the apply step builds routes from the path table)");
    for (int v = 0; v < n_veh; v++)
        if (start_nid[v] < 1 || start_nid[v] > ctx->n_nodes) return ctx->fail(AMOD_E_BADARG, "vehicle %d: start node out of range", v);
    CKR(amod_state_create(ctx, n_veh, max_stops, request_capacity));
    cudaStream_t st = ctx->stream;
    const size_t V = (size_t)n_veh, ML = (size_t)max_stops + 1, MS = (size_t)max_route_steps, R = (size_t)request_capacity, N = (size_t)ctx->n_nodes;
    FleetSim& f = ctx->fl;
    memset(&f, 0, sizeof f);
    f.n_veh = n_veh; f.max_stops = max_stops; f.max_legs = (int)ML; f.max_steps = max_route_steps;
    for (auto& b : ctx->fl_bufs) b.release();
    ctx->fl_bufs.clear(); ctx->fl_bufs.reserve(64);
    bool ok = true;
    auto alloc = [&](size_t bytes) -> void* {
        ctx->fl_bufs.emplace_back();
        DevBuf& b = ctx->fl_bufs.back();
        if (b.ensure(std::max(bytes, (size_t)16)) != cudaSuccess) { ok = false; return nullptr; }
        cudaMemsetAsync(b.p, 0, b.cap, st);
        return b.p;
    };
    f.lng = (double*)alloc(8 * V); f.lat = (double*)alloc(8 * V); f.d_to_nid = (double*)alloc(8 * V); f.state_time = (double*)alloc(8 * V);
    f.tot_t = (double*)alloc(8 * V); f.tot_d = (double*)alloc(8 * V); f.stats = (double*)alloc(8 * 8 * V); f.tnid = (int*)alloc(4 * V);
    f.has_step = (uint8_t*)alloc(V); f.cut_t = (double*)alloc(8 * V); f.cut_d = (double*)alloc(8 * V); f.cut_u = (int*)alloc(4 * V); f.cut_v = (int*)alloc(4 * V);
    f.leg_head = (int*)alloc(4 * V); f.leg_n = (int*)alloc(4 * V);
    f.leg_rid = (int*)alloc(4 * V * ML); f.leg_pod = (int*)alloc(4 * V * ML); f.leg_tnid = (int*)alloc(4 * V * ML); f.leg_o = (int*)alloc(4 * V * ML);
    f.leg_sb = (int*)alloc(4 * V * ML); f.leg_se = (int*)alloc(4 * V * ML);
    f.leg_ddl = (double*)alloc(8 * V * ML); f.leg_t = (double*)alloc(8 * V * ML); f.leg_d = (double*)alloc(8 * V * ML);
    f.st_t = (double*)alloc(8 * V * MS); f.st_d = (double*)alloc(8 * V * MS); f.st_u = (int*)alloc(4 * V * MS); f.st_v = (int*)alloc(4 * V * MS);
    f.rq_status = (int*)alloc(4 * R); f.rq_tp = (double*)alloc(8 * R); f.rq_td = (double*)alloc(8 * R);
    f.ev_cap = (int)(V * ML);
    f.ev_n = (int*)alloc(16); f.ev_veh = (int*)alloc(4 * V * ML); f.ev_seq = (int*)alloc(4 * V * ML); f.ev_rid = (int*)alloc(4 * V * ML);
    f.ev_pod = (int*)alloc(4 * V * ML); f.ev_time = (double*)alloc(8 * V * ML);
    double* geo = (double*)alloc(16 * N);
    f.err = (unsigned long long*)alloc(32);
    if (!ok) return ctx->fail(AMOD_E_CUDA, "out of device memory for the fleet state (%zu vehicles x %zu route steps)", V, MS);
    f.node_lng = geo; f.node_lat = geo + N;
    f.rq_onid = ctx->rq_onid.as<int>(); f.rq_dnid = ctx->rq_dnid.as<int>(); f.rq_clp = ctx->rq_clp.as<double>(); f.rq_cld = ctx->rq_cld.as<double>();
    f.rq_tr = ctx->rq_tr.as<double>();
    CK(cudaMemcpyAsync(geo, node_lng, 8 * N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(geo + N, node_lat, 8 * N, cudaMemcpyHostToDevice, st));
    // vehicle.py:41-70: every vehicle idle at its station
    std::vector<double> lng(V), lat(V);
    std::vector<int> one(V, FL_IDLE);
    for (size_t v = 0; v < V; v++) { lng[v] = node_lng[start_nid[v] - 1]; lat[v] = node_lat[start_nid[v] - 1]; }
    CK(cudaMemcpyAsync(ctx->fs.nid, start_nid, 4 * V, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(f.tnid, start_nid, 4 * V, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->fs.status, one.data(), 4 * V, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(f.lng, lng.data(), 8 * V, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(f.lat, lat.data(), 8 * V, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    ctx->fl_valid = true; ctx->fl_n_events = 0;
    return AMOD_OK;
}

#define FLEET_READY() do { if (!ctx) return AMOD_E_BADARG; if (!ctx->fl_valid) return ctx->fail(AMOD_E_BADARG, "amod_fleet_create first"); \
                           CK(cudaSetDevice(ctx->device)); } while (0)

extern "C" int amod_fleet_add_requests(amod_ctx* ctx, int32_t first, int32_t n, const int32_t* onid, const int32_t* dnid, const double* clp,
                                       const double* cld, const double* tr, const double* ts) {
    FLEET_READY();
    if (n <= 0) return n == 0 ? AMOD_OK : AMOD_E_BADARG;
    if (first != ctx->fs_req_n) return ctx->fail(AMOD_E_BADARG, "requests arrive in id order: expected first id %d", ctx->fs_req_n);
    CKR(amod_state_add_requests(ctx, first, n, onid, dnid, clp, cld, tr, ts));
    k_fleet_fill_requests<<<std::max(1, std::min(ctx->n_sm * 4, (n + 255) / 256)), 256, 0, ctx->stream>>>(ctx->fl, first, n);
    CK(cudaGetLastError());
    return AMOD_OK;
}

// Platform.upd_vehs_and_reqs_stat_to_time (platform.py:213-233)
extern "C" int amod_fleet_advance(amod_ctx* ctx, int64_t T, int32_t update_stats, int32_t* n_events) {
    FLEET_READY();
    cudaStream_t st = ctx->stream;
    const FleetSim& f = ctx->fl;
    CK(cudaMemsetAsync(f.ev_n, 0, 4, st));
    k_fleet_advance<<<std::max(1, std::min(ctx->n_sm * 8, (f.n_veh + 127) / 128)), 128, 0, st>>>(f, ctx->fs, (double)T, update_stats ? 1 : 0);
    if (ctx->fs_req_n) k_fleet_walkaway<<<std::max(1, std::min(ctx->n_sm * 8, (ctx->fs_req_n + 255) / 256)), 256, 0, st>>>(f, ctx->fs_req_n, (double)T);
    int n = 0;
    CK(cudaMemcpyAsync(&n, f.ev_n, 4, cudaMemcpyDeviceToHost, st));
    CKR(fleet_check(ctx, "amod_fleet_advance"));
    ctx->fl_n_events = std::min(n, f.ev_cap);
    if (n_events) *n_events = ctx->fl_n_events;
    return AMOD_OK;
}

// the (vehicle, sequence within the vehicle, rid, pod, time) tuples move_to_time returned in the last advance, in no particular
// order (sort by (veh, seq) for the reference's order)
extern "C" int amod_fleet_events(amod_ctx* ctx, int32_t* veh, int32_t* seq, int32_t* rid, int32_t* pod, double* time, int32_t cap) {
    FLEET_READY();
    const size_t n = (size_t)ctx->fl_n_events;
    if ((size_t)cap < n) return ctx->fail(AMOD_E_CAPACITY, "%zu events", n);
    if (!n) return AMOD_OK;
    cudaStream_t st = ctx->stream;
    const FleetSim& f = ctx->fl;
    CK(cudaMemcpyAsync(veh, f.ev_veh, 4 * n, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(seq, f.ev_seq, 4 * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(rid, f.ev_rid, 4 * n, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(pod, f.ev_pod, 4 * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(time, f.ev_time, 8 * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return AMOD_OK;
}

extern "C" int amod_fleet_set_status(amod_ctx* ctx, const int32_t* rids, int32_t n, int32_t status) {
    FLEET_READY();
    if (n < 0 || (n && !rids) || status < RQ_PENDING || status > RQ_WALKAWAY) return AMOD_E_BADARG;
    if (!n) return AMOD_OK;
    cudaStream_t st = ctx->stream;
    CK(ctx->fl_rids.ensure(4 * (size_t)n));
    CK(cudaMemcpyAsync(ctx->fl_rids.p, rids, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    k_fleet_set_status<<<std::max(1, std::min(ctx->n_sm * 4, (n + 255) / 256)), 256, 0, st>>>(ctx->fl, ctx->fl_rids.as<int>(), n, status, ctx->fs_req_n);
    return fleet_check(ctx, "amod_fleet_set_status");
}

template <typename TT>
static int fleet_apply_impl(amod_ctx* ctx, ApplySrc a, Table<TT> tt) {
    cudaStream_t st = ctx->stream;
    const FleetSim& f = ctx->fl;
    const int n = a.n;
    CK(ctx->fl_call.ensure(4 * (size_t)n));
    int* call = ctx->fl_call.as<int>();
    RouteTables rt; rt.pred = ctx->d_pred; rt.dist = ctx->d_dist; rt.n = ctx->n_nodes;
    const i64 work = (i64)n * f.max_legs;
    const int g1 = std::max(1, std::min(ctx->n_sm * 8, (n + 127) / 128));
    k_fleet_apply_rows<<<g1, 128, 0, st>>>(f, ctx->fs, a, ctx->fs_req_n, call);
    k_fleet_route_count<<<(int)std::max<i64>(1, std::min<i64>(ctx->n_sm * 8, (work + 127) / 128)), 128, 0, st>>>(f, rt, call, n);
    k_fleet_route_offsets<<<g1, 128, 0, st>>>(f, call, n);
    k_fleet_route_fill<TT><<<(int)std::max<i64>(1, std::min<i64>(ctx->n_sm * 8, (work + 7) / 8)), 256, 0, st>>>(f, rt, tt, call, n);
    k_fleet_apply_totals<<<g1, 128, 0, st>>>(f, call, n);
    return fleet_check(ctx, "amod_fleet_apply");
}

// Veh.build_route (vehicle.py:231-288) for a batch: call i gives vehicle veh[i] the stop codes [sche_off[i], sche_off[i+1])
// (2*rid + drop-off, -1 = the rebalancing stop (reb_node[i], reb_ddl[i])).  A vehicle appears at most once per call.
extern "C" int amod_fleet_apply(amod_ctx* ctx, int32_t n, const int32_t* veh, const int32_t* sche_off, const int32_t* sche_code,
                                const int32_t* reb_node, const double* reb_ddl) {
    FLEET_READY();
    if (n < 0 || (n && (!veh || !sche_off || !reb_node || !reb_ddl))) return AMOD_E_BADARG;
    if (!n) return AMOD_OK;
    cudaStream_t st = ctx->stream;
    const size_t N = (size_t)n, NS = (size_t)sche_off[n];
    if (NS && !sche_code) return AMOD_E_BADARG;
    for (int i = 0; i < n; i++) {
        if (sche_off[i + 1] < sche_off[i]) return ctx->fail(AMOD_E_BADARG, "call %d: negative schedule length", i);
        for (int k = sche_off[i]; k < sche_off[i + 1]; k++)
            if (sche_code[k] == -1 && (reb_node[i] < 1 || reb_node[i] > ctx->n_nodes)) return ctx->fail(AMOD_E_BADARG, "call %d: rebalancing node out of range", i);
    }
    // one packed staging buffer: reb_ddl | veh | off | reb_node | code
    const size_t o_veh = 8 * N, o_off = o_veh + 4 * N, o_reb = o_off + 4 * (N + 1), o_code = o_reb + 4 * N, total = o_code + 4 * NS;
    std::vector<char> host(total);
    memcpy(host.data(), reb_ddl, 8 * N); memcpy(host.data() + o_veh, veh, 4 * N); memcpy(host.data() + o_off, sche_off, 4 * (N + 1));
    memcpy(host.data() + o_reb, reb_node, 4 * N); if (NS) memcpy(host.data() + o_code, sche_code, 4 * NS);
    CK(ctx->fl_stage.ensure(total));
    CK(cudaMemcpyAsync(ctx->fl_stage.p, host.data(), total, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    const char* B = ctx->fl_stage.as<char>();
    ApplySrc a; memset(&a, 0, sizeof a);
    a.n = n; a.from_pool = 0; a.reb_ddl = (const double*)B; a.veh = (const int*)(B + o_veh); a.off = (const int*)(B + o_off);
    a.reb_node = (const int*)(B + o_reb); a.code = (const int*)(B + o_code);
    return dispatch_tt(ctx, [&](auto tt) { return fleet_apply_impl(ctx, a, tt); });
}

// scheduling.py:110-119 on the pool the last amod_fleet_build left on the device: entries[i] (host; pool entry indices) or, with
// entries == NULL, the selection of the last amod_greedy_assign (still on the device).  reset_considered: dispatch_osp.py:77-78.
extern "C" int amod_fleet_apply_selected(amod_ctx* ctx, const int32_t* entries, int32_t n, int32_t mark_picking, int32_t reset_considered) {
    FLEET_READY();
    if (!ctx->pool_valid) return ctx->fail(AMOD_E_BADARG, "no pool has been built on this context");
    cudaStream_t st = ctx->stream;
    ApplySrc a; memset(&a, 0, sizeof a);
    a.from_pool = 1; a.mark_picking = mark_picking ? 1 : 0; a.pb = pool_bufs_of(ctx);
    if (entries) {
        if (n < 0) return AMOD_E_BADARG;
        for (int i = 0; i < n; i++) if (entries[i] < 0 || entries[i] >= ctx->pool_entries) return ctx->fail(AMOD_E_BADARG, "entries[%d] is not a pool entry", i);
        CK(ctx->fl_stage.ensure(4 * (size_t)std::max(n, 1)));
        if (n) CK(cudaMemcpyAsync(ctx->fl_stage.p, entries, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        a.entries = ctx->fl_stage.as<int>(); a.n = n;
    } else {
        if (ctx->ga_n_sel < 0) return ctx->fail(AMOD_E_BADARG, "amod_greedy_assign first (or pass the entries)");
        a.sel = ctx->ga_sel.as<int>(); a.order = ctx->ga_order.as<int>(); a.n = ctx->ga_n_sel;
    }
    if (reset_considered && ctx->fs_req_n)
        k_fleet_reset_picking<<<std::max(1, std::min(ctx->n_sm * 8, (ctx->fs_req_n + 255) / 256)), 256, 0, st>>>(ctx->fl, ctx->fs_req_n);
    if (!a.n) return fleet_check(ctx, "amod_fleet_apply_selected");
    return dispatch_tt(ctx, [&](auto tt) { return fleet_apply_impl(ctx, a, tt); });
}

// The pool build of one epoch from the fleet state: OSP searches the considered requests (PICKING or PENDING, ascending ids,
// dispatch_osp.py:33-37, found on the device); SBA and OSP-NR search the n_new requests that arrived last (platform.py:140-150).
extern "C" int amod_fleet_build(amod_ctx* ctx, int32_t mode, int32_t cap, int64_t T, int32_t n_new, amod_stats* stats, int32_t* n_search_out) {
    FLEET_READY();
    cudaStream_t st = ctx->stream;
    const int R = ctx->fs_req_n;
    if (n_new < 0 || n_new > R) return AMOD_E_BADARG;
    int n_search = 0;
    CK(ctx->fs_search.ensure(4 * (size_t)std::max(R, 1)));
    if (mode == AMOD_MODE_OSP) {
        if (R) {
            CK(ctx->fl_pos.ensure(8 * ((size_t)R + 2)));
            CK(ctx->partial.ensure(2 * 8 * (size_t)(ctx->n_sm + 1)));
            scan_dev(ctx, InConsidered{ctx->fl.rq_status, (i64)R}, ctx->fl_pos.as<i64>(), FinNone{});
            k_fleet_considered<<<std::max(1, std::min(ctx->n_sm * 8, (R + 255) / 256)), 256, 0, st>>>(ctx->fl, R, ctx->fl_pos.as<i64>(), ctx->fs_search.as<int>());
            i64 tot = 0;
            CK(cudaMemcpyAsync(&tot, ctx->fl_pos.as<i64>() + R, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            n_search = (int)tot;
        }
    } else {
        std::vector<int> ids((size_t)n_new);
        for (int i = 0; i < n_new; i++) ids[(size_t)i] = R - n_new + i;
        if (n_new) CK(cudaMemcpyAsync(ctx->fs_search.p, ids.data(), 4 * (size_t)n_new, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        n_search = n_new;
    }
    if (n_search_out) *n_search_out = n_search;
    return state_build_dev(ctx, mode, cap, T, n_search, stats);
}

// Download one array of the fleet state (host mirror, tests).  `which` = AMOD_FLEET_* (include/amod_b200.h); bytes must equal the
// array's size.
extern "C" int amod_fleet_array(amod_ctx* ctx, int32_t which, void* out, int64_t bytes) {
    FLEET_READY();
    const FleetSim& f = ctx->fl; const FleetState& fs = ctx->fs;
    const size_t V = (size_t)f.n_veh, S = (size_t)f.max_stops, ML = (size_t)f.max_legs, R = (size_t)ctx->fs_req_n;
    const void* src = nullptr; size_t sz = 0;
    switch (which) {
        case AMOD_FLEET_NID: src = fs.nid; sz = 4 * V; break;
        case AMOD_FLEET_LNG: src = f.lng; sz = 8 * V; break;
        case AMOD_FLEET_LAT: src = f.lat; sz = 8 * V; break;
        case AMOD_FLEET_T_TO_NID: src = fs.t; sz = 8 * V; break;
        case AMOD_FLEET_D_TO_NID: src = f.d_to_nid; sz = 8 * V; break;
        case AMOD_FLEET_TNID: src = f.tnid; sz = 4 * V; break;
        case AMOD_FLEET_LOAD: src = fs.load; sz = 4 * V; break;
        case AMOD_FLEET_STATUS: src = fs.status; sz = 4 * V; break;
        case AMOD_FLEET_STATE_TIME: src = f.state_time; sz = 8 * V; break;
        case AMOD_FLEET_T: src = f.tot_t; sz = 8 * V; break;
        case AMOD_FLEET_D: src = f.tot_d; sz = 8 * V; break;
        case AMOD_FLEET_STATS: src = f.stats; sz = 64 * V; break;
        case AMOD_FLEET_HAS_STEP: src = f.has_step; sz = V; break;
        case AMOD_FLEET_STEP_T: src = f.cut_t; sz = 8 * V; break;
        case AMOD_FLEET_STEP_D: src = f.cut_d; sz = 8 * V; break;
        case AMOD_FLEET_STEP_U: src = f.cut_u; sz = 4 * V; break;
        case AMOD_FLEET_STEP_V: src = f.cut_v; sz = 4 * V; break;
        case AMOD_FLEET_SCHE_LEN: src = fs.len; sz = 4 * V; break;
        case AMOD_FLEET_SCHE_CODE: src = fs.code; sz = 4 * V * S; break;
        case AMOD_FLEET_SCHE_ONBOARD: src = fs.onboard; sz = V * S; break;
        case AMOD_FLEET_REB_NODE: src = fs.reb_node; sz = 4 * V; break;
        case AMOD_FLEET_REB_DDL: src = fs.reb_ddl; sz = 8 * V; break;
        case AMOD_FLEET_LEG_HEAD: src = f.leg_head; sz = 4 * V; break;
        case AMOD_FLEET_LEG_N: src = f.leg_n; sz = 4 * V; break;
        case AMOD_FLEET_LEG_RID: src = f.leg_rid; sz = 4 * V * ML; break;
        case AMOD_FLEET_LEG_POD: src = f.leg_pod; sz = 4 * V * ML; break;
        case AMOD_FLEET_LEG_TNID: src = f.leg_tnid; sz = 4 * V * ML; break;
        case AMOD_FLEET_LEG_DDL: src = f.leg_ddl; sz = 8 * V * ML; break;
        case AMOD_FLEET_LEG_T: src = f.leg_t; sz = 8 * V * ML; break;
        case AMOD_FLEET_LEG_D: src = f.leg_d; sz = 8 * V * ML; break;
        case AMOD_FLEET_LEG_SB: src = f.leg_sb; sz = 4 * V * ML; break;
        case AMOD_FLEET_LEG_SE: src = f.leg_se; sz = 4 * V * ML; break;
        case AMOD_FLEET_REQ_STATUS: src = f.rq_status; sz = 4 * R; break;
        case AMOD_FLEET_REQ_TP: src = f.rq_tp; sz = 8 * R; break;
        case AMOD_FLEET_REQ_TD: src = f.rq_td; sz = 8 * R; break;
        default: return ctx->fail(AMOD_E_BADARG, "unknown fleet array %d", which);
    }
    if ((size_t)bytes != sz) return ctx->fail(AMOD_E_BADARG, "fleet array %d holds %zu bytes, the caller asked for %lld", which, sz, (long long)bytes);
    if (!sz) return AMOD_OK;
    if (!out) return AMOD_E_BADARG;
    CK(cudaMemcpyAsync(out, src, sz, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return AMOD_OK;
}
