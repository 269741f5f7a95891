"""Record whole simulated horizons at the reference's seams (north-star target run): every epoch's
seam INPUTS and the outputs expected for them, so that a `-m gpu` test can replay the horizon through the product's
seams (PoolBuilder.compute_candidate_* + greedy_assignment + NPO) on CUDA and demand bit-exact pools and selections.

    python tests/golden/make_trace.py cfg1       # needs /root/reference; ~2 min
    python tests/golden/make_trace.py cfg2day    # needs /root/reference; hours (whole synthetic day, 2,878 epochs)

cfg1     BASELINE.json configs[0]: SBA + greedy_assignment, 100 vehicles cap 4, 30 + 60 + 39 min on the 4,091-node
         table.  The UNMODIFIED reference runs the whole horizon; at EVERY epoch the pool its own seam returned, the
         delays of its own scorer, the selection of its own greedy_assignment and the NPO matches are recorded.
cfg2day  BASELINE.json configs[1] as a whole synthetic DAY: OSP + greedy, 2,000 vehicles cap 4, 394,695 requests with a
         diurnal profile (README.md:9), 00:00 -> 23:59, 30 s epochs.  The reference's simulator (vehicle movement,
         routes, request life cycle, NPO) advances the day; the pool build is served by the C oracle (bit-identical to
         the reference's seam, tests/test_oracle_golden.py; the Python seam needs ~30 s per epoch at this size) and the
         greedy assigner by its set-based restatement (oracle.greedy_assign).  Every ~96th epoch plus the evening peak
         is recorded: seam inputs + fingerprints of the expected pool, the greedy selection and the NPO matches.
"""
from __future__ import annotations

import collections
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

from amod_b200 import marshal, synth  # noqa: E402
import ref_harness as rh  # noqa: E402

BIG_CUTOFF = 1e9


def digest_arrays(*arrs) -> str:
    h = hashlib.blake2b(digest_size=16)
    for a, dt in arrs:
        a = np.ascontiguousarray(a, dtype=np.dtype(dt))
        h.update(np.int64(a.size).tobytes())
        h.update(a.tobytes())
    return h.hexdigest()


def npo_inputs(ref, reqs, vehs):
    VS, OS = ref.types.VehicleStatus, ref.types.OrderStatus
    idle = [v for v in vehs if v.status == VS.IDLE]
    pend = [r for r in reqs if r.status == OS.PENDING]
    return idle, pend


def trace_cfg1():
    import make_golden as mg
    cap, fleet = 4, 100
    graph = synth.make_road_graph(seed=0, with_paths=True)
    t0 = 18 * 3600 + 30 * 60
    on, dn, tm = synth.make_request_stream(graph, 1700, t0, t0 + 135 * 60, seed=1, mean_trip_sec=600.0)
    ref = rh.setup_reference("/tmp/amod_ref_trace_cfg1", graph, on, dn, tm, fleet_size=fleet, capacity=cap, dispatcher="SBA",
                             warmup_min=30, main_min=60, winddown_min=39)
    out = {}
    state = {"n": 0, "cur": None}
    orig_sba = ref.dispatch_sba.compute_candidate_veh_req_pairs
    orig_assign = ref.ilp_assign.greedy_assignment
    orig_npo = ref.platform.reposition_idle_vehicles_to_nearest_pending_orders

    def sba_wrap(new_rids, reqs, vehs, T):
        recs = orig_sba(new_rids, reqs, vehs, T)
        ep, _objs = marshal.marshal_epoch(reqs, vehs, new_rids, T, cap, marshal.MODE_SBA)
        index_of = {int(rid): i for i, rid in enumerate(ep.req_id)}
        pool = marshal.records_to_pool(recs, vehs, index_of, marshal.MODE_SBA)
        nfeas, kinds = [], []
        for (veh, trip, _sche, _cost, _s) in recs:
            if trip:
                r = trip[0]
                fs = ref.scheduling.compute_schedule([veh.nid, veh.t_to_nid, veh.load], [veh.sche], [r.id, r.onid, r.dnid, r.Clp, r.Cld], T)[2]
                nfeas.append(len(fs)); kinds.append(marshal.KIND_SBA_PAIR)
            else:
                nfeas.append(1); kinds.append(marshal.KIND_SBA_EMPTY)
        pool.ent_nfeas = np.asarray(nfeas, dtype=np.int32)
        pool.ent_kind = np.asarray(kinds, dtype=np.int32)
        pool.ent_delay = np.asarray([ref.scheduling.compute_sche_delay(rec[0], rec[2], reqs, T) for rec in recs], dtype=np.float64)
        i = state["n"]
        out.update(ep.to_npz_dict(f"e{i}_ep_"))
        out.update(pool.to_npz_dict(f"e{i}_pool_"))
        state["cur"] = (i, {id(v): k for k, v in enumerate(vehs)})
        return recs

    def assign_wrap(pairs):
        sel = orig_assign(pairs)               # sorts `pairs` in place (ilp_assign.py:109), returns positions in it
        i, veh_index = state["cur"]
        out[f"e{i}_sel"] = np.asarray(sel, dtype=np.int32)
        out[f"e{i}_sel_veh"] = np.asarray([veh_index[id(pairs[s][0])] for s in sel], dtype=np.int32)
        out[f"e{i}_sel_nreq"] = np.asarray([len(pairs[s][1]) for s in sel], dtype=np.int32)
        out[f"e{i}_sel_rid"] = np.asarray([r.id for s in sel for r in pairs[s][1]], dtype=np.int32)
        return sel

    def npo_wrap(reqs, vehs, T, n_new, vf):
        idle, pend = npo_inputs(ref, reqs, vehs)
        orig_npo(reqs, vehs, T, n_new, vf)
        i = state["n"]
        VS = ref.types.VehicleStatus
        res = [(k, v.sche[0][2], v.sche[0][3]) for k, v in enumerate(idle) if v.status == VS.REBALANCING]
        out[f"e{i}_npo_idle_nid"] = np.asarray([v.nid for v in idle], dtype=np.int32)
        out[f"e{i}_npo_pend_onid"] = np.asarray([r.onid for r in pend], dtype=np.int32)
        out[f"e{i}_npo_m_idle"] = np.asarray([r[0] for r in res], dtype=np.int32)
        out[f"e{i}_npo_m_node"] = np.asarray([r[1] for r in res], dtype=np.int32)
        out[f"e{i}_npo_m_ddl"] = np.asarray([r[2] for r in res], dtype=np.float64)
        out[f"e{i}_T"] = np.asarray(T, dtype=np.int64)

    ref.dispatch_sba.compute_candidate_veh_req_pairs = sba_wrap
    ref.dispatch_sba.ilp_assignment = lambda pairs, rids, reqs, vehs, ensure=True: assign_wrap(pairs)
    ref.platform.reposition_idle_vehicles_to_nearest_pending_orders = npo_wrap
    pf = rh.make_platform(ref)
    t = time.time()
    for e in range(0, pf.system_shutdown_time_sec, 30):
        pf.run_cycle(e)
        state["n"] += 1
    pf.create_report(show=False)
    out["n_epochs"] = np.asarray(state["n"])
    out["spec"] = np.array(json.dumps(dict(kind="grid", n_nodes=graph.n_nodes, seed=0)))
    out["result"] = np.asarray(pf.main_sim_result, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "trace_cfg1.npz"), **out)
    print(f"cfg1: {state['n']} epochs in {time.time() - t:.0f}s, result {pf.main_sim_result}", flush=True)


def trace_cfg2day(sample_every: int = 96):
    import oracle
    cap, fleet = 4, 2000
    n_req = int(os.environ.get("DAY_REQUESTS", "394695"))
    graph = synth.make_road_graph(seed=0, with_paths=True)
    on, dn, _tm = synth.make_request_stream(graph, n_req, 0, 86400, seed=1, mean_trip_sec=600.0)
    tm = synth.diurnal_times(n_req, seed=1)
    on = np.append(on, on[-1]); dn = np.append(dn, dn[-1]); tm = np.append(tm, 10 ** 9)       # the reader looks one request ahead (platform.py:261)
    main_min = int(os.environ.get("DAY_MAIN_MIN", "1370"))
    ref = rh.setup_reference("/tmp/amod_ref_trace_cfg2day", graph, on, dn, tm, fleet_size=fleet, capacity=cap, dispatcher="OSP",
                             sim_start="2016-05-25 00:00:00", warmup_min=30, main_min=main_min, winddown_min=39)
    tt = graph.tt
    OS = ref.types.OrderStatus
    nthreads = oracle.max_threads()
    n_epochs = (30 + main_min + 39) * 2
    # evening peak (19:00) and morning peak (09:00) epochs are always recorded
    forced = {int(h * 120) for h in (9.0, 18.5, 19.0, 19.5, 20.0) if h * 120 < n_epochs}
    out = {}
    state = {"epoch": 0, "rec": None, "names": []}

    def want(e):
        return e % sample_every == 0 or e in forced

    def dispatch(mode, search_rids, reqs, vehs, T):
        ep, req_objs = marshal.marshal_epoch(reqs, vehs, search_rids, T, cap, mode)
        pool = (oracle.sba_build if mode == marshal.MODE_SBA else oracle.osp_build)(ep, tt, n_threads=nthreads)
        assert pool.error == 0, "trip-table level overflow"
        order, sel = oracle.greedy_assign(pool, ep.n_veh, ep.n_req)
        e = state["epoch"]
        if want(e):
            key = f"e{e}"
            out.update(ep.to_npz_dict(f"{key}_ep_"))
            out[f"{key}_expect"] = np.array(json.dumps({
                "mode": mode, "T": int(T), "pool_digest": marshal.pool_digest(pool), "n_entries": int(pool.n_entries),
                "n_evals": int(pool.stats["n_evals"]), "n_lookups": int(pool.stats["n_lookups"]),
                "order_digest": digest_arrays((order, "<i4")), "n_sel": int(sel.size)}))
            out[f"{key}_sel"] = sel.astype(np.int32)
            state["names"].append(e)
            print(f"  recorded epoch {e} (T={T}, mode {mode}): V={ep.n_veh} R={ep.n_search} entries={pool.n_entries} "
                  f"evals={pool.stats['n_evals']:,}", flush=True)
        if mode != marshal.MODE_SBA:
            for rid in search_rids:                              # dispatch_osp.py:77-78
                reqs[rid].status = OS.PENDING
        picked = pool.take(order[sel].astype(np.int64))
        recs = marshal.pool_to_records(picked, ep, vehs, req_objs)
        ref.scheduling.upd_schedule_for_vehicles_in_selected_vt_pairs(recs, list(range(len(recs))))

    def osp(new_rids, reqs, vehs, T, value_func, is_reoptimization=True):
        considered = [r.id for r in reqs if r.status in (OS.PICKING, OS.PENDING)]   # dispatch_osp.py:33-37
        dispatch(marshal.MODE_OSP, considered, reqs, vehs, T)

    def sba(new_rids, reqs, vehs, T, value_func):
        dispatch(marshal.MODE_SBA, new_rids, reqs, vehs, T)

    orig_npo = ref.platform.reposition_idle_vehicles_to_nearest_pending_orders

    def npo_wrap(reqs, vehs, T, n_new, vf):
        e = state["epoch"]
        if want(e):
            idle, pend = npo_inputs(ref, reqs, vehs)
            idle_nid = np.asarray([v.nid for v in idle], dtype=np.int32)
            pend_onid = np.asarray([r.onid for r in pend], dtype=np.int32)
            out[f"e{e}_npo_idle_nid"] = idle_nid
            out[f"e{e}_npo_pend_onid"] = pend_onid
            oi, op, od = oracle.npo_match(tt, idle_nid, pend_onid) if idle_nid.size and pend_onid.size else (np.zeros(0, np.int32),) * 2 + (np.zeros(0),)
            out[f"e{e}_npo_digest"] = np.array(digest_arrays((oi, "<i4"), (op, "<i4"), (od, "<f8")))
        # the day itself advances with the reference's own NPO on small inputs and with its array restatement on large
        # ones (the reference sorts idle x pending candidates in Python: minutes per epoch at 2,000 x 2,000)
        idle, pend = npo_inputs(ref, reqs, vehs)
        if len(idle) * len(pend) <= 20000:
            return orig_npo(reqs, vehs, T, n_new, vf)
        oi, op, od = oracle.npo_match(tt, np.asarray([v.nid for v in idle], dtype=np.int32), np.asarray([r.onid for r in pend], dtype=np.int32))
        for vi, ri, dt in zip(oi.tolist(), op.tolist(), od):
            idle[vi].build_route([(-1, 0, pend[ri].onid, T + np.float64(dt) + 240)])      # rebalancing_npo.py:37,62

    ref.platform.assign_orders_through_osp = osp
    ref.platform.assign_orders_through_sba = sba
    ref.platform.reposition_idle_vehicles_to_nearest_pending_orders = npo_wrap
    pf = rh.make_platform(ref)
    t_all = time.time()
    for e0 in range(0, pf.system_shutdown_time_sec, 30):
        state["epoch"] = e0 // 30 + 1
        t = time.time()
        pf.run_cycle(e0)
        if state["epoch"] % 20 == 0:
            st = [v.status.value for v in pf.vehs]
            print(f"epoch {state['epoch']}/{n_epochs}: {time.time() - t:.1f}s (total {time.time() - t_all:.0f}s), idle/working/reb = "
                  f"{st.count(1)}/{st.count(2)}/{st.count(3)}, reqs {len(pf.reqs)}", flush=True)
        if state["epoch"] % 480 == 0:          # checkpoint what was recorded so far
            out["epochs"] = np.asarray(state["names"], dtype=np.int32)
            np.savez_compressed(os.path.join(HERE, "trace_cfg2day.npz"), **out)
    pf.create_report(show=False)
    out["epochs"] = np.asarray(state["names"], dtype=np.int32)
    out["result"] = np.asarray(pf.main_sim_result, dtype=np.float64)
    out["spec"] = np.array(json.dumps(dict(kind="grid", n_nodes=graph.n_nodes, seed=0, requests=n_req, main_min=main_min)))
    np.savez_compressed(os.path.join(HERE, "trace_cfg2day.npz"), **out)
    print(f"cfg2day: {len(state['names'])} epochs recorded of {n_epochs}, {time.time() - t_all:.0f}s, result {pf.main_sim_result}", flush=True)


if __name__ == "__main__":
    assert rh.reference_available(), "run this where /root/reference exists"
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    if which == "cfg1":
        trace_cfg1()
    else:
        trace_cfg2day()
