"""The north-star target run with the epoch state resident on the device: the WHOLE synthetic day of BASELINE configs[1]
(2,000 vehicles cap 4, 394,695 requests with a diurnal profile, 2,878 epochs of 30 s, SBA warm-up / wind-down, OSP + greedy +
NPO in the main horizon) simulated by the device fleet in closed loop — only the request stream goes in — and compared with
the day the reference's simulator produced (tests/golden/trace_cfg2day.npz, make_trace.py cfg2day) at its 33 sampled epochs:
every vehicle row at dispatch time, the pool fingerprint, the evaluation counters, the greedy order and selection, the NPO
inputs and matches; and at the end of the day the service counts of the reference's report (platform.py:344-376)."""
import json

import numpy as np
import pytest

from amod_b200 import marshal, state as astate, synth
from amod_b200.fleet import DeviceFleetSim, RQ_COMPLETE, RQ_ONBOARD, RQ_PENDING, ST_IDLE
from amod_b200.pool import PoolBuilder
import trace_util as tu

pytestmark = pytest.mark.gpu


def test_whole_synthetic_day_on_the_device():
    z = tu.load_trace("trace_cfg2day")
    if z is None or "spec" not in z.files:
        pytest.skip("tests/golden/trace_cfg2day.npz not recorded")
    spec = json.loads(str(z["spec"]))
    V, cap, warm, main, wind = 2000, 4, 30, int(spec["main_min"]), 39
    graph = synth.make_road_graph(spec["n_nodes"], seed=spec["seed"], with_paths=True)
    n_req = int(spec["requests"])
    on, dn, _tm = synth.make_request_stream(graph, n_req, 0, 86400, seed=1, mean_trip_sec=600.0)      # make_trace.py trace_cfg2day
    tm = synth.diurnal_times(n_req, seed=1)
    ts_all = graph.tt[on - 1, dn - 1].astype(np.float64)                                               # request.py:31
    tr_all = tm.astype(np.float64)
    start = np.asarray([int(graph.stations[int(i * len(graph.stations) / V)]) for i in range(V)], dtype=np.int32)    # platform.py:50-55
    b = PoolBuilder(graph.tt)
    sim = DeviceFleetSim(b, graph, start, max_stops=24, max_route_steps=4096, request_capacity=n_req + 8)
    E = (warm + main + wind) * 2
    sampled = set(z["epochs"].tolist())
    in_main = lambda T: warm * 60 < T <= (warm + main) * 60
    nxt = 0
    checked = 0
    for e in range(1, E + 1):
        T = 30 * e
        sim.advance(T, in_main(T))
        hi = int(np.searchsorted(tm, T, side="left"))            # platform.py:261: request_time_sec < system_time_sec
        if hi > nxt:
            sl = slice(nxt, hi)
            sim.add_requests(nxt, on[sl], dn[sl], tr_all[sl], ts_all[sl], tr_all[sl] + 420, (tr_all[sl] + ts_all[sl]) + 840)
        n_new, nxt = hi - nxt, hi
        mode = marshal.MODE_OSP if in_main(T) else marshal.MODE_SBA
        rec = e in sampled
        if rec:
            ep = marshal.Epoch.from_npz_dict(z, f"e{e}_ep_")
            exp = json.loads(str(z[f"e{e}_expect"]))
            assert ep.mode == mode and int(ep.T) == T, e
            rows = astate.DeviceFleet.rows_of(ep, sim.max_stops)
            for k, name in (("nid", "nid"), ("t", "t_to_nid"), ("load", "load"), ("status", "status"), ("len", "sche_len"),
                            ("code", "sche_code"), ("onboard", "sche_onboard")):
                assert np.array_equal(np.asarray(rows[k]), sim.array(name)), f"epoch {e}: vehicle rows differ in {name}"
            has_reb = (rows["code"] == -1).any(axis=1) & (rows["len"] > 0)
            assert np.array_equal(rows["reb_node"][has_reb], sim.array("reb_node")[has_reb]), e
            assert np.array_equal(rows["reb_ddl"][has_reb], sim.array("reb_ddl")[has_reb]), e
        st = sim.build(mode, cap, T, n_new)
        if rec:
            assert np.array_equal(ep.req_id[ep.search], np.arange(hi - n_new, hi) if mode == marshal.MODE_SBA else ep.req_id[ep.search]), e
            assert sim.last_n_search == ep.n_search, (e, sim.last_n_search, ep.n_search)
            assert (st["n_entries"], st["n_evals"], st["n_lookups"]) == (exp["n_entries"], exp["n_evals"], exp["n_lookups"]), (e, st, exp)
            pool = b.fetch()                                     # names requests by id: back to the epoch's request indices
            rid = ep.req_id.astype(np.int64)
            idx_of = lambda ids: np.searchsorted(rid, ids)
            c = pool.sche_code.astype(np.int64)
            kw = {f: getattr(pool, f) for f in marshal.Pool.FIELDS}
            kw["sche_code"] = np.where(c < 0, -1, 2 * idx_of(np.maximum(c, 0) >> 1) + (c & 1)).astype(np.int32)
            kw["trip_req"] = idx_of(pool.trip_req.astype(np.int64)).astype(np.int32)
            assert marshal.pool_digest(marshal.Pool(stats=pool.stats, **kw)) == exp["pool_digest"], f"epoch {e}: pool differs"
        order, sel = b.greedy_assign()
        if rec:
            assert tu.digest_arrays((order, "<i4")) == exp["order_digest"], e
            assert np.array_equal(sel, z[f"e{e}_sel"]), e
        sim.apply_selected(None, mark_picking=True, reset_considered=(mode == marshal.MODE_OSP))
        if rec and f"e{e}_npo_idle_nid" in z.files:
            status, nid, rs = sim.array("status"), sim.array("nid"), sim.array("req_status")
            idle, pend = np.flatnonzero(status == ST_IDLE), np.flatnonzero(rs == RQ_PENDING)
            assert np.array_equal(nid[idle], z[f"e{e}_npo_idle_nid"]) and np.array_equal(sim._onid[pend], z[f"e{e}_npo_pend_onid"]), e
            if idle.size and pend.size:
                oi, op, od = b.npo_match(nid[idle], sim._onid[pend])
            else:
                oi, op, od = np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0)
            assert tu.digest_arrays((oi, "<i4"), (op, "<i4"), (od, "<f8")) == str(z[f"e{e}_npo_digest"]), e
        sim.rebalance(T)
        checked += rec
    assert checked == len(sampled)
    # the reference's report (platform.py:344-376): requests of the main horizon, served = complete + onboard
    rs = sim.array("req_status")
    m = (tr_all[:sim.n_req] > warm * 60) & (tr_all[:sim.n_req] <= (warm + main) * 60)
    served = int(((rs == RQ_COMPLETE) | (rs == RQ_ONBOARD))[m].sum())
    total = int(m.sum())
    want = z["result"]
    assert (served, total) == (int(want[0]), int(want[1])), (served, total, want)
    assert round(100.0 * served / total, 2) == float(want[2])
    b.close()
