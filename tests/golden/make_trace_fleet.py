"""Record the FLEET STATE of whole simulated horizons from the unmodified reference simulator (SURVEY.md §8f rank 3:
vehicle.py:73-302 move_to_time / build_route, platform.py:213-233 state advance + walk-away scan, scheduling.py:110-119
apply step), so that the device-resident fleet (amod_fleet_*, amod_b200/fleet.py) and its CPU restatement
(oracle/fleet_oracle.c) can be required to reproduce every vehicle and request field bit for bit, epoch by epoch.

    python tests/golden/make_trace_fleet.py sba     # cfg1: SBA + greedy + NPO, 100 vehicles cap 4, 258 epochs (~2 min)
    python tests/golden/make_trace_fleet.py osp     # OSP + greedy + NPO, 40 vehicles cap 4, 110 epochs (warm-up / wind-down in SBA)
    python tests/golden/make_trace_fleet.py ties    # SBA, 30 vehicles cap 3 on a 144-node grid whose edges take exactly 10 / 15 / 30 s:
                                                    # legs and steps end exactly on epoch boundaries (step.t == dT, pct == 1, leg.t == dT)

Everything recorded comes from the reference's own code (greedy_assignment in place of the Gurobi ILP as README.md:29
sanctions; the OSP wall-clock cut-off disabled as everywhere in this repository).  Per epoch:
  A  right after Platform.upd_vehs_and_reqs_stat_to_time: every vehicle field, the `done` events move_to_time returned,
     every request's status;
  B  the requests that arrived (the request table is stored once);
  C  every Veh.build_route call in call order (vehicle, the schedule passed, dispatch or rebalancing phase);
  D  at the end of the epoch: vehicle totals, schedules and the legs of every route; request statuses.
"""
from __future__ import annotations

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

from amod_b200 import synth  # noqa: E402
import ref_harness as rh  # noqa: E402

BIG_CUTOFF = 1e9
STATS = ("Ds", "Ts", "Ds_empty", "Ts_empty", "Dr", "Tr", "Lt", "Ld")


class Ragged:
    """Append-only CSR over (epoch) or (epoch, vehicle) rows."""

    def __init__(self, fields):
        self.off = [0]
        self.cols = {f: [] for f in fields}

    def row(self, items):
        for it in items:
            for f, v in zip(self.cols, it):
                self.cols[f].append(v)
        self.off.append(self.off[-1] + len(items))

    def dump(self, out, prefix, dtypes):
        out[prefix + "off"] = np.asarray(self.off, dtype=np.int64)
        for f, dt in zip(self.cols, dtypes):
            out[prefix + f] = np.asarray(self.cols[f], dtype=dt)


def record(which: str):
    cap = 4
    gspec = dict(kind="grid", n_nodes=synth.N_NODES_MANHATTAN, seed=0)
    trip = 600.0
    t0 = 18 * 3600 + 30 * 60
    if which == "sba":
        fleet, disp, warm, main, wind, n_req = 100, "SBA", 30, 60, 39, 1700
    elif which == "osp":
        fleet, disp, warm, main, wind, n_req = 40, "OSP", 10, 35, 10, 520
    else:       # "ties": every edge takes 10, 15 or 30 s, so legs and steps end exactly on the 30 s epoch boundaries
        fleet, disp, warm, main, wind, n_req = 30, "SBA", 5, 25, 5, 300
        gspec = dict(kind="grid", n_nodes=144, seed=7, edge_times=[10.0, 15.0, 30.0])
        trip, cap = 150.0, 3
    graph = synth.make_road_graph(gspec["n_nodes"], seed=gspec["seed"], with_paths=True, edge_times=gspec.get("edge_times"))
    horizon_min = warm + main + wind
    on, dn, tm = synth.make_request_stream(graph, n_req, t0, t0 + (horizon_min + 6) * 60, seed=1, mean_trip_sec=trip)
    ref = rh.setup_reference(f"/tmp/amod_ref_trace_fleet_{which}", graph, on, dn, tm, fleet_size=fleet, capacity=cap, dispatcher=disp,
                             warmup_min=warm, main_min=main, winddown_min=wind)
    Veh = ref.platform.Veh
    V = fleet
    cur = {"phase": 0, "events": [], "calls": []}

    orig_move = Veh.move_to_time

    def move_wrap(self, T, upd):
        done = orig_move(self, T, upd)
        for (rid, pod, t) in done:
            cur["events"].append((self.id, rid, pod, float(t)))
        return done

    orig_build = Veh.build_route

    def build_wrap(self, sche):
        cur["calls"].append((self.id, cur["phase"], [tuple(s) for s in sche]))
        return orig_build(self, sche)

    Veh.move_to_time = move_wrap
    Veh.build_route = build_wrap

    orig_osp_pairs = ref.dispatch_osp.compute_candidate_veh_trip_pairs
    ref.dispatch_osp.compute_candidate_veh_trip_pairs = \
        lambda new, cons, reqs, vehs, T, cutoff, reopt, fast: orig_osp_pairs(new, cons, reqs, vehs, T, BIG_CUTOFF, reopt, fast)
    orig_npo = ref.platform.reposition_idle_vehicles_to_nearest_pending_orders

    def npo_wrap(*a):
        cur["phase"] = 1
        try:
            return orig_npo(*a)
        finally:
            cur["phase"] = 0

    ref.platform.reposition_idle_vehicles_to_nearest_pending_orders = npo_wrap

    pf = rh.make_platform(ref)
    E = pf.system_shutdown_time_sec // 30
    z = lambda dt: np.zeros((E, V), dtype=dt)
    A = dict(nid=z(np.int32), lng=z(np.float64), lat=z(np.float64), t_to_nid=z(np.float64), d_to_nid=z(np.float64), tnid=z(np.int32),
             load=z(np.int32), status=z(np.int32), state_time=z(np.float64), t=z(np.float64), d=z(np.float64), has_step=z(np.uint8),
             step_t=z(np.float64), step_d=z(np.float64), step_u=z(np.int32), step_v=z(np.int32), n_legs=z(np.int32),
             leg0_t=z(np.float64), leg0_d=z(np.float64), leg0_steps=z(np.int32))
    for s in STATS:
        A[s] = z(np.float64)
    D = dict(t=z(np.float64), d=z(np.float64), tnid=z(np.int32), status=z(np.int32), n_legs=z(np.int32))
    a_sche = Ragged(("rid", "pod", "tnid", "ddl"))          # rows = (epoch, vehicle)
    d_sche = Ragged(("rid", "pod", "tnid", "ddl"))
    d_legs = Ragged(("rid", "pod", "tnid", "ddl", "t", "d", "n_steps"))
    a_events = Ragged(("veh", "rid", "pod", "time"))        # rows = epoch
    a_rstat = Ragged(("status",))
    d_rstat = Ragged(("status",))
    calls = Ragged(("veh", "phase", "n_stops"))
    call_stops = Ragged(("rid", "pod", "tnid", "ddl"))       # rows = call
    n_reqs_after = np.zeros(E, dtype=np.int64)
    Ts = np.zeros(E, dtype=np.int64)

    orig_upd = pf.upd_vehs_and_reqs_stat_to_time
    state = {"e": 0}

    def upd_wrap():
        orig_upd()
        e = state["e"]
        for v in pf.vehs:
            i = v.id
            A["nid"][e, i] = v.nid; A["lng"][e, i] = v.lng; A["lat"][e, i] = v.lat
            A["t_to_nid"][e, i] = v.t_to_nid; A["d_to_nid"][e, i] = v.d_to_nid; A["tnid"][e, i] = v.tnid
            A["load"][e, i] = v.load; A["status"][e, i] = v.status.value; A["state_time"][e, i] = v.state_time
            A["t"][e, i] = v.t; A["d"][e, i] = v.d
            for s in STATS:
                A[s][e, i] = getattr(v, s)
            st = v.step_to_nid
            if st is not None:
                A["has_step"][e, i] = 1
                A["step_t"][e, i] = st.t; A["step_d"][e, i] = st.d
                A["step_u"][e, i] = st.nid_pair[0]; A["step_v"][e, i] = st.nid_pair[1]
            A["n_legs"][e, i] = len(v.route)
            if v.route:
                A["leg0_t"][e, i] = v.route[0].t; A["leg0_d"][e, i] = v.route[0].d; A["leg0_steps"][e, i] = len(v.route[0].steps)
            a_sche.row([tuple(s) for s in v.sche])
        a_events.row(cur["events"])
        cur["events"] = []
        a_rstat.row([(r.status.value,) for r in pf.reqs])

    pf.upd_vehs_and_reqs_stat_to_time = upd_wrap
    t = time.time()
    for e in range(E):
        state["e"] = e
        cur["calls"] = []
        pf.run_cycle(e * 30)
        Ts[e] = pf.system_time_sec
        n_reqs_after[e] = len(pf.reqs)
        calls.row([(c[0], c[1], len(c[2])) for c in cur["calls"]])
        for c in cur["calls"]:
            call_stops.row(c[2])
        for v in pf.vehs:
            i = v.id
            D["t"][e, i] = v.t; D["d"][e, i] = v.d; D["tnid"][e, i] = v.tnid; D["status"][e, i] = v.status.value; D["n_legs"][e, i] = len(v.route)
            d_sche.row([tuple(s) for s in v.sche])
            d_legs.row([(l.rid, l.pod, l.tnid, l.ddl, l.t, l.d, len(l.steps)) for l in v.route])
        d_rstat.row([(r.status.value,) for r in pf.reqs])
    pf.create_report(show=False)

    out = {}
    for k, a in A.items():
        out["A_" + k] = a
    for k, a in D.items():
        out["D_" + k] = a
    stop_dt = (np.int32, np.int8, np.int32, np.float64)
    a_sche.dump(out, "A_sche_", stop_dt)
    d_sche.dump(out, "D_sche_", stop_dt)
    d_legs.dump(out, "D_leg_", stop_dt + (np.float64, np.float64, np.int32))
    a_events.dump(out, "A_ev_", (np.int32, np.int32, np.int8, np.float64))
    a_rstat.dump(out, "A_rstat_", (np.uint8,))
    d_rstat.dump(out, "D_rstat_", (np.uint8,))
    calls.dump(out, "C_", (np.int32, np.uint8, np.int32))
    call_stops.dump(out, "C_stop_", stop_dt)
    reqs = pf.reqs
    out["req_onid"] = np.asarray([r.onid for r in reqs], dtype=np.int32)
    out["req_dnid"] = np.asarray([r.dnid for r in reqs], dtype=np.int32)
    out["req_tr"] = np.asarray([r.Tr for r in reqs], dtype=np.float64)
    out["req_ts"] = np.asarray([r.Ts for r in reqs], dtype=np.float64)
    out["req_clp"] = np.asarray([r.Clp for r in reqs], dtype=np.float64)
    out["req_cld"] = np.asarray([r.Cld for r in reqs], dtype=np.float64)
    out["req_tp"] = np.asarray([r.Tp for r in reqs], dtype=np.float64)
    out["req_td"] = np.asarray([r.Td for r in reqs], dtype=np.float64)
    out["req_status"] = np.asarray([r.status.value for r in reqs], dtype=np.uint8)
    out["n_reqs_after"] = n_reqs_after
    out["T"] = Ts
    out["start_nid"] = np.asarray([int(graph.stations[int(i * len(graph.stations) / V)]) for i in range(V)], dtype=np.int32)
    out["result"] = np.asarray(pf.main_sim_result, dtype=np.float64)
    out["spec"] = np.array(json.dumps(dict(gspec, cap=cap, fleet=V, dispatcher=disp, warmup_min=warm, main_min=main, winddown_min=wind)))
    path = os.path.join(HERE, f"trace_fleet_{which}.npz")
    np.savez_compressed(path, **out)
    print(f"{which}: {E} epochs in {time.time() - t:.0f}s, {len(reqs)} requests, result {pf.main_sim_result}, "
          f"{os.path.getsize(path) / 1e6:.2f} MB", flush=True)


if __name__ == "__main__":
    assert rh.reference_available(), "run this where /root/reference exists"
    record(sys.argv[1] if len(sys.argv) > 1 else "sba")
