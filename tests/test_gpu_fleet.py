"""SURVEY.md 8f-3 (second half): the fleet's state advance, request life cycle and apply step on the device (amod_fleet_*,
amod_b200/fleet.py) against the fleet state recorded from the UNMODIFIED reference simulator (tests/golden/trace_fleet_*.npz,
made by make_trace_fleet.py): every vehicle and request field after every epoch's advance and at every epoch's end, bit for bit.

  open loop    the recorded arrivals, status changes and build_route calls drive the device fleet (pins advance + apply);
  closed loop  ONLY the request stream goes in: pool build, greedy assignment, apply, NPO and advance all run on the device for
               the whole horizon, and the state must still equal the reference's at every epoch (the north-star target run with
               the epoch state resident in HBM);
  at scale     2,000 vehicles / OSP in closed loop, the CPU restatement (oracle/fleet_oracle.c) following with the schedules the
               device chose: advance and routes agree at cfg2 size.
"""
import numpy as np
import pytest

from amod_b200 import marshal, synth
from amod_b200.fleet import DeviceFleetSim
from amod_b200.pool import PoolBuilder
import fleet_util as fu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def graph():
    return synth.make_road_graph(seed=0, with_paths=True)


@pytest.fixture(scope="module")
def builder(graph):
    b = PoolBuilder(graph.tt)
    yield b
    b.close()


_builders = {}


def _builder_for(tr):
    g = tr.graph()
    if id(g) not in _builders:
        _builders[id(g)] = PoolBuilder(g.tt)
    return g, _builders[id(g)]


def _final(tr, x, n):
    fu._same("final Tp", x["req_tp"], tr.z["req_tp"], n)
    fu._same("final Td", x["req_td"], tr.z["req_td"], n)
    fu._same("final status", x["req_status"], tr.z["req_status"], n)


@pytest.mark.parametrize("which", ["sba", "osp", "ties"])
def test_device_fleet_replays_the_reference_horizon_open_loop(which):
    tr = fu.FleetTrace(which)
    graph, builder = _builder_for(tr)
    sim = DeviceFleetSim(builder, graph, tr.z["start_nid"], max_stops=24, max_route_steps=4096, request_capacity=4096)
    n = fu.replay_open_loop(tr, sim)
    assert n == tr.E
    _final(tr, sim.export(), n)


@pytest.mark.parametrize("which", ["sba", "osp", "ties"])
def test_device_fleet_runs_the_reference_horizon_closed_loop(which):
    """Only the arrivals go in; everything else happens on the device, for the whole horizon."""
    tr = fu.FleetTrace(which)
    graph, builder = _builder_for(tr)
    z, cap = tr.z, tr.spec["cap"]
    sim = DeviceFleetSim(builder, graph, z["start_nid"], max_stops=24, max_route_steps=4096, request_capacity=4096)
    main_mode = marshal.MODE_OSP if tr.spec["dispatcher"] == "OSP" else marshal.MODE_SBA
    for e in range(tr.E):
        T = int(tr.T[e])
        ev = sim.advance(T, tr.in_main(T))
        fu.check_after_advance(tr, e, sim.export(), ev)
        lo, hi = tr.new_requests(e)
        if hi > lo:
            sim.add_requests(lo, z["req_onid"][lo:hi], z["req_dnid"][lo:hi], z["req_tr"][lo:hi], z["req_ts"][lo:hi], z["req_clp"][lo:hi], z["req_cld"][lo:hi])
        sim.dispatch(main_mode if tr.in_main(T) else marshal.MODE_SBA, cap, T, hi - lo)       # platform.py:140-150
        sim.rebalance(T)                                                                       # platform.py:153-155
        fu.check_after_epoch(tr, e, sim.export())
    _final(tr, sim.export(), tr.E)


def test_device_fleet_at_cfg2_scale_against_the_cpu_restatement(graph, builder):
    import fleet_oracle
    V, cap, E = 2000, 4, 24
    rng = np.random.default_rng(5)
    start = graph.stations[(np.arange(V) * len(graph.stations) // V)].astype(np.int32)
    on, dn, tm = synth.make_request_stream(graph, 150 * (E + 2), 0, 30 * (E + 2), seed=3, mean_trip_sec=600.0)
    sim = DeviceFleetSim(builder, graph, start, max_stops=24, max_route_steps=4096, request_capacity=1 << 14)
    orc = fleet_oracle.FleetOracle(graph, start)
    n_req = 0
    cols = [c for c in fu.A_EXACT]
    picked = dropped = 0
    for e in range(E):
        T = 30 * (e + 1)
        ev = sim.advance(T, True)
        ev_o = orc.advance(T, True)
        x, y = sim.export(), orc.export()
        for c in cols + ["sche_len", "sche_rid", "sche_pod", "sche_tnid", "sche_ddl", "n_legs", "leg_t", "leg_d", "leg_n_steps", "req_status", "req_tp", "req_td"]:
            fu._same(c, x[c], y[c], e)
        for c in ("veh", "rid", "pod", "time"):
            fu._same("ev_" + c, ev[c], ev_o[c], e)
        picked += int((ev["pod"] == 1).sum()); dropped += int((ev["pod"] == -1).sum())
        m = np.flatnonzero((tm >= T - 30) & (tm < T))
        if m.size:
            ts = graph.tt[on[m] - 1, dn[m] - 1].astype(np.float64)
            tr_ = tm[m].astype(np.float64)
            args = (n_req, on[m], dn[m], tr_, ts, tr_ + 420, (tr_ + ts) + 840)             # request.py:31-34
            sim.add_requests(*args); orc.add_requests(*args)
            n_req += m.size
        before = sim.array("req_status").copy()
        sim.dispatch(marshal.MODE_OSP, cap, T, m.size)
        # the CPU restatement follows with the schedules the device chose (every vehicle is selected exactly once)
        x = sim.export()
        for s in (1, 2):
            ch = np.flatnonzero((x["req_status"] != before) & (x["req_status"] == s))
            if ch.size:
                orc.set_status(ch, s)
        off = np.zeros(V + 1, np.int32); np.cumsum(x["sche_len"], out=off[1:])
        orc.apply(np.arange(V), off, x["sche_rid"], x["sche_pod"], x["sche_tnid"], x["sche_ddl"])
        status_before = x["status"].copy()
        sim.rebalance(T)
        x = sim.export()
        reb = np.flatnonzero((x["status"] == 3) & (status_before == 1))
        if reb.size:
            o2 = np.zeros(V + 1, np.int64); np.cumsum(x["sche_len"], out=o2[1:])
            take = o2[reb]
            orc.apply(reb, np.arange(reb.size + 1, dtype=np.int32), x["sche_rid"][take], x["sche_pod"][take], x["sche_tnid"][take], x["sche_ddl"][take])
        y = orc.export()
        for c in ("t", "d", "tnid", "status", "n_legs", "leg_rid", "leg_pod", "leg_tnid", "leg_ddl", "leg_t", "leg_d", "leg_n_steps"):
            fu._same("end." + c, x[c], y[c], e)
    assert picked > 200 and dropped > 0          # the fleet really moved and served


def test_fleet_errors_are_reported_not_swallowed(graph, builder):
    """Error behaviour of the C-ABI: capacity, bad arguments and broken invariants come back as codes + messages."""
    from amod_b200 import capi
    from amod_b200.pool import AmodError
    start = graph.stations[:8].astype(np.int32)
    # a route that does not fit the per-vehicle step region -> AMOD_E_CAPACITY naming the size needed
    sim = DeviceFleetSim(builder, graph, start, max_stops=8, max_route_steps=8, request_capacity=64)
    far = int(np.argmax(graph.tt[start[0] - 1])) + 1
    sim.add_requests(0, [far], [int(start[0])], [0.0], [float(graph.tt[far - 1, start[0] - 1])], [420.0], [5000.0])
    with pytest.raises(AmodError) as ei:
        sim.apply([0], [0, 2], [0, 0], [1, -1], [far, int(start[0])], [420.0, 5000.0])
    assert ei.value.code == capi.E_CAPACITY and "steps" in str(ei.value)
    # requests must arrive in id order; unknown vehicles / requests are refused
    sim = DeviceFleetSim(builder, graph, start, max_stops=8, max_route_steps=2048, request_capacity=64)
    with pytest.raises(AmodError):
        sim.add_requests(3, [far], [far], [0.0], [0.0], [420.0], [840.0])
    with pytest.raises(AmodError):
        sim.apply([99], [0, 0], [], [], [], [])
    with pytest.raises(AmodError):
        sim.apply([0], [0, 1], [5], [1], [far], [420.0])          # request 5 was never added
    with pytest.raises(AmodError):
        sim.set_status([7], 2)
    # picking up a request that is not PICKING trips the reference's assertion (request.py:49) as an error, not silently
    sim = DeviceFleetSim(builder, graph, start, max_stops=8, max_route_steps=2048, request_capacity=64)
    near = int(start[0])
    sim.add_requests(0, [near], [far], [0.0], [float(graph.tt[near - 1, far - 1])], [420.0], [9000.0])
    sim.apply([0], [0, 2], [0, 0], [1, -1], [near, far], [420.0, 9000.0])     # status left PENDING on purpose
    with pytest.raises(AmodError) as ei:
        sim.advance(30, False)
    assert "invariant" in str(ei.value)


def test_idle_fleet_and_empty_epochs(graph, builder):
    """No requests at all: vehicles stay idle at their stations, time advances, every build is the V basic pairs."""
    start = graph.stations[:16].astype(np.int32)
    sim = DeviceFleetSim(builder, graph, start, max_stops=8, max_route_steps=1024, request_capacity=64)
    for e in range(3):
        T = 30 * (e + 1)
        ev = sim.advance(T, True)
        assert ev["veh"].size == 0
        n_sel = sim.dispatch(marshal.MODE_OSP if e else marshal.MODE_SBA, 4, T, 0)
        assert n_sel == 16 and sim.rebalance(T) == 0
        x = sim.export()
        assert (x["status"] == 1).all() and (x["state_time"] == T).all() and (x["nid"] == start).all() and x["sche_len"].sum() == 0
        assert x["Ts"].sum() == 0 and x["n_legs"].sum() == 0
