"""Drop-in check of the seam plumbing against the LIVE reference simulator (runs only where the
reference tree exists, i.e. in the build container; skipped on the GPU box).

The reference simulator is run twice on the same synthetic day: (a) unmodified, (b) with the three
seams rebound exactly as amod_b200.install does, except that the array-level pool build / NPO match
is served by the oracle instead of the CUDA library (there is no GPU here).  Everything else on the
replacement path is the product's own code: marshal_epoch, pool_to_records, the scorer using device
delays, the NPO host wrapper.  Per-request outcomes must be identical."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_harness as rh  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")


def _run(dispatcher: str, patched, q, assigner: str = "greedy", scale: str = "small"):
    from amod_b200 import install as inst, marshal, pool as pool_mod, synth
    import oracle
    t0 = 18 * 3600 + 30 * 60
    if scale == "cfg1":     # BASELINE.json configs[0]: 100 vehicles cap 4, 1 h main horizon, the full-size 4,091-node table
        graph = synth.make_road_graph(seed=0, with_paths=True)
        on, dn, tm = synth.make_request_stream(graph, 1700, t0, t0 + 135 * 60, seed=1, mean_trip_sec=600.0)
        horizon = dict(fleet_size=100, warmup_min=30, main_min=60, winddown_min=39)
    else:
        graph = synth.make_road_graph(400, seed=3, with_paths=True)
        on, dn, tm = synth.make_request_stream(graph, 1500, t0, t0 + 40 * 60, seed=1, mean_trip_sec=420.0)
        horizon = dict(fleet_size=30, warmup_min=5, main_min=12, winddown_min=4)
    ref = rh.setup_reference(f"/tmp/amod_ref_dropin_{dispatcher}_{patched}_{assigner}_{scale}", graph, on, dn, tm, capacity=4,
                             dispatcher=dispatcher, **horizon)
    patched_tag = str(patched)
    # both runs must search exhaustively: disable the wall-clock cut-off of the unmodified seam as well
    orig = ref.dispatch_osp.compute_candidate_veh_trip_pairs
    ref.dispatch_osp.compute_candidate_veh_trip_pairs = lambda n, c, r, v, T, cutoff, reopt, fast: orig(n, c, r, v, T, 1e9, reopt, fast)
    if patched == "routes":     # all legs of an epoch built in one batch, served to the reference's unchanged Veh.build_route
        from amod_b200 import routes as aroutes
        from test_routes_cpu import OracleRouteBuilder
        import src.simulator.vehicle as vehicle_mod
        geo = np.array([ref.route_functions.get_node_geo(n) for n in range(1, graph.n_nodes + 1)], dtype=np.float64)
        rb = OracleRouteBuilder(graph, geo)
        aroutes.install_routes(rb, vehicle_mod, [ref.dispatch_osp, ref.dispatch_sba])
        patched = False
    if patched:
        tt = graph.tt

        class OracleBackedBuilder(pool_mod.PoolBuilder):       # test double for the CUDA library only
            def __init__(self):
                self.last_stats = {}

            def build_pool(self, ep, reuse=False):
                return (oracle.sba_build if ep.mode == marshal.MODE_SBA else oracle.osp_build)(ep, tt)

            def npo_match(self, idle_nid, pend_onid):
                return oracle.npo_match(tt, np.asarray(idle_nid, dtype=np.int32), np.asarray(pend_onid, dtype=np.int32))

            def close(self):
                pass

            # doubles for the device calls behind greedy_assignment (patched == "lazy")
            def build_pool(self, ep, reuse=False):          # noqa: F811  (keeps the last pool for greedy_assign)
                self._last = ((oracle.sba_build if ep.mode == marshal.MODE_SBA else oracle.osp_build)(ep, tt), ep)
                return self._last[0]

            def greedy_assign(self):
                pool, ep = self._last
                return oracle.greedy_assign(pool, ep.n_veh, ep.n_req)

        b = OracleBackedBuilder()
        if patched == "lazy":       # the install(greedy_on_gpu=True) wiring: lazy records + greedy on the "device"
            b.lazy_records = True
            ga = lambda pairs, rids=None, reqs=None, vehs=None, ensure=True: b.greedy_assignment(pairs)
            ref.dispatch_osp.ilp_assignment = ga
            ref.dispatch_sba.ilp_assignment = ga
        ref.dispatch_osp.compute_candidate_veh_trip_pairs = b.compute_candidate_veh_trip_pairs
        ref.dispatch_sba.compute_candidate_veh_req_pairs = b.compute_candidate_veh_req_pairs
        ref.dispatch_osp.score_vt_pairs_with_num_of_orders_and_sche_cost = pool_mod.score_vt_pairs_with_num_of_orders_and_sche_cost
        ref.dispatch_sba.score_vt_pairs_with_num_of_orders_and_sche_cost = pool_mod.score_vt_pairs_with_num_of_orders_and_sche_cost
        ref.platform.reposition_idle_vehicles_to_nearest_pending_orders = b.reposition_idle_vehicles_to_nearest_pending_orders
    if assigner == "milp":      # exact assignment on HiGHS in place of the Gurobi model (amod_b200/assign_milp.py)
        from amod_b200.assign_milp import milp_assignment
        ref.dispatch_osp.ilp_assignment = milp_assignment
        ref.dispatch_sba.ilp_assignment = milp_assignment
    pf = rh.make_platform(ref)
    for e in range(0, pf.system_shutdown_time_sec, 30):
        pf.run_cycle(e)
    pf.create_report(show=False)
    out = [(r.id, r.status.value, float(r.Tp), float(r.Td)) for r in pf.reqs]
    veh = [(v.id, v.nid, float(v.t_to_nid), v.load, v.status.value, float(v.Ds)) for v in pf.vehs]
    q.put((pf.main_sim_result, out, veh))


@pytest.mark.parametrize("dispatcher", ["OSP", "SBA"])
def test_simulation_outcome_identical_with_seams_replaced(dispatcher):
    ctx = mp.get_context("spawn")          # the reference freezes its config at import: one process per run
    res = []
    for patched in (False, True, "lazy"):
        q = ctx.Queue()
        p = ctx.Process(target=_run, args=(dispatcher, patched, q))
        p.start()
        res.append(q.get(timeout=600))
        p.join()
    (r0, reqs0, veh0) = res[0]
    assert len(reqs0) > 300 and r0[0] > 10
    for (r1, reqs1, veh1) in res[1:]:
        assert r0 == r1
        assert reqs0 == reqs1
        assert veh0 == veh1


def test_simulation_outcome_identical_with_batched_routes():
    """Apply step (scheduling.py:110-119 -> vehicle.py:231-302) fed by batched route construction: the simulated run is
    identical to the unmodified reference's, including every vehicle's travelled distance."""
    ctx = mp.get_context("spawn")
    res = []
    for patched in (False, "routes"):
        q = ctx.Queue()
        p = ctx.Process(target=_run, args=("OSP", patched, q))
        p.start()
        res.append(q.get(timeout=600))
        p.join()
    assert res[0] == res[1]


def test_simulation_outcome_identical_with_exact_assignment():
    """OSP with the exact assignment model (HiGHS stand-in for the absent Gurobi): the replaced seams hand the
    solver a bit-identical pool, so the simulated day is identical too."""
    ctx = mp.get_context("spawn")
    res = []
    for patched in (False, True, "lazy"):        # "lazy": the install(exact_assignment=True) wiring, model built from the pool arrays
        q = ctx.Queue()
        p = ctx.Process(target=_run, args=("OSP", patched, q, "milp"))
        p.start()
        res.append(q.get(timeout=900))
        p.join()
    (r0, reqs0, veh0) = res[0]
    assert len(reqs0) > 300 and r0[0] > 10
    for (r1, reqs1, veh1) in res[1:]:
        assert r0 == r1 and reqs0 == reqs1 and veh0 == veh1


def test_cfg1_simulated_hour_identical_with_seams_replaced():
    """BASELINE.json configs[0] at full size — SBA + greedy assignment, 100 vehicles of capacity 4, 30 + 60 + 39
    simulated minutes (258 epochs) on the 4,091-node synthetic table: per-request outcomes, per-vehicle end states
    and the report are identical for the unmodified reference, the replaced seams with eager records and the
    replaced seams with lazy records + the greedy assigner on the (stand-in) device."""
    ctx = mp.get_context("spawn")
    res = []
    for patched in (False, True, "lazy"):
        q = ctx.Queue()
        p = ctx.Process(target=_run, args=("SBA", patched, q, "greedy", "cfg1"))
        p.start()
        res.append(q.get(timeout=900))
        p.join()
    (r0, reqs0, veh0) = res[0]
    assert len(reqs0) > 1200 and len(veh0) == 100 and r0[0] > 300
    for (r1, reqs1, veh1) in res[1:]:
        assert r0 == r1 and reqs0 == reqs1 and veh0 == veh1
