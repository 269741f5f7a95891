"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py            # needs /root/reference (absent on the GPU box)

Every vector is: seam inputs (an `Epoch`, amod_b200/marshal.py) + the table recipe + the pool the
reference's own code returned for them, with the wall-clock cut-off disabled (the cut-off is an
argument of the seam, dispatch_osp.py:94-97; with the shipped 0.1 s the output depends on host speed).
Families:
  sim_*    live calls captured inside the reference simulator (Platform.run_cycle), i.e. real Veh/Req
           objects -> marshal_epoch -> Epoch, and the records the reference returned for them
  rand_*   random self-consistent epochs (synth.random_epoch) on small non-metric tables with many exact
           ties / arbitrary float64 values, turned into reference objects and pushed through the seams
  unit     scheduling.test_constraints_get_cost on random insertions, all six violation codes
  npo_*    the NPO candidate/sort/greedy loop (rebalancing_npo.py:26-59)
  levels   a case where the reference raises IndexError (trip of size cap+4, dispatch_osp.py:226)
"""
from __future__ import annotations

import collections
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from amod_b200 import marshal, synth  # noqa: E402
import ref_harness as rh  # noqa: E402

BIG_CUTOFF = 1e9


def table_from_spec(spec: dict) -> np.ndarray:
    """The single place a golden file's travel-time table is rebuilt from its recipe."""
    if spec["kind"] == "grid":
        return synth.make_road_graph(spec["n_nodes"], seed=spec["seed"]).tt
    return synth.random_table(spec["n_nodes"], spec["seed"], spec["kind"],
                              np.float64 if spec["dtype"] == "f64" else np.float32)


def epoch_to_objects(ref, ep: marshal.Epoch):
    """Epoch arrays -> reference Veh/Req objects carrying exactly the fields the seams read."""
    from src.simulator.request import Req
    from src.simulator.vehicle import Veh
    Leg = ref.types.Leg
    VS = ref.types.VehicleStatus
    OS = ref.types.OrderStatus
    num = lambda x: int(x) if float(x).is_integer() else float(x)
    reqs = [None] * (int(ep.req_id.max()) + 1 if ep.n_req else 0)
    objs = []
    for i in range(ep.n_req):
        r = Req.__new__(Req)
        r.id = int(ep.req_id[i]); r.status = OS.PENDING
        r.Tr = num(ep.req_tr[i]); r.onid = int(ep.req_onid[i]); r.dnid = int(ep.req_dnid[i])
        r.Ts = np.float64(ep.req_ts[i]); r.Clp = num(ep.req_clp[i]); r.Cld = np.float64(ep.req_cld[i])
        reqs[r.id] = r
        objs.append(r)
    vehs = []
    for v in range(ep.n_veh):
        veh = Veh.__new__(Veh)
        veh.id = v; veh.status = VS(int(ep.veh_status[v])); veh.nid = int(ep.veh_nid[v])
        veh.t_to_nid = float(ep.veh_t_to_nid[v]); veh.load = int(ep.veh_load[v])
        veh.sche, veh.onboard_rids, veh.picking_rids = [], [], []
        veh.route = collections.deque()
        for k in range(ep.veh_sche_off[v], ep.veh_sche_off[v + 1]):
            c = int(ep.sche_code[k])
            if c < 0:
                stop = (-1, 0, int(ep.veh_reb_node[v]), float(ep.veh_reb_ddl[v]))
            else:
                r = objs[c >> 1]
                stop = (r.id, -1, r.dnid, r.Cld) if c & 1 else (r.id, 1, r.onid, r.Clp)
                if ep.sche_onboard[k]:
                    veh.onboard_rids.append(r.id)
                if not (c & 1):
                    veh.picking_rids.append(r.id)
                    r.status = OS.PICKING
            veh.sche.append(stop)
            veh.route.append(Leg(stop[0], stop[1], stop[2], stop[3]))
        vehs.append(veh)
    return reqs, objs, vehs


def run_reference_seam(ref, ep: marshal.Epoch, reqs, objs, vehs):
    """Call the reference seam for `ep` and convert what it returned (plus delay / nfeas obtained
    from the reference's own functions) into Pool arrays.  Returns None if it raised IndexError."""
    import src.simulator.config as cfg
    cfg.config_change_vc(ep.cap)
    cfg.config_change_fs(max(ep.n_veh, 1))
    T = ep.T
    index_of = {int(rid): i for i, rid in enumerate(ep.req_id)}
    search_rids = [int(ep.req_id[i]) for i in ep.search]
    if ep.mode == marshal.MODE_SBA:
        recs = ref.dispatch_sba.compute_candidate_veh_req_pairs(search_rids, reqs, vehs, T)
        nfeas = []
        for (veh, trip, sche, cost, _s) in recs:
            if trip:
                r = trip[0]
                _b, _c, fs = ref.scheduling.compute_schedule([veh.nid, veh.t_to_nid, veh.load], [veh.sche],
                                                             [r.id, r.onid, r.dnid, r.Clp, r.Cld], T)
                nfeas.append(len(fs))
            else:
                nfeas.append(1)
        kinds = [marshal.KIND_SBA_PAIR if rec[1] else marshal.KIND_SBA_EMPTY for rec in recs]
    else:
        reopt = ep.mode == marshal.MODE_OSP
        try:
            recs = ref.dispatch_osp.compute_candidate_veh_trip_pairs(search_rids, search_rids, reqs, vehs, T,
                                                                     BIG_CUTOFF, reopt, False)
        except IndexError:
            return None
        nfeas, kinds = [], []
        table = ref.dispatch_osp.FEASIBLE_TRIP_TABLE
        for veh in vehs:
            nfeas.append(len(ref.dispatch_osp.compute_basic_sches_of_one_veh(veh, T, reopt)))
            kinds.append(marshal.KIND_BASIC)
            for level in table[veh.id]:
                for (_trip, _best, _cost, fs) in level:
                    nfeas.append(len(fs)); kinds.append(marshal.KIND_TRIP)
            if reopt:
                nfeas.append(1); kinds.append(marshal.KIND_CURRENT)
    pool = marshal.records_to_pool(recs, vehs, index_of, ep.mode)
    assert len(nfeas) == pool.n_entries
    pool.ent_nfeas = np.asarray(nfeas, dtype=np.int32)
    pool.ent_kind = np.asarray(kinds, dtype=np.int32)
    pool.ent_delay = np.asarray([ref.scheduling.compute_sche_delay(rec[0], rec[2], reqs, T) for rec in recs],
                                dtype=np.float64)
    return pool


def save_cases(name: str, spec: dict, cases: list):
    d = {"spec": np.array(json.dumps(spec)), "n_cases": np.array(len(cases))}
    for i, (ep, pool) in enumerate(cases):
        d.update(ep.to_npz_dict(f"c{i}_ep_"))
        if pool is None:
            d[f"c{i}_error"] = np.array(1)
        else:
            d.update(pool.to_npz_dict(f"c{i}_pool_"))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"  wrote {path}: {len(cases)} cases, {os.path.getsize(path) / 1024:.0f} KiB", flush=True)


def gen_rand(ref):
    """rand_*: stress epochs through the reference seams."""
    rf = ref.route_functions
    plans = [
        ("rand_ties_osp_cap4", dict(kind="ties", n_nodes=24, seed=11, dtype="f64"), marshal.MODE_OSP, 4,
         [dict(n_veh=12, n_pending=7, seed=s, near=(s % 2 == 0), wait_sec=150) for s in range(6)]),
        ("rand_real_osp_cap4", dict(kind="real", n_nodes=40, seed=12, dtype="f64"), marshal.MODE_OSP, 4,
         [dict(n_veh=14, n_pending=8, seed=100 + s, near=(s % 2 == 0), wait_sec=200) for s in range(6)]),
        ("rand_real32_osp_cap2", dict(kind="real", n_nodes=40, seed=13, dtype="f32"), marshal.MODE_OSP, 2,
         [dict(n_veh=14, n_pending=10, seed=200 + s, near=True, wait_sec=220) for s in range(4)]),
        ("rand_ties_osp_cap8", dict(kind="ties", n_nodes=30, seed=14, dtype="f64"), marshal.MODE_OSP, 8,
         [dict(n_veh=8, n_pending=5, seed=300 + s, near=True, max_picking=1, wait_sec=120) for s in range(3)]),
        ("rand_real_nr_cap4", dict(kind="real", n_nodes=40, seed=15, dtype="f64"), marshal.MODE_OSP_NR, 4,
         [dict(n_veh=14, n_pending=14, seed=400 + s, near=True, wait_sec=200) for s in range(4)]),
        ("rand_ties_sba_cap4", dict(kind="ties", n_nodes=24, seed=16, dtype="f64"), marshal.MODE_SBA, 4,
         [dict(n_veh=30, n_pending=30, seed=500 + s, near=(s % 2 == 0), frac_new=0.8, wait_sec=200) for s in range(4)]),
        ("rand_real_sba_cap6", dict(kind="real", n_nodes=60, seed=17, dtype="f32"), marshal.MODE_SBA, 6,
         [dict(n_veh=40, n_pending=40, seed=600 + s, near=True, frac_new=0.8, max_picking=3, wait_sec=300) for s in range(4)]),
    ]
    for name, spec, mode, cap, kws in plans:
        tt = table_from_spec(spec)
        rf.mean_travel_time_table = tt.astype(np.float64)
        cases = []
        t0 = time.time()
        import oracle   # sizing guard only: the Python reference does ~5e4 evaluations/s
        for kw in kws:
            kw = dict(kw)
            for attempt in range(50):   # walk seeds until the case is neither trivial nor explosive
                ep = synth.random_epoch(tt, cap=cap, mode=mode, **kw)
                est = (oracle.sba_build if mode == marshal.MODE_SBA else oracle.osp_build)(ep, tt, n_threads=4)
                if 1_000 <= est.stats["n_evals"] <= 250_000 and est.error == 0:
                    break
                kw["seed"] += 1000
            else:
                raise RuntimeError((name, kw))
            reqs, objs, vehs = epoch_to_objects(ref, ep)
            pool = run_reference_seam(ref, ep, reqs, objs, vehs)
            cases.append((ep, pool))
        print(f"  {name}: {[c[1].n_entries if c[1] else None for c in cases]} entries ({time.time() - t0:.1f}s)", flush=True)
        save_cases(name, spec, cases)


def gen_levels(ref):
    """A vehicle and cap+4 co-located zero-length requests: the reference runs out of trip-table
    levels and raises IndexError (dispatch_osp.py:8,226)."""
    spec = dict(kind="ties", n_nodes=6, seed=21, dtype="f64")
    tt = table_from_spec(spec)
    ref.route_functions.mean_travel_time_table = tt.astype(np.float64)
    cases = []
    for nreq in (4, 5):          # cap=1: 4 requests fit in the table (sizes 1..4), 5 do not
        T, cap = 1000, 1
        ep = marshal.Epoch(T=T, cap=cap, mode=marshal.MODE_OSP, veh_nid=[3], veh_t_to_nid=[0.0], veh_load=[0],
                           veh_status=[1], veh_sche_off=[0, 0], sche_code=[], sche_onboard=[], veh_reb_node=[0],
                           veh_reb_ddl=[0.0], req_id=np.arange(nreq), req_onid=[3] * nreq, req_dnid=[3] * nreq,
                           req_clp=[T + 400.0] * nreq, req_cld=[T + 800.0] * nreq, req_tr=[T - 20.0] * nreq,
                           req_ts=[0.0] * nreq, search=np.arange(nreq))
        reqs, objs, vehs = epoch_to_objects(ref, ep)
        pool = run_reference_seam(ref, ep, reqs, objs, vehs)
        print(f"  levels nreq={nreq}: {'IndexError' if pool is None else pool.n_entries}", flush=True)
        cases.append((ep, pool))
    assert cases[0][1] is not None and cases[1][1] is None
    save_cases("levels", spec, cases)


def gen_unit(ref):
    """unit: scheduling.test_constraints_get_cost on random insertions."""
    spec = dict(kind="real", n_nodes=40, seed=31, dtype="f64")
    tt = table_from_spec(spec)
    ref.route_functions.mean_travel_time_table = tt.astype(np.float64)
    import src.simulator.config as cfg
    rng = np.random.default_rng(5)
    rows, eps = [], []
    seen = collections.Counter()
    for case in range(4):
        cap = (2, 3, 4, 4)[case]
        ep = synth.random_epoch(tt, n_veh=20, n_pending=20, cap=cap, mode=marshal.MODE_OSP, seed=700 + case, near=True)
        cfg.config_change_vc(cap)
        reqs, objs, vehs = epoch_to_objects(ref, ep)
        eps.append(ep)
        for _ in range(400):
            v = int(rng.integers(0, ep.n_veh)); q = int(ep.search[rng.integers(0, ep.n_search)])
            veh, r = vehs[v], objs[q]
            if r.id in [s[0] for s in veh.sche]:
                continue
            l = len(veh.sche)
            i = int(rng.integers(0, l + 1)); j = int(rng.integers(i + 1, l + 2))
            sche = list(veh.sche)
            sche.insert(i, (r.id, 1, r.onid, r.Clp)); sche.insert(j, (r.id, -1, r.dnid, r.Cld))
            flag, cost, viol = ref.scheduling.test_constraints_get_cost([veh.nid, veh.t_to_nid, veh.load], sche,
                                                                        r.id, i, j, ep.T)
            if seen[viol] > 60 and viol in (-1, 3):
                continue
            seen[viol] += 1
            rows.append((case, v, q, i, j, viol, float(cost)))
    print(f"  unit: {len(rows)} rows, viol histogram {dict(seen)}", flush=True)
    assert set(seen) == {-1, 0, 1, 2, 3, 4}, seen
    d = {"spec": np.array(json.dumps(spec)), "n_cases": np.array(len(eps)),
         "rows": np.array([r[:6] for r in rows], dtype=np.int64), "costs": np.array([r[6] for r in rows])}
    for i, ep in enumerate(eps):
        d.update(ep.to_npz_dict(f"c{i}_ep_"))
    np.savez_compressed(os.path.join(HERE, "unit.npz"), **d)


def gen_sim(ref_kwargs, name, spec, n_requests, every, fleet, cap, dispatcher, horizon=(5, 12, 4), req_seed=1,
            mean_trip=420.0):
    """sim_*: capture live seam calls inside the reference simulator (fresh process per run)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    p = ctx.Process(target=_sim_worker, args=(name, spec, n_requests, every, fleet, cap, dispatcher, horizon,
                                              req_seed, mean_trip))
    p.start(); p.join()
    assert p.exitcode == 0, name


def _sim_worker(name, spec, n_requests, every, fleet, cap, dispatcher, horizon, req_seed, mean_trip):
    graph = synth.make_road_graph(spec["n_nodes"], seed=spec["seed"], with_paths=True)
    t0 = 18 * 3600 + 30 * 60
    total_min = sum(horizon)
    on, dn, tm = synth.make_request_stream(graph, n_requests, t0, t0 + (total_min + 10) * 60, seed=req_seed,
                                           mean_trip_sec=mean_trip)
    ref = rh.setup_reference(f"/tmp/amod_ref_{name}", graph, on, dn, tm, fleet_size=fleet, capacity=cap,
                             dispatcher=dispatcher, warmup_min=horizon[0], main_min=horizon[1],
                             winddown_min=horizon[2])
    cases, npo_cases = [], []
    counter = collections.Counter()
    orig_osp = ref.dispatch_osp.compute_candidate_veh_trip_pairs
    orig_sba = ref.dispatch_sba.compute_candidate_veh_req_pairs
    orig_npo = ref.platform.reposition_idle_vehicles_to_nearest_pending_orders

    def capture(mode, search_rids, reqs, vehs, T, recs, reopt=True):
        ep, req_objs = marshal.marshal_epoch(reqs, vehs, search_rids, T, cap, mode)
        index_of = {int(rid): i for i, rid in enumerate(ep.req_id)}
        pool = marshal.records_to_pool(recs, vehs, index_of, mode)
        nfeas, kinds = [], []
        if mode == marshal.MODE_SBA:
            for (veh, trip, sche, cost, _s) in recs:
                if trip:
                    r = trip[0]
                    fs = ref.scheduling.compute_schedule([veh.nid, veh.t_to_nid, veh.load], [veh.sche],
                                                         [r.id, r.onid, r.dnid, r.Clp, r.Cld], T)[2]
                    nfeas.append(len(fs)); kinds.append(marshal.KIND_SBA_PAIR)
                else:
                    nfeas.append(1); kinds.append(marshal.KIND_SBA_EMPTY)
        else:
            table = ref.dispatch_osp.FEASIBLE_TRIP_TABLE
            for veh in vehs:
                nfeas.append(len(ref.dispatch_osp.compute_basic_sches_of_one_veh(veh, T, reopt)))
                kinds.append(marshal.KIND_BASIC)
                for level in table[veh.id]:
                    for (_t, _b, _c, fs) in level:
                        nfeas.append(len(fs)); kinds.append(marshal.KIND_TRIP)
                if reopt:
                    nfeas.append(1); kinds.append(marshal.KIND_CURRENT)
        pool.ent_nfeas = np.asarray(nfeas, dtype=np.int32)
        pool.ent_kind = np.asarray(kinds, dtype=np.int32)
        pool.ent_delay = np.asarray([ref.scheduling.compute_sche_delay(rec[0], rec[2], reqs, T) for rec in recs])
        cases.append((ep, pool))

    def osp_wrap(new_rids, considered, reqs, vehs, T, cutoff, reopt, fast):
        recs = orig_osp(new_rids, considered, reqs, vehs, T, BIG_CUTOFF, reopt, fast)
        counter["osp"] += 1
        if counter["osp"] % every == 0:
            capture(marshal.MODE_OSP if reopt else marshal.MODE_OSP_NR, considered, reqs, vehs, T, recs, reopt)
        return recs

    def sba_wrap(new_rids, reqs, vehs, T):
        recs = orig_sba(new_rids, reqs, vehs, T)
        counter["sba"] += 1
        if counter["sba"] % every == 0:
            capture(marshal.MODE_SBA, new_rids, reqs, vehs, T, recs)
        return recs

    def npo_wrap(reqs, vehs, T, n_new, vf):
        VS, OS = ref.types.VehicleStatus, ref.types.OrderStatus
        idle = [v for v in vehs if v.status == VS.IDLE]
        pend = [r for r in reqs if r.status == OS.PENDING]
        orig_npo(reqs, vehs, T, n_new, vf)
        counter["npo"] += 1
        if idle and pend and counter["npo"] % 2 == 0:
            res = [(i, v.sche[0][2], v.sche[0][3]) for i, v in enumerate(idle) if v.status == VS.REBALANCING]
            npo_cases.append((np.array([v.nid for v in idle], dtype=np.int32),
                              np.array([r.onid for r in pend], dtype=np.int32), T,
                              np.array([r[0] for r in res], dtype=np.int32), np.array([r[1] for r in res], dtype=np.int32),
                              np.array([r[2] for r in res], dtype=np.float64)))

    ref.dispatch_osp.compute_candidate_veh_trip_pairs = osp_wrap
    ref.dispatch_sba.compute_candidate_veh_req_pairs = sba_wrap
    ref.platform.reposition_idle_vehicles_to_nearest_pending_orders = npo_wrap
    pf = rh.make_platform(ref)
    t = time.time()
    for e in range(0, pf.system_shutdown_time_sec, 30):
        pf.run_cycle(e)
    pf.create_report(show=False)
    print(f"  {name}: sim {time.time() - t:.1f}s, result {pf.main_sim_result}, calls {dict(counter)}, "
          f"captured {len(cases)} (+{len(npo_cases)} npo), entries {[c[1].n_entries for c in cases]}", flush=True)
    save_cases(name, spec, cases)
    d = {"spec": np.array(json.dumps(spec)), "n_cases": np.array(len(npo_cases))}
    for i, c in enumerate(npo_cases):
        for k, a in zip(("idle_nid", "pend_onid", "T", "m_idle", "m_node", "m_ddl"), c):
            d[f"c{i}_{k}"] = np.asarray(a)
    np.savez_compressed(os.path.join(HERE, name.replace("sim_", "npo_") + ".npz"), **d)


def main():
    assert rh.reference_available(), "run this where /root/reference exists"
    which = sys.argv[1:] or ["rand", "levels", "unit", "sim"]
    if "sim" in which:
        gen_sim({}, "sim_osp_cap4", dict(kind="grid", n_nodes=400, seed=3), 1500, 6, 30, 4, "OSP")
        gen_sim({}, "sim_osp_cap2", dict(kind="grid", n_nodes=300, seed=4), 1300, 7, 24, 2, "OSP", req_seed=2)
        gen_sim({}, "sim_ospnr_cap4", dict(kind="grid", n_nodes=400, seed=5), 1500, 7, 30, 4, "OSP-NR", req_seed=3)
        gen_sim({}, "sim_sba_cap6", dict(kind="grid", n_nodes=500, seed=6), 1800, 6, 40, 6, "SBA", req_seed=4)
    if {"rand", "levels", "unit"} & set(which):
        # one reference import serves all object-level families (tables/capacity are patched per family)
        g = synth.make_road_graph(16, seed=1, with_paths=True)
        ref = rh.setup_reference("/tmp/amod_ref_rand", g, np.array([1, 2], dtype=np.int32), np.array([2, 3], dtype=np.int32),
                                 np.array([70000, 80000]), fleet_size=4, capacity=4)
        if "rand" in which:
            gen_rand(ref)
        if "levels" in which:
            gen_levels(ref)
        if "unit" in which:
            gen_unit(ref)


if __name__ == "__main__":
    main()
