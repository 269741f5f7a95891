"""Produce the epoch snapshots bench.py replays (BASELINE.json configs[1]: OSP pool build,
2,000 vehicles, capacity 4, Manhattan-scale synthetic demand, 30 s epochs).

Runs HERE (needs /root/reference): the reference's own simulator (Platform.run_cycle, vehicle
movement, route building, NPO) advances the fleet on the synthetic 4,091-node table; only the
dispatcher's pool build is served by the C oracle (bit-identical to the reference, see
tests/test_oracle_golden.py, but ~100x faster than the Python loop) and the greedy assigner is
replaced by an equivalent set-based scan (the reference's `in list` scans are O(P*V)).
At chosen epochs the seam inputs are stored as Epoch arrays:  bench_data/cfg2_*.npz.
The table itself is rebuilt from its seed on the GPU box (synth.make_road_graph(seed=0)).

    python bench_data/make_snapshots.py [fleet] [cap] [req_per_hour] [tag]
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from amod_b200 import marshal, synth  # noqa: E402
import oracle  # noqa: E402
import ref_harness as rh  # noqa: E402


def main():
    fleet = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    cap = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    per_hour = int(sys.argv[3]) if len(sys.argv) > 3 else 27000
    tag = sys.argv[4] if len(sys.argv) > 4 else "cfg2"
    warm, main_min, wind = 30, int(os.environ.get("MAIN_MIN", "40")), 1
    snap_epochs = [int(x) for x in os.environ.get("SNAP", "70,80,90,100,110,120,130,138").split(",")]
    graph = synth.make_road_graph(seed=0, with_paths=True)
    t0 = 18 * 3600 + 30 * 60
    horizon = (warm + main_min + wind + 5) * 60
    nreq = int(per_hour * horizon / 3600)
    on, dn, tm = synth.make_request_stream(graph, nreq, t0, t0 + horizon, seed=1, mean_trip_sec=600.0)
    ref = rh.setup_reference(f"/tmp/amod_ref_{tag}", graph, on, dn, tm, fleet_size=fleet, capacity=cap, dispatcher="OSP",
                             warmup_min=warm, main_min=main_min, winddown_min=wind)
    tt = graph.tt
    OS = ref.types.OrderStatus
    nthreads = oracle.max_threads()
    log = []

    def greedy_on_arrays(pool):
        score = np.diff(pool.ent_trip_off)                       # pair[4] = len(trip) (scheduling.py:160-161)
        order = np.argsort(-score, kind="stable")                # list.sort(key=-score) is stable (ilp_assign.py:109)
        used_v, used_r, sel = set(), set(), []
        ev, toff, treq = pool.ent_veh, pool.ent_trip_off, pool.trip_req
        for e in order.tolist():
            v = int(ev[e])
            if v in used_v:
                continue
            rs = treq[toff[e]:toff[e + 1]].tolist()
            if any(r in used_r for r in rs):
                continue
            used_v.add(v); used_r.update(rs); sel.append(e)
        return sel

    def dispatch(mode, search_rids, reqs, vehs, T, epoch_idx):
        ep, req_objs = marshal.marshal_epoch(reqs, vehs, search_rids, T, cap, mode)
        t = time.time()
        pool = (oracle.sba_build if mode == marshal.MODE_SBA else oracle.osp_build)(ep, tt, n_threads=nthreads)
        dt = time.time() - t
        assert pool.error == 0, "trip-table level overflow"
        if mode != marshal.MODE_SBA:
            sizes = np.bincount(np.diff(pool.ent_trip_off), minlength=8)
            print(f"  epoch {epoch_idx} T={T}: V={ep.n_veh} R={ep.n_search} entries={pool.n_entries} evals={pool.stats['n_evals']:,} "
                  f"sches={pool.stats['n_sches_stored']:,} sizes={sizes.tolist()} oracle {dt:.1f}s", flush=True)
            if epoch_idx in snap_epochs:
                np.savez_compressed(os.path.join(HERE, f"{tag}_epoch{epoch_idx:04d}.npz"), **ep.to_npz_dict("ep_"),
                                    n_evals=pool.stats["n_evals"], n_lookups=pool.stats["n_lookups"], n_entries=pool.n_entries)
        sel = greedy_on_arrays(pool)
        if mode != marshal.MODE_SBA:
            for rid in search_rids:                              # dispatch_osp.py:77-78
                reqs[rid].status = OS.PENDING
        picked = pool.take(np.asarray(sel, dtype=np.int64))
        recs = marshal.pool_to_records(picked, ep, vehs, req_objs)
        ref.scheduling.upd_schedule_for_vehicles_in_selected_vt_pairs(recs, list(range(len(recs))))

    state = {"epoch": 0}

    sba_advance = os.environ.get("ADVANCE") == "sba"
    probe = int(os.environ.get("PROBE_STEP", "0"))

    def osp(new_rids, reqs, vehs, T, value_func, is_reoptimization=True):
        considered = [r.id for r in reqs if r.status in (OS.PICKING, OS.PENDING)]   # dispatch_osp.py:33-37
        if sba_advance:
            # dense configs (cfg4: demand x4): an exhaustive OSP build per epoch takes the CPU oracle many minutes, so the
            # day advances with the reference's SBA dispatcher (as it does during warm-up, platform.py:160-161) and the
            # OSP seam inputs of the chosen epochs are recorded without being applied
            if state["epoch"] in snap_epochs:
                ep, _objs = marshal.marshal_epoch(reqs, vehs, considered, T, cap, marshal.MODE_OSP)
                if probe:
                    t = time.time()
                    p = oracle.osp_build(ep, tt, n_threads=nthreads, veh_lo=0, veh_step=probe)
                    print(f"  probe epoch {state['epoch']}: every {probe}th vehicle: {p.stats['n_evals']:,} evals, {p.n_entries} entries, "
                          f"{p.stats['n_sches_stored']:,} sches in {time.time() - t:.1f}s  (V={ep.n_veh} R={ep.n_search})", flush=True)
                np.savez_compressed(os.path.join(HERE, f"{tag}_epoch{state['epoch']:04d}.npz"), **ep.to_npz_dict("ep_"),
                                    n_evals=0, n_lookups=0, n_entries=0)
            return dispatch(marshal.MODE_SBA, new_rids, reqs, vehs, T, state["epoch"])
        dispatch(marshal.MODE_OSP, considered, reqs, vehs, T, state["epoch"])

    def sba(new_rids, reqs, vehs, T, value_func):
        if sba_advance and state["epoch"] in snap_epochs:       # warm-up epochs of a dense run: the fleet in mixed states (cfg5's source)
            considered = [r.id for r in reqs if r.status in (OS.PICKING, OS.PENDING)]
            ep, _objs = marshal.marshal_epoch(reqs, vehs, considered, T, cap, marshal.MODE_OSP)
            np.savez_compressed(os.path.join(HERE, f"{tag}_epoch{state['epoch']:04d}.npz"), **ep.to_npz_dict("ep_"),
                                n_evals=0, n_lookups=0, n_entries=0)
        dispatch(marshal.MODE_SBA, new_rids, reqs, vehs, T, state["epoch"])

    ref.platform.assign_orders_through_osp = osp
    ref.platform.assign_orders_through_sba = sba
    pf = rh.make_platform(ref)
    t_all = time.time()
    for e in range(0, pf.system_shutdown_time_sec, 30):
        state["epoch"] = e // 30 + 1
        t = time.time()
        pf.run_cycle(e)
        if state["epoch"] % 10 == 0:
            st = [v.status.value for v in pf.vehs]
            print(f"epoch {state['epoch']}: {time.time() - t:.1f}s, idle/working/reb = {st.count(1)}/{st.count(2)}/{st.count(3)}, "
                  f"reqs {len(pf.reqs)}", flush=True)
        if state["epoch"] > max(snap_epochs):
            break
    print(f"done in {time.time() - t_all:.0f}s")


if __name__ == "__main__":
    main()
