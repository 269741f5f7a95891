"""What the REFERENCE's own Python spends per epoch on the rows SURVEY.md §8f ranks 2-3 name — the state advance
(Platform.upd_vehs_and_reqs_stat_to_time, platform.py:213-233), the apply step (upd_schedule_for_vehicles_in_selected_vt_pairs
-> Veh.build_route, scheduling.py:110-119, vehicle.py:231-302) and NPO — at cfg2 scale (2,000 vehicles cap 4, ~137
arrivals/epoch), for comparison with scripts/prof_fleet.py (the same epoch resident on the device).  Needs /root/reference
(runs in the build container, not on the GPU box); the pool build itself is served by the C oracle as in make_trace.py.

    python scripts/prof_ref_fleet_cpu.py --epochs 150
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from amod_b200 import marshal, synth  # noqa: E402
import oracle  # noqa: E402
import ref_harness as rh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--veh", type=int, default=2000)
    ap.add_argument("--epochs", type=int, default=150)
    ap.add_argument("--tail", type=int, default=60)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    cap, V, E = 4, a.veh, a.epochs
    graph = synth.make_road_graph(seed=0, with_paths=True)
    t0 = 18 * 3600 + 30 * 60
    on, dn, tm = synth.make_request_stream(graph, int(137 * (E + 4)), t0, t0 + 30 * (E + 4), seed=3, mean_trip_sec=600.0)
    on = np.append(on, on[-1]); dn = np.append(dn, dn[-1]); tm = np.append(tm, 10 ** 9)
    ref = rh.setup_reference("/tmp/amod_ref_prof_fleet", graph, on, dn, tm, fleet_size=V, capacity=cap, dispatcher="OSP",
                             warmup_min=0, main_min=(E * 30) // 60 + 1, winddown_min=1)
    OS = ref.types.OrderStatus
    tt = graph.tt
    nthreads = oracle.max_threads()
    acc = {"advance": [], "apply": [], "npo": [], "pool(oracle port)": [], "epoch": []}
    cur = {}

    def osp(new_rids, reqs, vehs, T, value_func, is_reoptimization=True):
        t = time.perf_counter()
        considered = [r.id for r in reqs if r.status in (OS.PICKING, OS.PENDING)]
        ep, req_objs = marshal.marshal_epoch(reqs, vehs, considered, T, cap, marshal.MODE_OSP)
        pool = oracle.osp_build(ep, tt, n_threads=nthreads)
        order, sel = oracle.greedy_assign(pool, ep.n_veh, ep.n_req)
        for rid in considered:
            reqs[rid].status = OS.PENDING
        picked = pool.take(order[sel].astype(np.int64))
        recs = marshal.pool_to_records(picked, ep, vehs, req_objs)
        t1 = time.perf_counter()
        ref.scheduling.upd_schedule_for_vehicles_in_selected_vt_pairs(recs, list(range(len(recs))))
        cur["apply"] = time.perf_counter() - t1
        cur["pool(oracle port)"] = t1 - t

    orig_npo = ref.platform.reposition_idle_vehicles_to_nearest_pending_orders

    def npo(*args):
        t = time.perf_counter()
        orig_npo(*args)
        cur["npo"] = time.perf_counter() - t

    ref.platform.assign_orders_through_osp = osp
    ref.platform.assign_orders_through_sba = lambda new, reqs, vehs, T, vf: osp(new, reqs, vehs, T, vf)
    ref.platform.reposition_idle_vehicles_to_nearest_pending_orders = npo
    pf = rh.make_platform(ref)
    orig_upd = pf.upd_vehs_and_reqs_stat_to_time

    def upd():
        t = time.perf_counter()
        orig_upd()
        cur["advance"] = time.perf_counter() - t

    pf.upd_vehs_and_reqs_stat_to_time = upd
    for e in range(E):
        cur.clear()
        t = time.perf_counter()
        pf.run_cycle(e * 30)
        cur["epoch"] = time.perf_counter() - t
        for k in acc:
            acc[k].append(1e3 * cur.get(k, 0.0))
        if e % 20 == 0:
            print(f"epoch {e}: " + ", ".join(f"{k} {acc[k][-1]:.1f} ms" for k in acc), flush=True)
    k = min(a.tail, E)
    out = {"workload": f"reference Python simulator, {V} veh cap {cap}, OSP + greedy + NPO, 137 arrivals/epoch, {E} epochs (means over the last {k}); "
                       f"pool build served by the C oracle port on {nthreads} threads", "cores": 1,
           "ms": {s: float(np.mean(v[-k:])) for s, v in acc.items()}}
    line = json.dumps(out)
    print(line)
    if a.out:
        open(a.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
