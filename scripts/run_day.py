"""The whole synthetic day of BASELINE configs[1] simulated with the fleet resident on the device (amod_b200/fleet.py):
2,000 vehicles cap 4, 394,695 requests (diurnal profile), 2,878 epochs, SBA warm-up / wind-down, OSP + greedy + NPO between.
Prints one JSON line: wall time of the day, per-hour mean epoch time, the heaviest epochs, the final report counts.

    python scripts/run_day.py --out gpurun_out/day.json          # on a B200; --epochs N stops early (profiling)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from amod_b200 import marshal, synth  # noqa: E402
from amod_b200.fleet import DeviceFleetSim, RQ_COMPLETE, RQ_ONBOARD  # noqa: E402
from amod_b200.pool import PoolBuilder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=0)
    ap.add_argument("--first-epoch-timed", type=int, default=1)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    V, cap, warm, main_min, wind, n_req = 2000, 4, 30, 1370, 39, 394695
    graph = synth.make_road_graph(seed=0, with_paths=True)
    on, dn, _ = synth.make_request_stream(graph, n_req, 0, 86400, seed=1, mean_trip_sec=600.0)
    tm = synth.diurnal_times(n_req, seed=1)
    ts_all = graph.tt[on - 1, dn - 1].astype(np.float64)
    tr_all = tm.astype(np.float64)
    start = np.asarray([int(graph.stations[int(i * len(graph.stations) / V)]) for i in range(V)], dtype=np.int32)
    b = PoolBuilder(graph.tt)
    sim = DeviceFleetSim(b, graph, start, max_stops=24, max_route_steps=4096, request_capacity=n_req + 8)
    E = (warm + main_min + wind) * 2
    if a.epochs:
        E = min(E, a.epochs)
    in_main = lambda T: warm * 60 < T <= (warm + main_min) * 60
    ms = np.zeros(E); evals = np.zeros(E, np.int64); entries = np.zeros(E, np.int64); nsearch = np.zeros(E, np.int64)
    nxt = 0
    t_day = time.perf_counter()
    for e in range(1, E + 1):
        T = 30 * e
        t0 = time.perf_counter()
        hi = int(np.searchsorted(tm, T, side="left"))
        arrivals = None
        if hi > nxt:
            sl = slice(nxt, hi)
            arrivals = (nxt, on[sl], dn[sl], tr_all[sl], ts_all[sl], tr_all[sl] + 420, (tr_all[sl] + ts_all[sl]) + 840)
        nxt = hi
        sim.epoch(T, in_main(T), marshal.MODE_OSP if in_main(T) else marshal.MODE_SBA, cap, arrivals)
        ms[e - 1] = 1e3 * (time.perf_counter() - t0)
        evals[e - 1] = sim.last_stats["n_evals"]; entries[e - 1] = sim.last_stats["n_entries"]; nsearch[e - 1] = sim.last_n_search
    wall = time.perf_counter() - t_day
    rs = sim.array("req_status")
    m = (tr_all[:sim.n_req] > warm * 60) & (tr_all[:sim.n_req] <= (warm + main_min) * 60)
    served, total = int(((rs == RQ_COMPLETE) | (rs == RQ_ONBOARD))[m].sum()), int(m.sum())
    top = np.argsort(-ms)[:5]
    per_hour = {}
    for h in range(24):
        sel = np.arange(E)[(np.arange(1, E + 1) * 30 > h * 3600) & (np.arange(1, E + 1) * 30 <= (h + 1) * 3600)]
        if sel.size:
            per_hour[str(h)] = {"ms_per_epoch": float(ms[sel].mean()), "requests_considered": float(nsearch[sel].mean()),
                                "pool_entries": float(entries[sel].mean()), "schedules": float(evals[sel].mean())}
    out = {"workload": f"whole synthetic day on the device: {V} veh cap {cap}, {n_req} requests, {E} epochs (closed loop, only arrivals uploaded)",
           "wall_s": wall, "ms_per_epoch_mean": float(ms.mean()), "ms_per_epoch_p99": float(np.percentile(ms, 99)), "ms_per_epoch_max": float(ms.max()),
           "schedules_evaluated": int(evals.sum()), "pool_entries": int(entries.sum()),
           "heaviest_epochs": [{"epoch": int(i + 1), "ms": float(ms[i]), "requests_considered": int(nsearch[i]), "pool_entries": int(entries[i]),
                                "schedules": int(evals[i])} for i in top],
           "per_hour": per_hour, "report": {"served": served, "requests_in_main_horizon": total, "service_rate_pct": round(100.0 * served / max(total, 1), 2)}}
    line = json.dumps(out)
    print(line)
    if a.out:
        open(a.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
