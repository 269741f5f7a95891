"""Per-stage cost of one DEVICE-RESIDENT epoch (amod_b200/fleet.py: advance -> arrivals -> pool build -> greedy assignment ->
apply -> NPO -> apply) at cfg2 scale: 2,000 vehicles cap 4, ~137 arrivals per 30 s epoch (README.md:9), OSP + greedy + NPO.
Host wall clock around each call (every call ends with its own stream synchronisation).  Run on the GPU box:

    python scripts/prof_fleet.py --epochs 150 --out gpurun_out/fleet_epoch.json
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
from amod_b200.fleet import DeviceFleetSim  # noqa: E402
from amod_b200.pool import PoolBuilder  # noqa: E402


def measure(veh=2000, cap=4, epochs=150, per_epoch=137.0, tail=60, graph=None, device=0):
    graph = graph if graph is not None else synth.make_road_graph(seed=0, with_paths=True)
    V, E = veh, epochs
    start = graph.stations[(np.arange(V) * len(graph.stations) // V)].astype(np.int32)
    on, dn, tm = synth.make_request_stream(graph, int(per_epoch * (E + 2)), 0, 30 * (E + 2), seed=3, mean_trip_sec=600.0)
    b = PoolBuilder(graph.tt, device=device)
    sim = DeviceFleetSim(b, graph, start, max_stops=24, max_route_steps=4096, request_capacity=1 << 17)
    stages = ("advance", "arrivals", "build", "greedy", "apply", "rebalance", "epoch")
    ms = {s: [] for s in stages}
    extra = {k: [] for k in ("n_search", "entries", "evals", "selected", "rebalanced", "events", "h2d_bytes")}
    n_req = 0
    for e in range(E):
        T = 30 * (e + 1)
        t0 = time.perf_counter()
        ev = sim.advance(T, True)
        t1 = time.perf_counter()
        m = np.flatnonzero((tm >= T - 30) & (tm < T))
        h2d = 0
        if m.size:
            ts = graph.tt[on[m] - 1, dn[m] - 1].astype(np.float64)
            tr_ = tm[m].astype(np.float64)
            sim.add_requests(n_req, on[m], dn[m], tr_, ts, tr_ + 420, (tr_ + ts) + 840)
            n_req += m.size
            h2d = m.size * (4 + 4 + 8 * 4)
        t2 = time.perf_counter()
        st = sim.build(marshal.MODE_OSP, cap, T, m.size)
        t3 = time.perf_counter()
        _order, sel = b.greedy_assign()
        t4 = time.perf_counter()
        sim.apply_selected(None, True, True)
        t5 = time.perf_counter()
        nreb = sim.rebalance(T)
        t6 = time.perf_counter()
        for s, d in zip(stages, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t6 - t0)):
            ms[s].append(1e3 * d)
        for k, v in zip(extra, (sim.last_n_search, st["n_entries"], st["n_evals"], sel.size, nreb, ev["veh"].size, h2d)):
            extra[k].append(int(v))
    k = min(tail, E)
    status = sim.array("status"); rs = sim.array("req_status")
    out = {"workload": f"device-resident epoch (amod_fleet_*: advance -> arrivals -> OSP pool build -> greedy assignment -> apply -> NPO -> apply), "
                       f"{V} veh cap {cap}, {per_epoch:g} arrivals/epoch, {E} epochs (means over the last {k}; host wall clock per call)",
           "ms": {s: float(np.mean(v[-k:])) for s, v in ms.items()},
           "ms_p95": {s: float(np.percentile(v[-k:], 95)) for s, v in ms.items()},
           "per_epoch": {s: float(np.mean(v[-k:])) for s, v in extra.items()},
           "fleet_at_end": {"idle": int((status == 1).sum()), "working": int((status == 2).sum()), "rebalancing": int((status == 3).sum())},
           "requests_at_end": {n: int((rs == i).sum()) for i, n in enumerate(("pending", "picking", "onboard", "complete", "walkaway"), start=1)}}
    b.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--veh", type=int, default=2000)
    ap.add_argument("--cap", type=int, default=4)
    ap.add_argument("--epochs", type=int, default=150)
    ap.add_argument("--per-epoch", type=float, default=137.0)
    ap.add_argument("--tail", type=int, default=60, help="epochs averaged (the last ones: steady state)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    line = json.dumps(measure(a.veh, a.cap, a.epochs, a.per_epoch, a.tail))
    print(line)
    if a.out:
        open(a.out, "w").write(line + "\n")


if __name__ == "__main__":
    main()
