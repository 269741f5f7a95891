#!/usr/bin/env python
"""Steady-state profile driver for ncu (profiles/r02_*): builds every snapshot of a config twice to warm up (graph
captured, buffers grown, table resident in L2), then brackets ONE pass over the snapshots with cudaProfilerStart/Stop
and writes, for exactly those launches, the per-level algorithmic work (reference evaluations / lookups per trip
level = per k_eval<count>/<fill> launch pair) to a JSON file, so that ncu's per-launch sectors and bytes can be
divided by the algorithmic lookups of the SAME launches.

    ncu --profile-from-start off --cache-control none --clock-control none -k regex:k_eval --metrics ... \
        python scripts/prof_eval.py --out gpurun_out/r02_eval_work.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "eval_work.json"))
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--fleet-mult", type=int, default=1)
    args = ap.parse_args()
    import torch
    import bench
    from amod_b200 import capi, synth
    from amod_b200.pool import PoolBuilder
    tt = synth.make_road_graph(seed=0).tt
    snaps = bench.load_snapshots(args.config)
    dev = torch.device("cuda", 0)
    b = PoolBuilder(tt, device=0)
    if args.no_graph:
        b.set_graph(False)

    class DevEp:
        pass

    eps = []
    for name, ep, _ne, _nl, _extra in snaps:
        ep = synth.tile_fleet(ep, args.fleet_mult)
        d = DevEp()
        d.n_veh, d.n_req, d.n_search, d.cap, d.mode, d.T = ep.n_veh, ep.n_req, ep.n_search, ep.cap, ep.mode, ep.T
        d.keep = {f: torch.from_numpy(getattr(ep, f)).to(dev) for f in capi.EPOCH_PTR_FIELDS}
        for f, t in d.keep.items():
            setattr(d, f, t)
        eps.append((name, d))
    ptr_of = lambda t: t.data_ptr()
    for _ in range(2):
        for _name, d in eps:
            b.build(d, loc=capi.LOC_DEVICE, ptr_of=ptr_of)
    torch.cuda.synchronize()
    rows = []
    torch.cuda.profiler.start()
    for _ in range(args.passes):
        for name, d in eps:
            st = b.build(d, loc=capi.LOC_DEVICE, ptr_of=ptr_of)
            rows.append({"snapshot": name, "n_evals": st["n_evals"], "n_lookups_eval": st["n_lookups_eval"], "n_levels": st["n_levels"],
                         "n_eval_launches": st["n_eval_launches"], "n_kernel_launches": st["n_kernel_launches"], "ms_build": st["ms_build"],
                         "evals_level": st["evals_level"], "lookups_level": st["lookups_level"], "n_attempts": st["n_attempts"],
                         "n_chunks": st["n_chunks"], "n_items": st["n_items"]})
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"config": args.config, "graph": not args.no_graph, "fleet_mult": args.fleet_mult, "steps": rows}, open(args.out, "w"), indent=1)
    print(f"{len(rows)} profiled steps, mean ms_build {np.mean([r['ms_build'] for r in rows]):.3f}")


if __name__ == "__main__":
    main()
