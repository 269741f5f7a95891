"""cpu_baseline.python_reference — the reference's OWN Python implementation of the hot path, timed on the GPU box.

The UNMODIFIED reference sources travel to the box as a git-ignored scratch copy under baseline/_ref/src (made by
__graft_entry__.build() in the build container, where /root/reference exists; nothing of it is committed).  This
module (bench.py's cpu_baseline leg only) imports that copy through oracle/ref_harness.py (stub modules for
matplotlib / gurobipy / the value network, synthetic data files), rebuilds reference Veh/Req objects from one bench
snapshot and times dispatch_osp.compute_candidate_veh_trip_pairs (cut-off disabled) on a bounded vehicle sample:
  * one core, as shipped (a single Python process);
  * all host cores: a multiprocessing (fork) pool mapping the same seam over vehicle chunks.
The numerator (schedules evaluated) of the sample is counted by the instrumented C oracle on the same vehicles, and
the number of records the reference returns is checked against the oracle's pool for them.

    python baseline/ref_arm.py --budget 20      -> one JSON line
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "dispatcher"))


def time_reference_python(root: str = ROOT, budget_s: float = 20.0) -> dict:
    """Runs the measurement in a subprocess (the reference freezes its config at import and star-imports the world)."""
    if not available():
        return {"unavailable": "baseline/_ref/src is not populated (run __graft_entry__.build() where /root/reference exists)"}
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--budget", str(budget_s)], capture_output=True, text=True,
                         timeout=600, env=env, cwd=root)
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    return {"unavailable": ("reference run failed: " + (out.stderr.strip().splitlines() or ["?"])[-1])[:240]}


_G = {}


def _chunk_worker(args):
    lo, hi, step = args
    g = _G
    vehs = g["vehs"][lo:hi:step] if step > 1 else g["vehs"][lo:hi]
    recs = g["ref"].dispatch_osp.compute_candidate_veh_trip_pairs(g["rids"], g["rids"], g["reqs"], vehs, g["T"], 1e9, True, False)
    return len(recs)


def _main(budget_s: float):
    import glob
    import types

    import numpy as np
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    os.environ["AMOD_REFERENCE_ROOT"] = REF
    from amod_b200 import marshal, synth
    import oracle
    import ref_harness as rh
    rh.REFERENCE_ROOT = REF
    graph = synth.make_road_graph(seed=0)
    light = types.SimpleNamespace(tt=graph.tt, dist=np.zeros((2, 2), np.float32), pred=np.zeros((2, 2), np.int32), lng=graph.lng,
                                  lat=graph.lat, stations=graph.stations, n_nodes=graph.n_nodes)
    f = sorted(glob.glob(os.path.join(ROOT, "bench_data", "cfg2_epoch*.npz")))[-2]
    ep = marshal.Epoch.from_npz_dict(np.load(f), "ep_")
    one = np.array([1], dtype=np.int32)
    ref = rh.setup_reference("/tmp/amod_ref_arm", light, one, one + 1, np.array([10 ** 9]), fleet_size=ep.n_veh, capacity=ep.cap,
                             dispatcher="OSP")
    import make_golden
    reqs, objs, vehs = make_golden.epoch_to_objects(ref, ep)
    import src.simulator.config as cfg
    cfg.config_change_vc(ep.cap)
    cfg.config_change_fs(max(ep.n_veh, 1))
    rids = [int(ep.req_id[i]) for i in ep.search]
    seam = ref.dispatch_osp.compute_candidate_veh_trip_pairs
    tt = graph.tt

    def oracle_sample(lo, step):
        p = oracle.osp_build(ep, tt, veh_lo=lo, veh_step=step)
        return int(p.stats["n_evals"]), int(p.n_entries)

    # calibrate on every 64th vehicle, then size the 1-core sample for ~budget/2
    t = time.perf_counter()
    n_rec = len(seam(rids, rids, reqs, vehs[3::64], ep.T, 1e9, True, False))
    dt = time.perf_counter() - t
    ev, ent = oracle_sample(3, 64)
    assert n_rec == ent, (n_rec, ent)
    rate = max(ev, 1) / max(dt, 1e-6)
    total_ev, _ = oracle_sample(0, 1)
    step = int(max(1, np.ceil(total_ev / (rate * budget_s / 2))))
    t = time.perf_counter()
    n_rec = len(seam(rids, rids, reqs, vehs[0::step], ep.T, 1e9, True, False))
    dt1 = time.perf_counter() - t
    ev1, ent1 = oracle_sample(0, step)
    out = {"value": ev1 / dt1, "unit": "schedules/s", "cores": 1, "kind": "reference",
           "sample": f"{os.path.basename(f)}: every {step}th vehicle ({len(vehs[0::step])} of {ep.n_veh}), {ev1} schedules in {dt1:.1f} s, "
                     f"dispatch_osp.compute_candidate_veh_trip_pairs as shipped (cut-off disabled), one Python process",
           "records_match_oracle": bool(n_rec == ent1)}
    # all cores: fork pool over contiguous vehicle chunks of the same sample
    import multiprocessing as mp
    cores = max(1, len(os.sched_getaffinity(0)))
    _G.update(ref=ref, vehs=vehs, rids=rids, reqs=reqs, T=ep.T)
    mstep = int(max(1, np.ceil(total_ev / (rate * cores * 0.6 * budget_s / 2))))
    sample = list(range(0, ep.n_veh, mstep))
    # interleave the sampled vehicles over 8 x cores chunks (work per vehicle is heavy-tailed)
    nchunk = min(len(sample), 8 * cores)
    _G["vehs"] = [vehs[i] for i in sample]
    jobs = [(c, len(sample), nchunk) for c in range(nchunk)]
    ctx = mp.get_context("fork")
    t = time.perf_counter()
    with ctx.Pool(cores) as pool:
        counts = pool.map(_chunk_worker, jobs, chunksize=1)
    dtm = time.perf_counter() - t
    evm, entm = oracle_sample(0, mstep)
    out["all_cores"] = {"value": evm / dtm, "unit": "schedules/s", "cores": cores,
                        "sample": f"every {mstep}th vehicle ({len(sample)}), {evm} schedules in {dtm:.1f} s, multiprocessing fork pool over "
                                  f"{nchunk} interleaved vehicle chunks", "records_match_oracle": bool(sum(counts) == entm)}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=float, default=20.0)
    _main(ap.parse_args().budget)
