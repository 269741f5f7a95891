#!/usr/bin/env python
"""bench.py — OSP pool-build throughput on B200 (BASELINE.json metric: schedules evaluated/s and
pool-build ms/epoch, 2,000 vehicles, capacity 4, Manhattan-scale synthetic demand, 30 s epochs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path (dispatch_osp.compute_candidate_veh_trip_pairs, cut-off
disabled) over one recorded epoch snapshot (bench_data/cfg2_epoch*.npz, produced by running the
reference simulator on the synthetic table, see bench_data/make_snapshots.py).  The unit, a
"schedule evaluated", is one execution of scheduling.test_constraints_get_cost AS THE REFERENCE
WOULD PERFORM IT; the kernels count it with the reference's early-exit semantics and the tests
check the count against the oracle, so GPU and CPU arms share the numerator.

N > 1 (torchrun, one rank per GPU): weak scaling — the fleet is N x 2,000 vehicles (the snapshot's
fleet N times against the same request set, further copies in a seeded vehicle order; BASELINE.json
configs[3] at N=4), vehicles are interleaved over ranks, every rank builds its slice and the
vehicle-major pool is assembled on every rank over NVLink peer memory (amod_b200/dist.py PeerPool:
bulk peer stores + epoch flags, no host-side collective in the timed region).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

print_json = None
try:
    METRIC = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
except Exception:
    METRIC = "schedules evaluated/sec; OSP pool-build ms/epoch (2000 veh cap4) at 1/2/4/8 B200"
UNIT = "schedules/s"


def load_snapshots(tag="cfg2"):
    from amod_b200.marshal import Epoch
    files = sorted(glob.glob(os.path.join(ROOT, "bench_data", f"{tag}_epoch*.npz")))
    if not files:
        raise SystemExit("bench_data/ holds no epoch snapshots (run bench_data/make_snapshots.py)")
    snaps = []
    for f in files:
        z = np.load(f)
        snaps.append((os.path.basename(f), Epoch.from_npz_dict(z, "ep_"), int(z["n_evals"]), int(z["n_lookups"])))
    return snaps


def tile_fleet(ep, n):
    """Fleet of n x V vehicles on the same requests: the snapshot's fleet n times, every further copy in a
    different (seeded) vehicle order, so that an interleaved rank slice is a sample of the whole fleet and not
    the same V/n vehicles n times over."""
    if n == 1:
        return ep
    from amod_b200.marshal import Epoch
    parts = [ep] + [ep.take(np.random.default_rng(1000 + c).permutation(ep.n_veh)) for c in range(1, n)]
    kw = {f: getattr(ep, f) for f in Epoch.FIELDS}
    for f in ("veh_nid", "veh_t_to_nid", "veh_load", "veh_status", "veh_reb_node", "veh_reb_ddl", "sche_code", "sche_onboard"):
        kw[f] = np.concatenate([getattr(p, f) for p in parts])
    lens = np.concatenate([np.diff(p.veh_sche_off) for p in parts])
    off = np.zeros(lens.size + 1, dtype=np.int32)
    np.cumsum(lens, out=off[1:])
    kw["veh_sche_off"] = off
    return Epoch(T=ep.T, cap=ep.cap, mode=ep.mode, **kw)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_threads() -> int:
    """All host threads this process may use (torchrun pins OMP_NUM_THREADS=1; the CPU arm must not inherit that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_leg(snaps, tt, budget_s=10.0):
    """Oracle (the C port of the reference's path) on the host cores, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    threads = host_threads()
    # warm (thread pool, table pages), calibrate on every 16th vehicle, then whole snapshots for ~budget_s
    t = time.perf_counter()
    p = oracle.osp_build(snaps[0][1], tt, n_threads=threads, veh_step=16)
    rate = max(p.stats["n_evals"], 1) / max(time.perf_counter() - t, 1e-6)
    evals, secs, n_done = 0, 0.0, 0
    while secs < budget_s and n_done < 5000:
        name, ep, n_evals, _ = snaps[n_done % len(snaps)]
        step = 1 if n_evals / rate < budget_s else int(np.ceil(n_evals / (rate * budget_s)))
        t = time.perf_counter()
        p = oracle.osp_build(ep, tt, n_threads=threads, veh_step=step)
        secs += time.perf_counter() - t
        evals += p.stats["n_evals"]
        n_done += 1
    mean_evals = float(np.mean([s[2] for s in snaps]))
    return {"value": evals / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_done} snapshot builds (all vehicles) in {secs:.2f} s, {evals} schedules, "
                      f"oracle/amod_oracle.c with OpenMP over vehicles",
            "ms_per_epoch": 1e3 * mean_evals / (evals / secs)}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port: the Python reference cannot travel
    to the GPU box) on all host threads, same snapshots, each step a bounded vehicle sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from amod_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    tt = synth.make_road_graph(seed=0).tt
    snaps = load_snapshots()
    threads = host_threads()
    name, ep, n_evals, _ = snaps[0]
    t = time.perf_counter()
    p = oracle.osp_build(ep, tt, n_threads=threads, veh_step=64)
    rate = max(p.stats["n_evals"], 1) / max(time.perf_counter() - t, 1e-6)
    per_step_budget = max(2.0, min(20.0, 150.0 / max(args.steps + args.warmup, 1)))
    step = int(max(1, min(64, np.ceil(n_evals / (rate * per_step_budget)))))
    evals, t_total = 0, 0.0
    for i in range(args.warmup + args.steps):
        _n, e, _ne, _nl = snaps[i % len(snaps)]
        t = time.perf_counter()
        p = oracle.osp_build(e, tt, n_threads=threads, veh_lo=i % step, veh_step=step)
        dt = time.perf_counter() - t
        if i >= args.warmup:
            evals += p.stats["n_evals"]; t_total += dt
    val = evals / t_total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "OSP pool build, 2000 veh cap 4, Manhattan-scale synthetic epoch snapshots (bench_data/cfg2_*)",
                       "sample": f"every {step}th vehicle per step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"every {step}th vehicle of each snapshot, {args.steps} steps, oracle/amod_oracle.c (OpenMP); "
                                       f"the Python reference itself measured 45-71k schedules/s/core (BASELINE.md)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print_json(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything else (NCCL banner, warnings) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global print_json
    print_json = lambda obj: os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from amod_b200 import capi, synth
    from amod_b200.pool import PoolBuilder

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the pool builder has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1

    tt = synth.make_road_graph(seed=0).tt
    snaps = load_snapshots()
    b = PoolBuilder(tt, device=local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    b.set_stream(stream.cuda_stream)          # CUDA events below bracket the stream the kernels run on

    # per-rank slice of the (tiled) fleet, resident in HBM before the timed region
    work = []
    for name, ep, n_evals, n_lookups in snaps:
        full = tile_fleet(ep, world)
        mine, sel = full.shard(rank, world)
        tens = {f: torch.from_numpy(getattr(mine, f)).to(dev) for f in capi.EPOCH_PTR_FIELDS}
        work.append(dict(name=name, host=mine, dev=tens, sel=sel, n_evals_full=n_evals * world, n_veh_full=full.n_veh))

    class DevEp:
        pass

    def dev_epoch(w):
        d = DevEp()
        h = w["host"]
        d.n_veh, d.n_req, d.n_search, d.cap, d.mode, d.T = h.n_veh, h.n_req, h.n_search, h.cap, h.mode, h.T
        for f, t in w["dev"].items():
            setattr(d, f, t)
        return d

    peer = {}

    def allgather_pool():
        """Assemble the fleet-wide vehicle-major pool on every rank: NCCL all-gather of the per-vehicle
        counts (a few KB), then each rank's scatter kernel writes its entries into all ranks' global pools
        with peer stores over NVLink, then a barrier (amod_b200/dist.py PeerPool)."""
        from amod_b200 import dist as adist
        if "pp" not in peer:
            P, NT, NS = b.pool_sizes()
            peer["pp"] = adist.PeerPool(b, world, rank, work[0]["n_veh_full"], 3 * world * P, 3 * world * NT + 1024,
                                        3 * world * NS + 1024, dev)
        peer["pp"].assemble()

    ptr_of = lambda t: t.data_ptr()

    def step_resident(i):
        w = work[i % len(work)]
        st = b.build(dev_epoch(w), loc=capi.LOC_DEVICE, ptr_of=ptr_of)
        if world > 1:
            allgather_pool()
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm (value) ---------------------------------------------------------------
    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    barrier()                       # every rank enters the timed region together
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evals = launches = entries = 0
    ev0.record(stream)
    for i in range(args.steps):
        st = step_resident(args.warmup + i)
        evals += st["n_evals"]; launches += st["n_kernel_launches"]; entries += st["n_entries"]
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    # dominant-kernel time: the same steps once more with plain stream launches, where every k_eval launch is
    # bracketed by CUDA events on the build stream (a captured graph cannot be timed per node)
    b.set_graph(False)
    ms_eval, n_eval_launches, lookups_eval_p = 0.0, 0, 0
    for i in range(min(args.steps, 2 * len(work))):
        st = b.build(dev_epoch(work[(args.warmup + i) % len(work)]), loc=capi.LOC_DEVICE, ptr_of=ptr_of)
        ms_eval += st["ms_eval"]; n_eval_launches += st["n_eval_launches"]
        lookups_eval_p += st["n_lookups_eval"]
    b.set_graph(True)
    prof_steps = min(args.steps, 2 * len(work))
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(evals), float(launches), float(entries)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t_ms.item())
    evals_all, launches_all, entries_all = (float(x) for x in tot.tolist())

    # ---- end-to-end arm: C-ABI call with HOST buffers, H2D + D2H inside the timed region ------------
    def step_e2e(i):
        w = work[i % len(work)]
        pool = b.build_pool(w["host"], reuse=True)
        if world > 1:
            allgather_pool()
        h2d = sum(getattr(w["host"], f).nbytes for f in capi.EPOCH_PTR_FIELDS)
        d2h = sum(getattr(pool, f).nbytes for f, _d, _k in capi.POOL_FIELDS)
        return pool.stats["n_evals"], h2d, d2h

    for i in range(min(args.warmup, 2)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    e_evals = e_h2d = e_d2h = 0
    for i in range(args.steps):
        a, h, d = step_e2e(args.warmup + i)
        e_evals += a; e_h2d += h; e_d2h += d
    barrier()
    e_s = time.perf_counter() - t0
    t_e = torch.tensor([e_s], dtype=torch.float64, device=dev)
    tot_e = torch.tensor([float(e_evals)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_e, op=dist.ReduceOp.SUM)
    e2e_val = float(tot_e.item()) / float(t_e.item())
    clocks = sampler.stop() if sampler else None       # sampled across both timed regions (resident and end-to-end)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_eval: the insertion search) -----------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # L2 random-gather peak of the table on this box (the bound SURVEY.md §8d names), measured live
    rng = np.random.default_rng(0)
    ng = 1 << 24
    o_t = torch.from_numpy(rng.integers(1, tt.shape[0] + 1, size=ng).astype(np.int32)).to(dev)
    d_t = torch.from_numpy(rng.integers(1, tt.shape[0] + 1, size=ng).astype(np.int32)).to(dev)
    out_t = torch.empty(ng, dtype=torch.float32, device=dev)
    import ctypes as C
    gms = C.c_double(0.0)
    b._check(b.lib.amod_gather_bench(b.h, o_t.data_ptr(), d_t.data_ptr(), out_t.data_ptr(), ng, 5, C.byref(gms)))
    l2_gather_gbs = ng * 32 / (gms.value * 1e-3) / 1e9
    alg_bytes = lookups_eval_p * 32.0          # one 32 B L2 sector per reference lookup
    achieved = alg_bytes / (ms_eval * 1e-3) / 1e9 if ms_eval > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "eval_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_eval (insertion search, count + fill launches)", "achieved": achieved,
                "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(n_eval_launches, 1), "launches": n_eval_launches,
                "avg_launch_ms": ms_eval / max(n_eval_launches, 1), "kernel_share_of_step": (ms_eval / prof_steps) / (ms / args.steps),
                "l2_gather_peak_gbs": l2_gather_gbs, "frac_of_l2_gather_peak": achieved / l2_gather_gbs,
                "note": "algorithmic bytes = reference travel-time lookups inside evaluations x 32 B sector; the table is "
                        "L2-resident so HBM peak is the fallback bound and the measured L2 random-gather rate the real one"}

    cpu = None if (args.no_cpu or world > 1) else cpu_leg(snaps, tt)
    line = {"metric": METRIC, "value": evals_all / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"OSP pool build (cut-off disabled), {work[0]['n_veh_full']} veh cap {work[0]['host'].cap}, "
                                   f"Manhattan-scale synthetic epoch snapshots x{len(work)} (bench_data/cfg2_*), 4091-node fp32 table",
                       "fleet": work[0]["n_veh_full"], "requests_considered": int(np.mean([w["host"].n_search for w in work])),
                       "parallelism": f"vehicles interleaved over {world} rank(s)" + (", NCCL all-gather of per-vehicle counts + peer-store scatter of the pool over NVLink" if world > 1 else ""),
                       "l2_policy": "inputs differ every step (8 snapshots cycled); table deliberately L2-resident (67 MB persisting window)",
                       "pool_entries_per_step": entries_all / args.steps, "schedules_per_step": evals_all / args.steps},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": e_h2d // args.steps, "d2h_bytes_per_step": e_d2h // args.steps,
                    "ms_per_step": 1e3 * float(t_e.item()) / args.steps,
                    "path": "amod_osp_build(host pointers) + amod_pool_fetch(host buffers) through ctypes"},
            "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roofline}
    if cpu:
        line["cpu_baseline"] = cpu
    print_json(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
