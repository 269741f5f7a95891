#!/usr/bin/env python
"""bench.py — OSP pool-build throughput on B200 (BASELINE.json metric: schedules evaluated/s and
pool-build ms/epoch, 2,000 vehicles, capacity 4, Manhattan-scale synthetic demand, 30 s epochs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config cfg2|cfg3|cfg4|cfg5] [--fleet-mult M] [--scaling weak|strong]

A *step* is one pass of the hot path (dispatch_osp.compute_candidate_veh_trip_pairs, cut-off
disabled) over one recorded epoch snapshot (bench_data/<cfg>_epoch*.npz, produced by running the
reference simulator on the synthetic table, see bench_data/make_snapshots.py).  The unit, a
"schedule evaluated", is one execution of scheduling.test_constraints_get_cost AS THE REFERENCE
WOULD PERFORM IT; the kernels count it with the reference's early-exit semantics and the tests
check the count against the oracle, so GPU and CPU arms share the numerator.

Workloads (BASELINE.json configs[1..4]; configs[0] is the CPU parity run in tests/):
  cfg2 (default)  OSP, 2,000 vehicles cap 4, 8 snapshots.  N > 1: WEAK scaling, the fleet is N x 2,000 vehicles
                  (synth.tile_fleet) against the same requests.  --fleet-mult M fixes the fleet at M x 2,000 instead
                  (single-GPU size sweep, or STRONG scaling with --gpus N --scaling strong).
  cfg3            the same with capacity 8 (large permutation search).
  cfg4            OSP scale-out, 8,000 vehicles cap 4 (the cfg2 fleet x4): STRONG scaling, the fleet is sharded over N ranks.
                  (The survey's "demand x4" reading cannot be built exhaustively: it holds trips of size cap+4, where the
                  reference raises IndexError; its 100-vehicle sample is a parity test, bench_data/cfg4dense_*.)
  cfg5            peak epoch: SBA vehicle-request matrix 20,000 x 5,000 + NPO nearest-idle-vehicle match.

N > 1 (torchrun, one rank per GPU): vehicles are sharded over the ranks, every rank builds its slice and the
vehicle-major pool is assembled over NVLink peer memory on the rank that hosts the (unchanged) assigner
(amod_b200/dist.py PeerPool; no host-side collective in the timed region).

Parity is part of the line: before timing, every snapshot's pool (at N > 1 the ASSEMBLED pool on rank 0) is
fingerprinted (marshal.pool_digest) and compared with the oracle's fingerprint stored in bench_data/digests.json
(bench_data/make_digests.py) -> "parity": {"checked": true, "ok": ..., "snapshots": n}.
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

CONFIGS = {
    "cfg2": dict(tag="cfg2", scaling="weak", what="OSP pool build (cut-off disabled), {fleet} veh cap 4, Manhattan-scale synthetic epoch "
                 "snapshots x{n} (bench_data/cfg2_*), 4091-node fp32 table"),
    "cfg3": dict(tag="cfg3", scaling="weak", what="OSP pool build (cut-off disabled), {fleet} veh cap 8 (large permutation search), "
                 "synthetic epoch snapshots x{n} (bench_data/cfg3_*), 4091-node fp32 table"),
    "cfg4": dict(tag="cfg2", scaling="strong", mult=4, what="OSP scale-out, {fleet} veh cap 4 (the cfg2 epochs' fleet x4, synth.tile_fleet), synthetic "
                 "epoch snapshots x{n} (bench_data/cfg2_*), vehicle-sharded, 4091-node fp32 table"),
    "cfg5": dict(tag="cfg5", scaling="strong", what="peak epoch: SBA feasibility/cost matrix {fleet} veh x {req} new requests + NPO "
                 "nearest-idle-vehicle match (bench_data/cfg5_*), 4091-node fp32 table"),
}


def load_digests():
    try:
        return json.load(open(os.path.join(ROOT, "bench_data", "digests.json")))
    except Exception:
        return {}


def load_snapshots(tag="cfg2"):
    from amod_b200.marshal import Epoch
    files = sorted(glob.glob(os.path.join(ROOT, "bench_data", f"{tag}_epoch*.npz")))
    if not files:
        raise SystemExit(f"bench_data/ holds no {tag} epoch snapshots (run bench_data/make_snapshots.py)")
    snaps = []
    for f in files:
        z = np.load(f)
        extra = {k: z[k] for k in ("npo_idle_nid", "npo_pend_onid") if k in z.files}
        snaps.append((os.path.basename(f), Epoch.from_npz_dict(z, "ep_"), int(z["n_evals"]), int(z["n_lookups"]), extra))
    return snaps


def tile_fleet(ep, n):
    from amod_b200 import synth
    return synth.tile_fleet(ep, n)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []          # (perf_counter at arrival, fields)
        self.proc = None
        self.t_mark = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def wait_first(self, timeout_s: float = 8.0):
        """nvidia-smi needs a moment to come up (longer on a fresh box): the timed region starts once it is streaming."""
        t = time.perf_counter()
        while not self.rows and time.perf_counter() - t < timeout_s:
            time.sleep(0.02)

    def mark(self):
        self.t_mark = time.perf_counter()

    def needs_sample(self) -> bool:
        """True while no row has arrived since mark() (for at most 1.5 s after the first question)."""
        if any(ts >= (self.t_mark or 0.0) for ts, _r in self.rows):
            return False
        if not hasattr(self, "_asked"):
            self._asked = time.perf_counter()
        self.extended = True
        return time.perf_counter() - self._asked < 1.5

    def stop(self, load_fn=None) -> dict:
        """Rows that arrived since mark().  The timed regions of a short run (K x 0.5 ms) can be shorter than the 100 ms sampling
        period: then the SAME steps keep running (untimed) until one row has arrived, so that the clocks are always read under the
        load that was timed, and the JSON says so."""
        t_mark = self.t_mark if self.t_mark is not None else 0.0
        inside = [r for ts, r in self.rows if ts >= t_mark]
        how = "during the timed regions"
        if not inside and load_fn is not None:
            while self.needs_sample():
                load_fn()
            inside = [r for ts, r in self.rows if ts >= t_mark]
        if getattr(self, "extended", False):
            how = "under the same steps (untimed) right after the timed regions, which were shorter than the 100 ms sampling period"
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in inside if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in inside if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "sampled": how}


def host_threads() -> int:
    """All host threads this process may use (torchrun pins OMP_NUM_THREADS=1; the CPU arm must not inherit that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def _oracle_build(oracle, ep, tt, **kw):
    from amod_b200 import marshal
    if ep.mode == marshal.MODE_SBA:
        kw.pop("veh_lo", None); kw.pop("veh_step", None)
        return oracle.sba_build(ep, tt, n_threads=kw.get("n_threads", 1))
    return oracle.osp_build(ep, tt, **kw)


def cpu_leg(work_eps, tt, budget_s=10.0):
    """Oracle (the C port of the reference's path) on the host cores, bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    from amod_b200 import marshal
    threads = host_threads()
    sba = work_eps[0].mode == marshal.MODE_SBA
    _oracle_build(oracle, work_eps[0], tt, n_threads=threads, veh_step=16)      # warm (thread pool, table pages)
    evals, secs, n_done = 0, 0.0, 0
    while secs < budget_s and n_done < 5000:
        ep = work_eps[n_done % len(work_eps)]
        t = time.perf_counter()
        p = _oracle_build(oracle, ep, tt, n_threads=threads)
        secs += time.perf_counter() - t
        evals += p.stats["n_evals"]
        n_done += 1
    return {"value": evals / secs, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_done} snapshot builds (all vehicles) in {secs:.2f} s, {evals} schedules, "
                      f"oracle/amod_oracle.c with OpenMP over {'requests' if sba else 'vehicles'}",
            "ms_per_epoch": 1e3 * secs / max(n_done, 1)}


def python_reference_leg(budget_s=20.0):
    """The reference's OWN Python implementation (git-ignored scratch copy under baseline/_ref, shipped to the box by
    gpurun), one core, on a bounded vehicle sample of the first snapshot; None when the copy is not there."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline"))
        import ref_arm
    except Exception:
        return None
    try:
        return ref_arm.time_reference_python(ROOT, budget_s=budget_s)
    except Exception as e:      # the baseline is a report, never a reason to lose the bench line
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def workload_for(args, world):
    cfg = CONFIGS[args.config]
    scaling = args.scaling or cfg["scaling"]
    mult = args.fleet_mult if args.fleet_mult else (world if scaling == "weak" else cfg.get("mult", 1))
    return cfg, scaling, mult


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the Python reference itself is timed on one core
    as `python_reference` inside cpu_baseline) on all host threads, same workload as our arm at this N (the same
    tiled fleet), each step a bounded vehicle sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from amod_b200 import marshal, synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    cfg, scaling, mult = workload_for(args, max(world, args.gpus))
    tt = synth.make_road_graph(seed=0).tt
    snaps = load_snapshots(cfg["tag"])
    eps = [tile_fleet(s[1], mult) for s in snaps]
    threads = host_threads()
    sba = eps[0].mode == marshal.MODE_SBA
    t = time.perf_counter()
    p = _oracle_build(oracle, eps[0], tt, n_threads=threads, veh_step=64)
    rate = max(p.stats["n_evals"], 1) / max(time.perf_counter() - t, 1e-6)
    dg0 = load_digests().get(f"{snaps[0][0]}@x{mult}", {})
    n_evals = dg0.get("n_evals", 0) * dg0.get("sample_step", 1) or snaps[0][2] * mult
    per_step_budget = max(2.0, min(20.0, 150.0 / max(args.steps + args.warmup, 1)))
    step = 1 if sba else int(max(1, min(64 * mult, np.ceil(n_evals / (rate * per_step_budget)))))
    evals, t_total = 0, 0.0
    for i in range(args.warmup + args.steps):
        e = eps[i % len(eps)]
        t = time.perf_counter()
        p = _oracle_build(oracle, e, tt, n_threads=threads, veh_lo=i % step, veh_step=step)
        dt = time.perf_counter() - t
        if i >= args.warmup:
            evals += p.stats["n_evals"]; t_total += dt
    val = evals / t_total
    fleet = eps[0].n_veh
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["what"].format(fleet=fleet, n=len(eps), req=eps[0].n_search), "fleet": fleet,
                       "sample": f"every {step}th vehicle per step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"every {step}th vehicle of each snapshot, {args.steps} steps, oracle/amod_oracle.c (OpenMP), the C "
                                       f"restatement of the reference's path; the Python reference itself: see cpu_baseline.python_reference "
                                       f"of our arm (BASELINE.md)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print_json(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--fleet-mult", type=int, default=0, help="fleet = M x the snapshot's fleet (default: N for weak scaling, 1 for strong)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="plain stream launches instead of the captured epoch graph (profiling)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything else (NCCL banner, warnings) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global print_json
    print_json = lambda obj: os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    if args.impl == "reference":
        return run_reference(args)

    import ctypes as C

    import torch
    import torch.distributed as dist
    from amod_b200 import capi, marshal, synth
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

    cfg, scaling, mult = workload_for(args, world)
    tt = synth.make_road_graph(seed=0).tt
    snaps = load_snapshots(cfg["tag"])
    digests = load_digests()
    b = PoolBuilder(tt, device=local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    b.set_stream(stream.cuda_stream)          # CUDA events below bracket the stream the kernels run on
    if args.no_graph:
        b.set_graph(False)
    sba = snaps[0][1].mode == marshal.MODE_SBA

    fulls = [(name, tile_fleet(ep, mult), extra) for name, ep, _ne, _nl, extra in snaps]
    n_veh_full = fulls[0][1].n_veh
    peer = {}
    placement = {"veh_rank": None, "how": "interleaved (v % N)"}

    def make_work(sels):
        """per-rank slice of the (tiled) fleet, resident in HBM before the timed region"""
        work = []
        for name, full, extra in fulls:
            mine = full.take(sels[rank]) if sels is not None else full.shard(rank, world)[0]
            tens = {f: torch.from_numpy(getattr(mine, f)).to(dev) for f in capi.EPOCH_PTR_FIELDS}
            w = dict(name=name, host=mine, dev=tens, n_veh_full=full.n_veh, digest=digests.get(f"{name}@x{mult}"))
            if extra:        # cfg5: the NPO match of the same epoch (idle vehicles x pending requests)
                w["npo_host"] = (np.ascontiguousarray(extra["npo_idle_nid"], dtype=np.int32), np.ascontiguousarray(extra["npo_pend_onid"], dtype=np.int32))
                w["npo_dev"] = tuple(torch.from_numpy(a).to(dev) for a in w["npo_host"])
            work.append(w)
        return work

    work = make_work(None)
    if world > 1:
        # Load-balanced placement (work per vehicle is heavy-tailed): one untimed pass with the interleaved
        # placement measures every vehicle's share of the pool (feasible schedules behind its entries); the fleet is then
        # placed by longest-processing-time-first bin packing on the running mean over the epochs seen so far.
        from amod_b200 import dist as adist
        b.build(work[0]["host"])
        P, NT, NS = b.pool_sizes()
        peer["pp"] = adist.PeerPool(b, world, rank, n_veh_full, 3 * world * P, 3 * world * NT + 1024, 3 * world * NS + 1024, dev)
        peer["pp"].attach(root=0)
        cost = np.zeros(n_veh_full)
        for w in work:
            b.build(w["host"])
            if rank == 0:
                got = peer["pp"].fetch()
                cost += np.bincount(got.ent_veh, weights=got.ent_nfeas.astype(np.float64) + 1.0, minlength=n_veh_full)
            dist.barrier()
        vr = torch.zeros(n_veh_full, dtype=torch.int32, device=dev)
        if rank == 0 and not os.environ.get("AMOD_BENCH_INTERLEAVED"):
            # the gathering rank also merges: it takes less than an even share of the build (measured: profiles/r02_root_share_ab.txt)
            root_share = float(os.environ.get("AMOD_BENCH_ROOT_SHARE", {2: 0.7, 4: 0.55, 8: 0.45}.get(world, max(0.3, 1.0 - 0.63 * (world - 1) / world))))
            veh_rank, _vl, _sels = adist.balanced_placement(cost, world, root=0, root_share=root_share)
            vr.copy_(torch.from_numpy(veh_rank))
        elif rank == 0:
            vr.copy_(torch.from_numpy((np.arange(n_veh_full) % world).astype(np.int32)))
        dist.broadcast(vr, src=0)
        veh_rank = vr.cpu().numpy()
        veh_local = np.zeros(n_veh_full, dtype=np.int32)
        sels = []
        for r in range(world):
            sel = np.flatnonzero(veh_rank == r)
            veh_local[sel] = np.arange(sel.size, dtype=np.int32)
            sels.append(sel.astype(np.int64))
        work = make_work(sels)
        peer["pp"].attach(root=0, veh_rank=veh_rank, veh_local=veh_local)
        if not os.environ.get("AMOD_BENCH_INTERLEAVED"):
            placement = {"veh_rank": veh_rank, "how": "balanced: LPT bin packing on each vehicle's mean pool share over the epochs seen (untimed pass); the gathering rank, which also merges, takes 0.7 / 0.55 / 0.45 of an even share at N = 2 / 4 / 8"}

    class DevEp:
        pass

    def dev_epoch(w):
        d = DevEp()
        h = w["host"]
        d.n_veh, d.n_req, d.n_search, d.cap, d.mode, d.T = h.n_veh, h.n_req, h.n_search, h.cap, h.mode, h.T
        for f, t in w["dev"].items():
            setattr(d, f, t)
        return d

    ptr_of = lambda t: t.data_ptr()
    npo_out = {}

    def npo_resident(w):
        idle, pend = w["npo_dev"]
        m = min(idle.numel(), pend.numel())
        if "oi" not in npo_out:
            npo_out.update(oi=np.zeros(max(m, 1), np.int32), op=np.zeros(max(m, 1), np.int32), od=np.zeros(max(m, 1), np.float64))
        n_out, n_rounds = C.c_int32(0), C.c_int32(0)
        b._check(b.lib.amod_npo_match(b.h, idle.data_ptr(), idle.numel(), pend.data_ptr(), pend.numel(), capi.LOC_DEVICE,
                                      npo_out["oi"].ctypes.data, npo_out["op"].ctypes.data, npo_out["od"].ctypes.data,
                                      C.byref(n_out), C.byref(n_rounds)))
        return n_out.value, n_rounds.value

    def step_resident(i):
        w = work[i % len(work)]
        st = b.build(dev_epoch(w), loc=capi.LOC_DEVICE, ptr_of=ptr_of)      # (N > 1: the build also assembles, PeerPool.attach)
        if "npo_dev" in w and rank == 0:      # NPO runs where the assigner lives: its input is the idle vehicles' nodes, a few KB
            npo_resident(w)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity (untimed): every snapshot's pool against the oracle's fingerprint ---------------------------------
    def npo_digest(oi, op, od):
        import hashlib
        h = hashlib.blake2b(digest_size=16)
        for a, dt in ((oi, "<i4"), (op, "<i4"), (od, "<f8")):
            h.update(np.ascontiguousarray(a, dtype=np.dtype(dt)).tobytes())
        return h.hexdigest()

    parity = {"checked": True, "ok": True, "snapshots": 0, "against": "oracle fingerprints in bench_data/digests.json (entries, order, "
              "trips, schedules, float64 cost/delay bits; n_evals)", "pool": "assembled on rank 0" if world > 1 else "local"}
    for i, w in enumerate(work):
        st = step_resident(i)
        ne = torch.tensor([float(st["n_evals"])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ne, op=dist.ReduceOp.SUM)
        if rank == 0:
            got = peer["pp"].fetch() if world > 1 else b.fetch()
            dg = w["digest"]
            if dg is None:
                parity["checked"] = False
                parity.setdefault("missing", []).append(f"{w['name']}@x{mult}")
            else:
                step = int(dg.get("sample_step", 1))         # cfg4: the oracle built every step-th vehicle's part of the pool
                sub = marshal.pool_of_vehicles(got, step)
                ok = marshal.pool_digest(sub) == dg["digest"] and sub.n_entries == dg["n_entries"] and (step > 1 or int(ne.item()) == dg["n_evals"])
                if step > 1:
                    parity["vehicle_sample"] = f"every {step}th vehicle of the fleet ({dg.get('n_veh', 0) // step} vehicles, {dg['n_entries']} pool entries)"
                if "npo_host" in w and "npo_digest" in dg:
                    oi, op, od = b.npo_match(*w["npo_host"])
                    ok = ok and npo_digest(oi, op, od) == dg["npo_digest"]
                parity["ok"] = parity["ok"] and bool(ok)
                if not ok:
                    parity.setdefault("failed", []).append(w["name"])
                parity["snapshots"] += 1
    if not parity["checked"]:
        parity["ok"] = None

    # ---- device-resident arm (value) ---------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_first()
    for i in range(args.warmup):
        step_resident(i)
    if sampler:
        sampler.mark()
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
    ms_eval, n_eval_launches, lookups_eval_p, ms_screen, screens_p, ms_npo, npo_pairs = 0.0, 0, 0, 0.0, 0, 0.0, 0
    prof_steps = min(args.steps, 2 * len(work))
    for i in range(prof_steps):
        w = work[(args.warmup + i) % len(work)]
        st = b.build(dev_epoch(w), loc=capi.LOC_DEVICE, ptr_of=ptr_of)
        ms_eval += st["ms_eval"]; n_eval_launches += st["n_eval_launches"]
        lookups_eval_p += st["n_lookups_eval"]
        ms_screen += st.get("ms_screen", 0.0); screens_p += st["n_screens"]
        if "npo_dev" in w and rank == 0:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _n, rounds = npo_resident(w)
            e1.record(stream)
            torch.cuda.synchronize()
            ms_npo += e0.elapsed_time(e1)
            npo_pairs += w["npo_dev"][0].numel() * w["npo_dev"][1].numel()
    b.set_graph(not args.no_graph)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(evals), float(launches), float(entries)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t_ms.item())
    evals_all, launches_all, entries_all = (float(x) for x in tot.tolist())

    # ---- end-to-end arm: C-ABI call with HOST buffers, H2D + D2H inside the timed region ------------
    # every rank uploads its slice from host memory; the rank hosting the assigner (rank 0) reads the result back:
    # its local pool at N = 1, the ASSEMBLED fleet-wide pool at N > 1
    def step_e2e(i):
        w = work[i % len(work)]
        h2d = sum(getattr(w["host"], f).nbytes for f in capi.EPOCH_PTR_FIELDS)
        d2h = 0
        if world == 1:
            pool = b.build_pool(w["host"], reuse=True)
            n_ev = pool.stats["n_evals"]
            d2h = sum(getattr(pool, f).nbytes for f, _d, _k in capi.POOL_FIELDS)
        else:
            n_ev = b.build(w["host"])["n_evals"]
            if rank == 0:
                pool = peer["pp"].fetch(reuse=True)
                d2h = sum(getattr(pool, f).nbytes for f, _d, _k in capi.POOL_FIELDS)
        if "npo_host" in w and rank == 0:
            oi, _op, _od = b.npo_match(*w["npo_host"])
            h2d += sum(a.nbytes for a in w["npo_host"]); d2h += 16 * oi.size
        return n_ev, h2d, d2h

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
    tot_e = torch.tensor([float(e_evals), float(e_h2d), float(e_d2h)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_e, op=dist.ReduceOp.SUM)
    e_evals_all, e_h2d_all, e_d2h_all = (float(x) for x in tot_e.tolist())
    e2e_val = e_evals_all / float(t_e.item())
    # clocks: sampled across both timed regions (resident and end-to-end); a run shorter than the sampling period keeps stepping (every
    # rank together, untimed) until one reading has arrived
    if world > 1:
        while True:
            need = torch.tensor([1.0 if (sampler is not None and sampler.needs_sample()) else 0.0], device=dev)
            dist.broadcast(need, 0)
            if need.item() == 0.0:
                break
            step_resident(0)
        clocks = sampler.stop() if sampler else None
    else:
        clocks = sampler.stop(lambda: step_resident(0)) if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------
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
    gms = C.c_double(0.0)
    b._check(b.lib.amod_gather_bench(b.h, o_t.data_ptr(), d_t.data_ptr(), out_t.data_ptr(), ng, 5, C.byref(gms)))
    l2_gather_gbs = ng * 32 / (gms.value * 1e-3) / 1e9
    traffic_info = {}
    try:
        traffic_info = json.load(open(os.path.join(ROOT, "profiles", "eval_traffic.json")))
    except Exception:
        pass
    step_ms = ms / args.steps
    if sba and ms_screen > ms_eval:     # cfg5: the 10^8-pair reachability screen dominates (one lookup per pair)
        alg_bytes, k_ms, k_launches = screens_p * 32.0, ms_screen, 2 * prof_steps
        kname = "k_screen_count + k_screen_fill (reachability screen, dispatch_sba.py:59)"
        traffic = None
    else:
        alg_bytes, k_ms, k_launches = lookups_eval_p * 32.0, ms_eval, n_eval_launches
        kname = "k_eval (insertion search, one pass per trip level)"
        traffic = traffic_info.get(args.config, {}).get("dram_bytes_per_launch") if isinstance(traffic_info.get(args.config), dict) else None
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved,
                "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes / max(k_launches, 1), "launches": k_launches,
                "avg_launch_ms": k_ms / max(k_launches, 1), "kernel_share_of_step": (k_ms / prof_steps) / step_ms,
                "l2_gather_peak_gbs": l2_gather_gbs, "frac_of_l2_gather_peak": achieved / l2_gather_gbs,
                "measured_on": "rank 0" + (" (the gathering rank: it builds less than an even share of the fleet, so its launches are smaller than the other ranks')" if world > 1 else ""),
                "traffic_source": traffic_info.get(args.config, {}).get("source") if isinstance(traffic_info.get(args.config), dict) else None,
                "note": "algorithmic bytes = reference travel-time lookups inside the kernel's work x 32 B sector; the table is "
                        "L2-resident so HBM peak is the fallback bound and the measured L2 random-gather rate the real one"}

    fleet = work[0]["n_veh_full"]
    par = f"vehicles interleaved over {world} rank(s)"
    if world > 1:
        par += (", placement " + placement["how"] + "; pool gathered and merged on rank 0 over NVLink peer memory inside the build's own "
                "graph (bulk peer stores + epoch flags, one host sync per epoch, no host collective; NCCL only for start-up, "
                "barriers and the final timing reduction)")
    line = {"metric": METRIC, "value": evals_all / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["what"].format(fleet=fleet, n=len(work), req=work[0]["host"].n_search), "name": args.config,
                       "fleet": fleet, "fleet_mult": mult, "requests_considered": int(np.mean([w["host"].n_search for w in work])),
                       "parallelism": par,
                       "l2_policy": f"inputs differ every step ({len(work)} snapshots cycled); table deliberately L2-resident (67 MB persisting window)",
                       "pool_entries_per_step": entries_all / args.steps, "schedules_per_step": evals_all / args.steps,
                       "graph": not args.no_graph},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(e_h2d_all // args.steps), "d2h_bytes_per_step": int(e_d2h_all // args.steps),
                    "ms_per_step": 1e3 * float(t_e.item()) / args.steps,
                    "path": "amod_osp_build(host pointers) on every rank + " + ("peer assembly inside the build + amod_gpool_fetch of the assembled pool on rank 0"
                                                                                 if world > 1 else "amod_pool_fetch(host buffers)") + " through ctypes"},
            "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roofline, "parity": parity}
    if sba:
        line["config"].update({"sba_screens_per_step": screens_p / max(prof_steps, 1), "screen_ms_per_step": ms_screen / max(prof_steps, 1),
                               "eval_ms_per_step": ms_eval / max(prof_steps, 1), "npo_ms_per_step": ms_npo / max(prof_steps, 1),
                               "npo_pairs_per_step": npo_pairs / max(prof_steps, 1)})
    if not (args.no_cpu or world > 1):
        cpu = cpu_leg([w["host"] for w in work], tt)
        pyref = python_reference_leg() if args.config == "cfg2" else None
        if pyref:
            cpu["python_reference"] = pyref
        line["cpu_baseline"] = cpu
    if not (args.no_cpu or world > 1) and args.config == "cfg2":
        # beyond the metric: one whole simulated epoch with the fleet resident on the device, for information
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import prof_fleet
            line["fleet_epoch"] = prof_fleet.measure(veh=2000, cap=4, epochs=110, tail=40, device=local)
        except Exception as exc:      # never costs the bench line
            line["fleet_epoch"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    print_json(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
