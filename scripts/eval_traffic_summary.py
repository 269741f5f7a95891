#!/usr/bin/env python
"""Turn one steady-state ncu capture of the insertion-search kernels into the figures bench.py reports as roofline.traffic.

    ncu --profile-from-start off --cache-control none --clock-control none -k regex:'k_eval|k_place|k_classify' --metrics <M> --csv \
        --log-file X.csv python scripts/prof_eval.py --out X_work.json          (scripts/gpu_call_t.sh)
    python scripts/eval_traffic_summary.py X.csv X_work.json [--write profiles/eval_traffic.json --tag r02_vNN]

Prints the per-kernel summary (profiles/r02_*_k_eval_place_warm_ncu_summary.txt) and, with --write, stores the per-launch DRAM bytes
and the L2 sectors per algorithmic lookup of k_eval for the config that was profiled."""
import argparse
import collections
import csv
import json
import statistics


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv"); ap.add_argument("work")
    ap.add_argument("--write", default=None); ap.add_argument("--tag", default="")
    ap.add_argument("--hbm-peak", type=float, default=6545.0)
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr, per = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            unit = d["Metric Unit"]
            if unit in ("ns", "nsecond"): v /= 1000.0
            elif unit in ("Kbyte",): v *= 1e3
            elif unit in ("Mbyte",): v *= 1e6
            elif unit in ("Gbyte",): v *= 1e9
            per.setdefault(d["ID"], {"name": d["Kernel Name"]})[d["Metric Name"]] = v
    work = json.load(open(a.work))
    steps = work["steps"]
    evals = sum(s["n_evals"] for s in steps); lookups = sum(s["n_lookups_eval"] for s in steps)
    kern = collections.OrderedDict()
    for k in per.values():
        n = k["name"]; n = n[:n.index("(")] if "(" in n else n
        kern.setdefault(n.replace("void ", ""), []).append(k)
    print(f"{len(per)} launches over {len(steps)} steps; algorithmic work of exactly these launches: {evals} evaluations, {lookups} reference lookups = {lookups * 32 / 1e6:.1f} MB of sectors")
    out = {}
    for n, ks in kern.items():
        t = sum(k.get("gpu__time_duration.sum", 0) for k in ks)
        lts = sum(k.get("lts__t_sectors.sum", 0) for k in ks)
        dr = sum(k.get("dram__bytes_read.sum", 0) for k in ks); dw = sum(k.get("dram__bytes_write.sum", 0) for k in ks)
        hit = statistics.median(k.get("lts__t_sector_hit_rate.pct", 0) for k in ks)
        l1s = sum(k.get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", 0) for k in ks); l1r = sum(k.get("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", 0) for k in ks)
        thr = statistics.mean(k.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0) for k in ks)
        wa = statistics.mean(k.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0) for k in ks)
        ia = statistics.mean(k.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0) for k in ks)
        print(f"{n:<24} n={len(ks):3d} time {t:8.1f} us ({t / len(steps):6.1f} us/step)  lts sectors {lts / 1e6:7.2f} M  L2 hit (median per launch) {hit:5.1f} %  "
              f"dram R {dr / 1e6:7.1f} MB W {dw / 1e6:6.1f} MB  l1 ld sectors/request {l1s / max(l1r, 1):4.2f}  threads/inst {thr:4.1f}  warps active {wa:4.1f} %  issue active {ia:4.1f} %")
        if n.startswith("k_eval"):
            out = {"dram_bytes_per_launch": (dr + dw) / len(ks), "algorithmic_bytes_per_launch": lookups * 32 / len(ks),
                   "lts_sectors_per_algorithmic_lookup": lts / max(lookups, 1), "time_us": t, "launches": len(ks)}
    if out:
        rate = lookups * 32 / (out["time_us"] * 1e-6) / 1e9
        print(f"k_eval: {out['lts_sectors_per_algorithmic_lookup']:.3f} L2 sectors per algorithmic lookup; DRAM bytes / algorithmic bytes = "
              f"{out['dram_bytes_per_launch'] / out['algorithmic_bytes_per_launch']:.3f}; algorithmic sector rate {rate:.0f} GB/s = {rate / a.hbm_peak:.3f} of the measured HBM peak ({a.hbm_peak:,.0f} GB/s)")
        print(f"k_eval dram bytes per launch (mean over {out['launches']} launches): {out['dram_bytes_per_launch']:.0f}; algorithmic bytes per launch: {out['algorithmic_bytes_per_launch']:.0f}")
        if a.write:
            try:
                cur = json.load(open(a.write))
            except Exception:
                cur = {}
            cur[work.get("config", "cfg2")] = {
                "dram_bytes_per_launch": out["dram_bytes_per_launch"], "algorithmic_bytes_per_launch": out["algorithmic_bytes_per_launch"],
                "lts_sectors_per_algorithmic_lookup": out["lts_sectors_per_algorithmic_lookup"],
                "source": f"profiles/{a.tag}_k_eval_place_warm_ncu.csv + {a.tag}_k_eval_warm_work.json: ncu --cache-control none on warm graph replays, dram__bytes_read.sum + "
                          f"dram__bytes_write.sum of the {out['launches']} k_eval launches of one pass over the {len(steps)} {work.get('config', 'cfg2')} snapshots, per launch (scripts/eval_traffic_summary.py)"}
            json.dump(cur, open(a.write, "w"), indent=1)


if __name__ == "__main__":
    main()
