#!/bin/bash
# final records of the round: whole GPU suite, smoke, both bench arms as the driver runs them, then (after those exited 0 without ncu)
# the warm launch list and the steady-state metrics of the insertion-search kernels, and one full capture of k_eval
TAG=${1:-r02_v29}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench default rc=$?"
python - <<PY
import json
for f in ("bench", "bench_default", "bench_reference"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e", d["e2e"].get("ms_per_step"), "%.4g" % d["e2e"]["value"], "frac", (d.get("roofline") or {}).get("frac"), "parity", (d.get("parity") or {}).get("ok"), "clocks", d.get("clocks"), (d.get("fleet_epoch") or {}).get("ms"))
    except Exception as e:
        print(f, "no json", e)
PY
timeout 300 ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches_warm.csv python scripts/prof_eval.py --out gpurun_out/${TAG}_eval_work0.json > gpurun_out/${TAG}_ncu0.log 2>&1; echo "ncu launches rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches_warm.csv 3 > gpurun_out/${TAG}_launches_warm_summary.txt 2>&1; head -12 gpurun_out/${TAG}_launches_warm_summary.txt
M=lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 400 ncu --profile-from-start off --cache-control none --clock-control none -k regex:"k_eval|k_place|k_classify" --metrics $M --csv --log-file gpurun_out/${TAG}_k_eval_place_warm_ncu.csv python scripts/prof_eval.py --out gpurun_out/${TAG}_k_eval_warm_work.json > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu metrics rc=$?"
python scripts/eval_traffic_summary.py gpurun_out/${TAG}_k_eval_place_warm_ncu.csv gpurun_out/${TAG}_k_eval_warm_work.json > gpurun_out/${TAG}_k_eval_place_warm_ncu_summary.txt 2>&1; tail -3 gpurun_out/${TAG}_k_eval_place_warm_ncu_summary.txt
timeout 400 ncu --profile-from-start off --set full --cache-control none --clock-control none --import-source on -k regex:k_eval -s 14 -c 4 -f -o gpurun_out/${TAG}_k_eval_full python scripts/prof_eval.py --out gpurun_out/${TAG}_eval_work2.json > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu full rc=$?"
timeout 300 ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches_warm_cfg5.csv python scripts/prof_eval.py --config cfg5 --out gpurun_out/${TAG}_eval_work_cfg5.json > gpurun_out/${TAG}_ncu5.log 2>&1; echo "ncu cfg5 launches rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches_warm_cfg5.csv 0 > gpurun_out/${TAG}_launches_warm_cfg5_summary.txt 2>&1; head -24 gpurun_out/${TAG}_launches_warm_cfg5_summary.txt
