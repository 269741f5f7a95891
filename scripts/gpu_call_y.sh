#!/bin/bash
# full GPU suite on the tree's library, then A/B of library variants (AMOD_B200_LIB), then the warm launch list
TAG=${1:-r03y}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for v in ${VARIANTS:-S E64 E80}; do
  lib=/root/repo/amod_b200/csrc/libamod_b200_$v.so
  [ -f $lib ] || continue
  for rep in 1 2; do
    AMOD_B200_LIB=$lib timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu > gpurun_out/${TAG}_bench_${v}_$rep.json 2> gpurun_out/${TAG}_bench_${v}_$rep.err; echo "bench $v rc=$?"
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "eval avg ms %.4f" % d["roofline"]["avg_launch_ms"], "parity", d["parity"]["ok"])
    except Exception as e:
        print(f, "no json", e)
PY
timeout 400 ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches_warm.csv python scripts/prof_eval.py --out gpurun_out/${TAG}_eval_work.json > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_launches_warm.csv 3 > gpurun_out/${TAG}_launches_warm_summary.txt 2>&1; head -24 gpurun_out/${TAG}_launches_warm_summary.txt
