#!/bin/bash
# full GPU test batch + benches for every single-GPU config (tag $1)
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
for CFG in cfg2 cfg3 cfg5; do
  timeout 600 python bench.py --steps 20 --warmup 5 --config $CFG > gpurun_out/${TAG}_bench_${CFG}.json 2> gpurun_out/${TAG}_bench_${CFG}.err; echo "bench $CFG rc=$?"
done
for M in 4 8 16; do
  timeout 600 python bench.py --steps 10 --warmup 3 --fleet-mult $M --no-cpu > gpurun_out/${TAG}_bench_cfg2_x${M}.json 2> gpurun_out/${TAG}_bench_cfg2_x${M}.err; echo "bench x$M rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms/step %.3f" % d["ms_per_step"], "value %.3g" % d["value"], "e2e ms %.3f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "parity", d["parity"]["ok"], d["parity"]["snapshots"], d.get("cpu_baseline", {}).get("value"), (d.get("cpu_baseline", {}).get("python_reference") or {}).get("value"))
    except Exception as e:
        print(f, "no json", e)
PY
