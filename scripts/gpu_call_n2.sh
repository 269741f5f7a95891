#!/bin/bash
# NPO rewrite: the whole GPU suite, then the cfg5 bench (SBA + NPO at the peak epoch) and the default bench
TAG=${1:-r03n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --config cfg5 --no-cpu > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err; echo "bench cfg5 rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_cfg5.json").read().strip().splitlines()[-1])
print("ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], {k: d["config"][k] for k in ("screen_ms_per_step", "eval_ms_per_step", "npo_ms_per_step") if k in d["config"]}, d["parity"], d["roofline"]["kernel"], d["roofline"]["frac"])
PY
for v in ${VARIANTS:-N F}; do
  lib=/root/repo/amod_b200/csrc/libamod_b200_$v.so
  [ -f $lib ] || continue
  for rep in 1 2; do
    AMOD_B200_LIB=$lib timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu > gpurun_out/${TAG}_bench_${v}_$rep.json 2> gpurun_out/${TAG}_bench_${v}_$rep.err; echo "bench $v rc=$?"
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_[A-Z]*_[12].json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "eval avg ms %.4f" % d["roofline"]["avg_launch_ms"], "parity", d["parity"]["ok"])
    except Exception as e:
        print(f, "no json", e)
PY
