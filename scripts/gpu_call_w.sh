#!/bin/bash
# k_eval block shapes (library variants of the same tree) + the new tests
TAG=${1:-r03w2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fingerprint or wide or npo" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
for v in W8B4 W4B8 W4B9 W4B10 W2B18; do
  lib=/root/repo/amod_b200/csrc/libamod_b200_$v.so
  [ -f $lib ] || continue
  for rep in 1 2; do
    AMOD_B200_LIB=$lib timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu > gpurun_out/${TAG}_bench_${v}_$rep.json 2> gpurun_out/${TAG}_bench_${v}_$rep.err; echo "bench $v rc=$?"
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_W*_[12].json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "eval avg ms %.4f" % d["roofline"]["avg_launch_ms"], "parity", d["parity"]["ok"])
    except Exception as e:
        print(f, "no json", e)
PY
