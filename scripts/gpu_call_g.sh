#!/bin/bash
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -k "npo or cfg5 or trace or peak" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 240 python bench.py --steps 20 --warmup 5 --config cfg5 > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err; echo "bench cfg5 rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_cfg5.json").read().strip().splitlines()[-1])
print("ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], {k: d["config"][k] for k in ("screen_ms_per_step", "eval_ms_per_step", "npo_ms_per_step")}, d["parity"], d["roofline"]["kernel"], d["roofline"]["frac"])
PY
