#!/bin/bash
TAG=${1:-r03c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --config cfg5 --no-cpu > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err; echo "bench cfg5 rc=$?"
timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --config cfg3 --no-cpu > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; echo "bench cfg3 rc=$?"
python - <<PY
import json
for f in ("bench", "bench_cfg5", "bench_cfg3"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f" % d["ms_per_step"], "e2e", d["e2e"].get("ms_per_step"), "frac", (d.get("roofline") or {}).get("frac"), "parity", (d.get("parity") or {}).get("ok"), {k: round(v, 3) for k, v in d["config"].items() if k.endswith("_ms_per_step")})
    except Exception as e:
        print(f, "no json", e)
PY
