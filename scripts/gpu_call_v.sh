#!/bin/bash
# verification of the tree as it stands: whole GPU suite, smoke, the bench as the driver runs it (+ reference arm), cfg5
TAG=${1:-r03v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --config cfg5 --no-cpu > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err; echo "bench cfg5 rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --config cfg3 --no-cpu > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; echo "bench cfg3 rc=$?"
python - <<PY
import json
for f in ("bench", "bench_reference", "bench_cfg5", "bench_cfg3"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e", d["e2e"].get("ms_per_step"), "%.4g" % d["e2e"]["value"], "frac", (d.get("roofline") or {}).get("frac"), "parity", (d.get("parity") or {}).get("ok"), {k: round(v, 3) for k, v in d["config"].items() if k.endswith("_ms_per_step")}, (d.get("fleet_epoch") or {}).get("ms"))
    except Exception as e:
        print(f, "no json", e)
PY
