#!/bin/bash
TAG=${1:-r03m}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "npo or cfg5 or peak or fleet or trace" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for rep in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --config cfg5 --no-cpu > gpurun_out/${TAG}_bench_cfg5_$rep.json 2> gpurun_out/${TAG}_bench_cfg5_$rep.err; echo "bench cfg5 rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_cfg5_$rep.json").read().strip().splitlines()[-1])
print("ms/step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], {k: d["config"][k] for k in ("screen_ms_per_step", "eval_ms_per_step", "npo_ms_per_step") if k in d["config"]}, d["parity"]["ok"])
PY
done
cat > /tmp/npo_once.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "/root/repo")
from amod_b200 import synth
from amod_b200.pool import PoolBuilder
z = np.load("/root/repo/bench_data/cfg5_epoch0025.npz")
tt = synth.make_road_graph(seed=0).tt
b = PoolBuilder(tt, device=0)
oi, op, od = b.npo_match(z["npo_idle_nid"], z["npo_pend_onid"])
print(oi.size, b.last_npo_rounds)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_npo_launches.csv -k regex:k_npo python /tmp/npo_once.py > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_npo_launches.csv 0 2>/dev/null | head -8
tail -2 gpurun_out/${TAG}_ncu.log
