#!/bin/bash
# tests + bench + warm launch list (tag = $1)
TAG=${1:-r02b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json | cut -c1-600
python scripts/prof_eval.py --out gpurun_out/${TAG}_eval_work.json > gpurun_out/${TAG}_prof_plain.log 2>&1 && \
ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches_warm.csv python scripts/prof_eval.py --out gpurun_out/${TAG}_eval_work.json > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
