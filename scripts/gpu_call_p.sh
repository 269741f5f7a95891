#!/bin/bash
# source-level ncu captures (warm graph replays, no cache flush): k_eval (2 launches), k_gen_pairs<4>, k_gen_emit, k_classify (1 each)
TAG=${1:-r03p}
mkdir -p gpurun_out
timeout 200 python scripts/prof_eval.py --out gpurun_out/${TAG}_eval_work.json > gpurun_out/${TAG}_prof_plain.log 2>&1 || exit 1
timeout 500 ncu --profile-from-start off --set full --cache-control none --clock-control none --import-source on -k regex:k_eval -s 14 -c 2 -f -o gpurun_out/${TAG}_k_eval python scripts/prof_eval.py --out gpurun_out/${TAG}_w1.json > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 500 ncu --profile-from-start off --set full --cache-control none --clock-control none --import-source on -k regex:"k_gen_pairs|k_gen_emit" -s 14 -c 4 -f -o gpurun_out/${TAG}_k_gen python scripts/prof_eval.py --out gpurun_out/${TAG}_w2.json > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/${TAG}_*.ncu-rep
