#!/bin/bash
# N GPUs ($1): the bench as the driver launches it (cfg2 weak), cfg3 (cap 8) and cfg4 (8,000 vehicles, strong)
N=${1:-8}; TAG=${2:-r03g8}
mkdir -p gpurun_out
bash scripts/gpu_call_scale.sh $N ${TAG}_bench_n$N --steps 20 --warmup 5
bash scripts/gpu_call_scale.sh $N ${TAG}_bench_cfg3_n$N --steps 20 --warmup 5 --config cfg3 --no-cpu
bash scripts/gpu_call_scale.sh $N ${TAG}_bench_cfg4_n$N --steps 20 --warmup 5 --config cfg4 --no-cpu
