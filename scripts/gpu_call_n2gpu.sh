#!/bin/bash
# 2 GPUs: peer tests, then the bench as the driver launches it at N=2 (cfg2 weak) and cfg5 at N=2
TAG=${1:-r03g2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/${TAG}_peer.log 2>&1; echo "peer pytest rc=$?"; tail -3 gpurun_out/${TAG}_peer.log
bash scripts/gpu_call_scale.sh 2 ${TAG}_bench_n2 --steps 20 --warmup 5
bash scripts/gpu_call_scale.sh 2 ${TAG}_bench_cfg5_n2 --steps 10 --warmup 3 --config cfg5 --no-cpu
