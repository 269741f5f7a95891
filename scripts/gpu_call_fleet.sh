#!/bin/bash
# fleet (8f-3) tests on the GPU box; output tail kept under gpurun_out/
TAG=${1:-r03f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fleet.py -x -q > gpurun_out/${TAG}_pytest_fleet.log 2>&1
echo "pytest rc=$?"
tail -n 40 gpurun_out/${TAG}_pytest_fleet.log
