#!/bin/bash
TAG=${1:-r03d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fleet_day.py -x -q --durations=3 > gpurun_out/${TAG}_pytest_day.log 2>&1
echo "pytest rc=$?"
tail -n 40 gpurun_out/${TAG}_pytest_day.log
