#!/bin/bash
TAG=${1:-r03e}
mkdir -p gpurun_out
timeout 600 python scripts/run_day.py --out gpurun_out/${TAG}_day.json > gpurun_out/${TAG}_day.log 2>&1; echo "day rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_day.json')); print({k: d[k] for k in ('wall_s','ms_per_epoch_mean','ms_per_epoch_p99','ms_per_epoch_max','schedules_evaluated','report')}); print(d['heaviest_epochs'][:3])"
# launch list of 50 epochs (2,000 vehicles, 137 arrivals per epoch) of the device-resident loop (kernel names + durations)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_fleet_launches.csv python scripts/prof_fleet.py --epochs 50 --tail 10 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/${TAG}_fleet_launches.csv 2>/dev/null | head -40
