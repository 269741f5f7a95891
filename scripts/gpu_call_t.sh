#!/bin/bash
TAG=${1:-r02t}
mkdir -p gpurun_out
timeout 200 python scripts/prof_e2e.py > gpurun_out/${TAG}_prof_e2e.txt 2>&1; cat gpurun_out/${TAG}_prof_e2e.txt | tail -2
M=lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 200 python scripts/prof_eval.py --out gpurun_out/${TAG}_eval_work.json > gpurun_out/${TAG}_prof_plain.log 2>&1 && \
timeout 400 ncu --profile-from-start off --cache-control none --clock-control none -k regex:"k_eval|k_place|k_classify" --metrics $M --csv --log-file gpurun_out/${TAG}_eval_warm.csv python scripts/prof_eval.py --out gpurun_out/${TAG}_eval_work.json > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 400 ncu --profile-from-start off --set full --cache-control none --clock-control none --import-source on -k regex:k_eval -s 14 -c 4 -o gpurun_out/${TAG}_k_eval_full python scripts/prof_eval.py --out gpurun_out/${TAG}_eval_work2.json > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu2 rc=$?"
ls -la gpurun_out/${TAG}_k_eval_full.ncu-rep
