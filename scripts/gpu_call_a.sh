#!/bin/bash
# round-2 GPU call A: baseline tests + bench + L2-policy A/B + steady-state (no cache flush) ncu evidence
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
AMOD_NO_L2_POLICY=1 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02a_bench_nol2.json 2> gpurun_out/r02a_bench_nol2.err; echo "bench nol2 rc=$?"
M=lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
python scripts/prof_eval.py --out gpurun_out/r02a_eval_work.json > gpurun_out/r02a_prof_plain.log 2>&1 && \
ncu --profile-from-start off --cache-control none --clock-control none -k regex:k_eval --metrics $M --csv --log-file gpurun_out/r02a_eval_warm.csv python scripts/prof_eval.py --out gpurun_out/r02a_eval_work.json > gpurun_out/r02a_ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02a_launches_warm.csv python scripts/prof_eval.py --out gpurun_out/r02a_eval_work.json > gpurun_out/r02a_ncu2.log 2>&1
echo "ncu2 rc=$?"
AMOD_NO_L2_POLICY=1 ncu --profile-from-start off --cache-control none --clock-control none -k regex:k_eval --metrics $M --csv --log-file gpurun_out/r02a_eval_warm_nol2.csv python scripts/prof_eval.py --out gpurun_out/r02a_eval_work_nol2.json > gpurun_out/r02a_ncu3.log 2>&1
echo "ncu3 rc=$?"
tail -3 gpurun_out/r02a_pytest.log
