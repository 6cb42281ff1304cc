#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, fp64 op counts + DRAM traffic and a
# full-set capture of the CHOMP launch (never a bench value: the numbers printed under ncu are discarded)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:chomp_iterate -s 3 -c 1 --csv --log-file gpurun_out/r2_fp64_ops.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:chomp_iterate_jit -s 3 -c 1 -o gpurun_out/r2_chomp_final -f python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_final.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_headline.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_headline.json')); print('value', d['value'], 'kern_ms', d['kernel_ms_per_step'], 'run_iters', d['run_iterations_per_step'])"
echo done
