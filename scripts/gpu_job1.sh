#!/bin/bash
# round-2 GPU job: full GPU suite, bench line, ncu fp64 op counts of the CHOMP kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
nproc >> gpurun_out/r2_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -40 > gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:chomp_iterate -c 2 --csv --log-file gpurun_out/r2_fp64_ops.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sdf > gpurun_out/r2_bench_ncu.log 2>&1
echo done
