#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --timeout=1500 -p no:cacheprovider > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu_full.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
tail -2 gpurun_out/r2_bench_full.err
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
# fp64 op counts + DRAM traffic of the CHOMP launch
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:chomp_iterate -s 3 -c 1 --csv --log-file gpurun_out/r2_fp64_ops.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
# full-set capture of the CHOMP launch
timeout 900 ncu --set full --clock-control none -k regex:chomp_iterate_jit -s 3 -c 1 -o gpurun_out/r2_chomp_final -f python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_final.log 2>&1
echo done
