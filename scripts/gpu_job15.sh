#!/bin/bash
mkdir -p gpurun_out
for cfg in "4096 0.05" "512 0.05" "512 0.3" "4096 0.3" "444 0.3" "148 0.3"; do
set -- $cfg
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --runs $1 --shrink $2 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_e.json')); print('runs $1 shrink $2 value', d['value'], 'kern_ms', d['kernel_ms_per_step'], 'failed', d['runs_failed_joint_limits'])"
done
echo done
