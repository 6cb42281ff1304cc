#!/bin/bash
mkdir -p gpurun_out
for R in 4096 512; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --runs $R > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_e.json')); print('runs $R value', d['value'], 'kern_ms', d['kernel_ms_per_step'], d['config']['kernel'], 'failed', d['runs_failed_joint_limits'])"
done
timeout 2400 python -m pytest tests/test_gpu_chomp.py tests/test_gpu_fullsize.py -m gpu -q --timeout=1500 -p no:cacheprovider > gpurun_out/r2_pytest_chomp.log 2>&1
tail -6 gpurun_out/r2_pytest_chomp.log
echo done
