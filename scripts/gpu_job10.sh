#!/bin/bash
mkdir -p gpurun_out
OCB_JIT_FLAGS="" timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_e.json')); print('value', d['value'], 'kern_ms', d['kernel_ms_per_step'], d['config']['kernel'], 'failed', d['runs_failed_joint_limits'])"
timeout 600 python scripts/dev_phase_clocks.py > gpurun_out/r2_phase_clocks.txt 2>&1
cat gpurun_out/r2_phase_clocks.txt | tail -18
timeout 2400 python -m pytest tests/test_gpu_chomp.py -m gpu -q --timeout=1500 -p no:cacheprovider > gpurun_out/r2_pytest_chomp.log 2>&1
tail -8 gpurun_out/r2_pytest_chomp.log
echo done
