#!/bin/bash
mkdir -p gpurun_out
for flags in "" "-DJR_NO_TRIG_CACHE"; do
OCB_JIT_FLAGS="$flags" timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_e.json')); print('flags [$flags] value', d['value'], 'kern_ms', d['kernel_ms_per_step'], 'failed', d['runs_failed_joint_limits'])"
done
timeout 1200 python -m pytest tests/test_gpu_chomp.py -m gpu -q --timeout=600 -p no:cacheprovider -x 2>&1 | tail -3
