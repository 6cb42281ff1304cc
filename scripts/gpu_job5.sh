#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --only cfg3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c.json')); print('value', d['value'], 'kern_ms', d['kernel_ms_per_step'], 'sdf ms', d['sdf_build']['ms'], d['sdf_build']['roofline']['frac'])"
timeout 600 python scripts/dev_phase_clocks.py > gpurun_out/r2_phase_clocks.txt 2>&1
cat gpurun_out/r2_phase_clocks.txt | tail -20
timeout 2400 python -m pytest tests -m gpu -q --timeout=1500 -p no:cacheprovider -x > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu_full.log
echo done
