#!/bin/bash
mkdir -p gpurun_out
for mb in 3 2 1; do
OCB_JIT_MINBLOCKS=$mb timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_e.json')); print('minblocks $mb value', d['value'], 'kern_ms', d['kernel_ms_per_step'])"
done
echo done
