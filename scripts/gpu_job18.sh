#!/bin/bash
mkdir -p gpurun_out
for tw in 32 16 8; do
OCB_TILE_W=$tw timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --only cfg5 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_e.json'))['configs']['cfg5_dense']; print('tile $tw value', d['value'], 'kern_ms', d['kernel_ms_per_step'], d['tile_width'])"
done
echo done
