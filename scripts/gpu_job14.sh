#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_sdf.py tests/test_gpu_fullsize.py -m gpu -q --timeout=1500 -p no:cacheprovider -k "sdf or cfg3 or mesh or flood or occup or distance" > gpurun_out/r2_pytest_sdf.log 2>&1
tail -5 gpurun_out/r2_pytest_sdf.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --only cfg3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c.json')); s=d['sdf_build']; print('sdf ms', s['ms'], s['value'], s['roofline']['frac'], 'e2e', s['e2e'])"
timeout 900 ncu --set full --clock-control none -k regex:"edt_|pack_rows" -s 6 -c 3 -o gpurun_out/r2_sdf_after -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --only cfg3 > gpurun_out/r2_sdf_ncu.log 2>&1
echo done
