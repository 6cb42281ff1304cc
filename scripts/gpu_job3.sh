#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --timeout=1500 -p no:cacheprovider > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu_full.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -3 gpurun_out/r2_bench_b.err
echo done
