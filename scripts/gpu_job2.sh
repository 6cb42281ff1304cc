#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sdf > gpurun_out/r2_bench_robot.json 2> gpurun_out/r2_bench_robot.err
OCB_JIT_ROBOT=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sdf > gpurun_out/r2_bench_generic.json 2> gpurun_out/r2_bench_generic.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sdf --no-jit > gpurun_out/r2_bench_lib.json 2> gpurun_out/r2_bench_lib.err
timeout 2400 python -m pytest tests -m gpu -q --timeout=1200 2>&1 | tail -60 > gpurun_out/r2_pytest_gpu.log
echo done
