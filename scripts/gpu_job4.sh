#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_chomp.py tests/test_gpu_fullsize.py -m gpu -q --timeout=1500 -p no:cacheprovider -k "rest or cfg2 or joint_limit or multi or split or two_engines" > gpurun_out/r2_pytest_sel.log 2>&1
tail -5 gpurun_out/r2_pytest_sel.log
OCB_JIT_FLAGS="-lineinfo" timeout 900 ncu --set full --import-source on --clock-control none -k regex:chomp_iterate_jit -s 3 -c 1 -o gpurun_out/r2_chomp_robot -f python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_robot.log 2>&1
tail -3 gpurun_out/r2_ncu_robot.log
echo done
