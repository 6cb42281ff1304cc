#!/bin/bash
# compute-sanitizer over the round-2 kernels that are new since the round-1 run (constraints, compiled-robot
# kernel, scan solve); writes gpurun_out/r2_sanitizer.txt
mkdir -p gpurun_out
out=gpurun_out/r2_sanitizer.txt
echo "# compute-sanitizer on a B200 (gpurun), round 2" > $out
run() {
  tool=$1; shift
  echo "compute-sanitizer --tool $tool python -m pytest $*" >> $out
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python -m pytest "$@" -m gpu -q -x --timeout=1400 -p no:cacheprovider > gpurun_out/san.log 2>&1
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/san.log | sed 's/^/   -> /' >> $out
  grep -E "Race reported|Invalid|hazard" gpurun_out/san.log | head -5 >> $out
}
run memcheck tests/test_gpu_constraints.py
run racecheck tests/test_gpu_constraints.py -k "tsr_constraint_matches or start_tsr or wider_metric or floating"
run memcheck tests/test_gpu_chomp.py -k "config1 or smallest or per_iteration"
run racecheck tests/test_gpu_chomp.py -k "config1"
cat $out
