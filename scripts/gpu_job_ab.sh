#!/bin/bash
# A/B of a development knob on the headline: bench headline with and without, then the CHOMP suites
mkdir -p gpurun_out
KNOB=${1:-OCB_COST_PERM}
for v in 1 0 1 0; do
  env $KNOB=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('$KNOB=$v value', d['value'], 'kern_ms', d['kernel_ms_per_step'], 'failed', d['runs_failed_joint_limits'])"
done
timeout 1500 python -m pytest tests/test_gpu_chomp.py tests/test_gpu_fullsize.py -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -4
