#!/bin/bash
# A/B of a development knob on the headline: bench headline with and without, then the CHOMP suites
mkdir -p gpurun_out
for v in 1 0; do
  OCB_LINE_FORM=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_$v.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('OCB_LINE_FORM=$v value', d['value'], 'kern_ms', d['kernel_ms_per_step'], 'failed', d['runs_failed_joint_limits'])"
done
timeout 1500 python -m pytest tests/test_gpu_chomp.py tests/test_gpu_fullsize.py tests/test_gpu_constraints.py tests/test_gpu_module.py -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -5
python -c "
import json
for k in ('jit','library'):
    c=json.load(open('gpurun_out/cfg2_census_%s.json'%k)); print(k, c['within_1e6'], c['over_1e6'], c['gpu_fail_ref_ok'], c['ref_fail_gpu_ok'], c['one_iteration_from_reference_state'])
"
