#!/bin/bash
# round-end style validation: smoke, full GPU suite, reference arm, default bench line
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 3000 python -m pytest tests -m gpu -q --timeout=1500 -p no:cacheprovider > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_full.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> /dev/null
t0=$(date +%s); timeout 1200 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo "default bench.py took $(( $(date +%s) - t0 )) s" 
python -c "
import json
r=json.load(open('gpurun_out/r2_bench_reference_arm.json')); d=json.load(open('gpurun_out/r2_bench_full.json'))
print('reference arm', r['value'], r['cpu_baseline']['cores'], 'ours', d['value'], 'e2e', d['e2e']['value'], 'ratio e2e', d['e2e']['value']/r['value'])
print('fp64', d['roofline_fp64']); print('hbm', d['roofline']['frac'])
for k,v in d['configs'].items(): print(k, v['value'], (v.get('roofline') or {}).get('frac'), v['e2e']['value'], v['cpu_baseline']['value'])
"
echo done
