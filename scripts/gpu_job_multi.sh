#!/bin/bash
# multi-GPU measurements: weak + strong scaling lines at N = 8, 4, 2 and the one-process multi-engine test
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
for N in 8 4 2; do
  if [ $N -le $NG ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
    tail -1 gpurun_out/r2_bench_${N}gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d.get('strong_scaling') or {}; print('N=$N weak', d['value'], d['ms_per_step'], 'kern', d['kernel_ms_per_step'], 'strong', s.get('value'), s.get('ms_per_step'), s.get('kernel_ms_per_step'))"
  fi
done
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
timeout 900 python -m pytest tests/test_gpu_chomp.py -m gpu -q --timeout=600 -p no:cacheprovider -k "multi_engine or two_engines" > gpurun_out/r2_pytest_multi.log 2>&1
tail -5 gpurun_out/r2_pytest_multi.log
echo done
