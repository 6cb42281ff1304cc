#!/bin/bash
# ncu counters of the constrained launch (library kernel with the TSR projection): profiles/r2_tsr_ncu.csv
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:chomp_iterate_kernel -s 2 -c 1 --csv --log-file gpurun_out/r2_tsr_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --only tsr > /dev/null 2>&1
grep -c . gpurun_out/r2_tsr_ncu.csv
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_tsr_ncu.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][-60:], r[-3], r[-2], r[-1])
PY
