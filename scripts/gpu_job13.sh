#!/bin/bash
mkdir -p gpurun_out
# CHOMP against the HBM-resident 400^3 field: DRAM traffic and hit rates of one launch
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:chomp_iterate -s 8 -c 2 --csv --log-file gpurun_out/r2_hbm_field_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --only hbm > gpurun_out/r2_hbm_field.log 2>&1
grep -v "^==" gpurun_out/r2_hbm_field_ncu.csv | cut -d, -f5,13,14,15 | tail -24
# SDF build kernels, full set
timeout 900 ncu --set full --clock-control none -k regex:"edt_|pack_rows" -s 6 -c 3 -o gpurun_out/r2_sdf -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --only cfg3 > gpurun_out/r2_sdf_ncu.log 2>&1
# tiled path kernels, full set (one cost launch + one update launch)
timeout 900 ncu --set full --clock-control none -k regex:"chomp_tile_cost|chomp_run_update" -s 4 -c 2 -o gpurun_out/r2_tiled -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --only cfg5 > gpurun_out/r2_tiled_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
echo done
