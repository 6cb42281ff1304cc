"""Print a compact per-kernel summary of an .ncu-rep (raw page): duration, DRAM traffic, occupancy, pipes."""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    if len(r) < len(hdr): continue
    print("== %s" % r[ki][:90])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("   %-66s %-14s %s" % (k, units[i], r[i]))
