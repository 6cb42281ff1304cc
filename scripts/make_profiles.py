"""Turn the raw ncu output of scripts/gpu_job_profiles.sh (gpurun_out/) into the tracked summaries under
profiles/: r2_fp64_ops.json, r2_launch_shares.txt, r2_launches.csv, r2_chomp_final_ncu.txt and the CHOMP
entries of traffic.json."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rows = [r for r in csv.reader(open(os.path.join(G, "r2_fp64_ops.csv"))) if len(r) > 10 and r[0].isdigit()]
m = {r[-3]: float(r[-1].replace(",", "")) for r in rows}
head = json.load(open(os.path.join(G, "r2_bench_headline.json")))
ri = head["run_iterations_per_step"]
dfma, dmul, dadd = (m["smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % k] for k in ("dfma", "dmul", "dadd"))
ops = {"kernel": "chomp_iterate_jit (run-time specialised, WAM7 compiled as code)",
       "launch": "bench step: 4096 runs x 100 iterations + final cost pass", "run_iterations": ri,
       "dfma": dfma, "dmul": dmul, "dadd": dadd, "warp_instructions": m["smsp__inst_executed.sum"],
       "threads_per_warp_instruction": m["smsp__thread_inst_executed_per_inst_executed.ratio"],
       "fp64_pipe_pct": m["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"],
       "issue_active_pct": m["smsp__issue_active.avg.pct_of_peak_sustained_active"],
       "gpu_time_ns": m["gpu__time_duration.sum"], "dram_bytes": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
       "flop_per_run_iteration": (2 * dfma + dmul + dadd) / ri,
       "source": "ncu --metrics smsp__sass_thread_inst_executed_op_d{fma,mul,add}_pred_on.sum ... --clock-control none "
                 "-k regex:chomp_iterate -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras"}
json.dump(ops, open(os.path.join(P, "r2_fp64_ops.json"), "w"), indent=1)
t = json.load(open(os.path.join(P, "traffic.json")))
t["chomp_iterate_kernel_bytes_per_launch"] = int(ops["dram_bytes"])
t["algorithmic_bytes_per_launch"] = int(58128 * ri)
json.dump(t, open(os.path.join(P, "traffic.json"), "w"), indent=1)
shutil.copy(os.path.join(G, "r2_launches.csv"), os.path.join(P, "r2_launches.csv"))
rows = [r for r in csv.reader(open(os.path.join(G, "r2_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4], [0, 0.0]); a[0] += 1; a[1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, "r2_launch_shares.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400, python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n"
            "# (set-up SDF build; headline: 3 warm-up + 2 timed + 3 end-to-end steps of chomp_iterate_jit; sub-records cfg3 / cfg4 / cfg5 / HBM field / TSR-constrained batch -- the latter runs the library kernel chomp_iterate_kernel<128,0,0,0,1>)\n"
            "# per-launch times are cold-cache and serialised: shares, not absolutes\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-64s launches %4d  total %10.3f ms  avg %8.3f ms  share %5.1f%%\n" % (k[-64:], v[0], v[1] / 1e6, v[1] / 1e6 / v[0], 100 * v[1] / tot))
rep = os.path.join(G, "r2_chomp_final.ncu-rep")
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
iS, iI = hdr.index("Source"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
opc, st, n_tot = collections.Counter(), collections.Counter(), 0
for r in rows[2:]:
    try: n = int(r[iI])
    except Exception: continue
    toks = r[iS].strip().split()
    if toks and toks[0].startswith("@"): toks = toks[1:]
    opc[(toks[0] if toks else "?").split(".")[0]] += n; n_tot += n
    for i in stall:
        try: st[hdr[i]] += int(r[i])
        except Exception: pass
T = sum(st.values())
with open(os.path.join(P, "r2_chomp_final_ncu.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none -k regex:chomp_iterate_jit -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline\n")
    f.write(out)
    f.write("\n# SASS: %d static instructions; executed warp-instruction mix:\n   " % (len(rows) - 2))
    f.write("  ".join("%s %.1f%%" % (k, 100 * v / n_tot) for k, v in opc.most_common(16)))
    f.write("\n# warp stall samples (all): " + "  ".join("%s %.1f%%" % (k[6:], 100 * v / T) for k, v in st.most_common(9)) + "\n")
print(open(os.path.join(P, "r2_chomp_final_ncu.txt")).read())
print(json.dumps({k: ops[k] for k in ("flop_per_run_iteration", "fp64_pipe_pct", "issue_active_pct", "warp_instructions", "dram_bytes")}))
