"""profiles/<name>_ncu_raw.csv (ncu -i x.ncu-rep --page raw --csv, one kernel row) -> profiles/<name>_ncu_summary.json:
the handful of counters bench.py's `roofline.executed` is computed from, so that a reader can recompute every roofline
number from profiles/ plus the bench line.   usage: python scripts/ncu_summary.py raw.csv out.json robots [row]"""
import csv, json, sys

raw, out, robots = sys.argv[1], sys.argv[2], int(sys.argv[3])
row = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rows = list(csv.reader(open(raw)))
h, units, v = rows[0], rows[1], rows[2 + row]
def f(k):
    x = float(v[h.index(k)].replace(",", ""))
    u = units[h.index(k)]
    return x * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(u, 1)
cyc = f("smsp__cycles_elapsed.avg") if "smsp__cycles_elapsed.avg" in h else f("sm__cycles_elapsed.avg")
d = {"kernel": v[h.index("Kernel Name")], "robots": robots, "duration_s_under_ncu": f("gpu__time_duration.sum"), "cycles_elapsed": cyc,
     "thread_inst_ffma": f("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed") * cyc,
     "thread_inst_fmul": f("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed") * cyc,
     "thread_inst_fadd": f("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed") * cyc,
     "warp_inst": f("smsp__inst_executed.sum"), "lanes_active": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
     "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "pipe_fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
     "pipe_alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
     "pipe_fp64_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
     "registers": f("launch__registers_per_thread"), "grid": f("launch__grid_size"),
     "dram_bytes": f("dram__bytes_read.sum") + f("dram__bytes_write.sum"),
     "sm_cycles_active_over_elapsed": f("sm__cycles_active.avg") / f("sm__cycles_elapsed.avg")}
d["executed_flop_per_robot_step"] = (2 * d["thread_inst_ffma"] + d["thread_inst_fmul"] + d["thread_inst_fadd"]) / robots
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d, indent=1))
