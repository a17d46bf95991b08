"""Turns the files scripts/prof_cmd.sh leaves in gpurun_out/ into the tracked summaries under profiles/."""
import collections, csv, pathlib, shutil, subprocess, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
shutil.copy(G / "launches_s3.csv", P / "r1_launches.csv")
shutil.copy(G / "raw_s3.csv", P / "r1_step_kernel_ncu_raw.csv")
shutil.copy(G / "raw_ppo.csv", P / "r1_ppo_kernels_ncu_raw.csv")
shutil.copy(G / "bench_r1_s3.json", P / "bench_r1_final.json")
# SASS with line info of the library that was profiled
sass = pathlib.Path("/tmp/sass"); sass.mkdir(exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "brb_kernels.sm_100a.cubin", str(ROOT / "balance_robot_b200/csrc/libbrb_cuda.so")], cwd=sass, check=True, capture_output=True)
(sass / "cur.txt").write_text(subprocess.run(["nvdisasm", "-g", "-c", "brb_kernels.sm_100a.cubin"], cwd=sass, check=True, capture_output=True, text=True).stdout)
rows = list(csv.reader(open(G / "src_s3.csv")))
hdr, data = rows[1], rows[2:]
ie, it, isrc = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
tot = sum(int(r[ie]) for r in data); tt = sum(int(r[it]) for r in data)
out = ["brb_step_kernel<Env01-v2>, 65,536 envs, steady state (step launch 200 after reset_all), from ncu --set full --import-source on (source page)",
       f"warp instructions executed: {tot}  ({tot / (2048 * 250):.0f} per warp per substep at 2048 warps x 250 substeps)",
       f"thread instructions executed: {tt}  -> {tt / tot:.2f} of 32 lanes active on average", "", f"{'opcode':10s}{'share %':>8s}{'lanes':>9s}"]
ops = collections.defaultdict(lambda: [0, 0])
for r in data:
    tok = r[isrc].split()
    op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
    ops[op][0] += int(r[ie]); ops[op][1] += int(r[it])
for op, (c, t) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:22]:
    out.append(f"{op:10s}{100 * c / tot:8.2f}{t / max(1, c):9.1f}")
out += ["", "by source function (scripts/attribute_sass.py: nvdisasm -g line info joined with the ncu source page)",
        subprocess.run([sys.executable, str(ROOT / "scripts/attribute_sass.py"), str(sass / "cur.txt"), str(G / "src_s3.csv"), "brb_step_kernelILi1E"],
                       capture_output=True, text=True).stdout]
(P / "r1_step_kernel_opcode_mix.txt").write_text("\n".join(out))
rows = list(csv.reader(open(G / "raw_s3.csv"))); h, v = rows[0], rows[2]
for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_active.avg", "sm__cycles_elapsed.avg"):
    print(k, rows[1][h.index(k)], v[h.index(k)])
for k, x in zip(h, v):
    if "stalled" in k and "per_issue_active" in k and float(x or 0) > 0.08: print(k, x)
