# profiles/ recipe of round 2: gpu tests, bench line, launch list, full capture of the step kernel
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; tail -4 gpurun_out/r2_gputests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err || { tail -20 gpurun_out/r2_bench.err; exit 1; }
B="python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --sustained-s 0 --ppo-envs-per-gpu 0"
$B > gpurun_out/r2_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:brb_step -s 200 -c 1 -f -o gpurun_out/r2_prof_step $B > gpurun_out/r2_ncu2.log 2>&1
ncu -i gpurun_out/r2_prof_step.ncu-rep --page raw --csv > gpurun_out/r2_raw_step.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_step.ncu-rep --page source --csv > gpurun_out/r2_src_step.csv 2>/dev/null
cat gpurun_out/r2_bench.json
