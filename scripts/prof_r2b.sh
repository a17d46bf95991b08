# profiles/ recipe of round 2 (step kernel): launch list + full capture of the step kernel, after the plain run exited 0
set -x
B="python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --sustained-s 0 --ppo-envs-per-gpu 0"
$B > gpurun_out/r2f_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2f_launches.csv $B > gpurun_out/r2f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:brb_step -s 200 -c 1 -f -o gpurun_out/r2f_prof_step $B > gpurun_out/r2f_ncu2.log 2>&1
ncu -i gpurun_out/r2f_prof_step.ncu-rep --page raw --csv > gpurun_out/r2f_raw_step.csv 2>/dev/null
ncu -i gpurun_out/r2f_prof_step.ncu-rep --page source --csv > gpurun_out/r2f_src_step.csv 2>/dev/null
tail -2 gpurun_out/r2f_plain.log | cut -c1-300
