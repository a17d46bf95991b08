set -x
B="python bench.py --env Env03-v2 --steps 5 --warmup 3 --spinup 40 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_e3.log 2>&1 || { tail -5 gpurun_out/plain_e3.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:brb_step -s 44 -c 1 -f -o gpurun_out/prof_e3 $B > gpurun_out/ncu_e3.log 2>&1
ncu -i gpurun_out/prof_e3.ncu-rep --page raw --csv > gpurun_out/raw_e3.csv 2>/dev/null
ncu -i gpurun_out/prof_e3.ncu-rep --page source --csv > gpurun_out/src_e3.csv 2>/dev/null
tail -1 gpurun_out/plain_e3.log | cut -c1-300
