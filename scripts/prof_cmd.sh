# profiles/ recipe of round 1: bench line, launch list, full capture of the step kernel, full capture of the PPO kernels
set -x
python bench.py > gpurun_out/bench_r1_s3.json 2> gpurun_out/bench_r1_s3.err || exit 1
B="python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_s3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_s3.csv $B > gpurun_out/ncu1_s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:brb_step -s 200 -c 1 -f -o gpurun_out/prof_s3 $B > gpurun_out/ncu2_s3.log 2>&1
ncu -i gpurun_out/prof_s3.ncu-rep --page raw --csv > gpurun_out/raw_s3.csv 2>/dev/null
ncu -i gpurun_out/prof_s3.ncu-rep --page source --csv > gpurun_out/src_s3.csv 2>/dev/null
P="python scripts/ppo_prof.py"
$P > gpurun_out/plain_ppo.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"brb_ppo_grad|brb_policy_act" -s 8 -c 2 -f -o gpurun_out/prof_ppo $P > gpurun_out/ncu_ppo.log 2>&1
ncu -i gpurun_out/prof_ppo.ncu-rep --page raw --csv > gpurun_out/raw_ppo.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
cat gpurun_out/bench_r1_s3.json
