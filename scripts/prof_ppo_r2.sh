# round 2: gpu test suite, then ncu captures of the tcgen05 PPO gradient kernels inside one small PPO iteration (scripts/ppo_prof.py)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; tail -4 gpurun_out/r2_gputests.log
P="python scripts/ppo_prof.py"
$P > gpurun_out/r2_plain_ppo.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_ppo.csv $P > gpurun_out/r2_ncu_ppo1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"brb_ppo_grad_tc|brb_adam" -s 6 -c 3 -f -o gpurun_out/r2_prof_ppo $P > gpurun_out/r2_ncu_ppo2.log 2>&1
ncu -i gpurun_out/r2_prof_ppo.ncu-rep --page raw --csv > gpurun_out/r2_raw_ppo.csv 2>/dev/null
ncu -i gpurun_out/r2_prof_ppo.ncu-rep --page source --csv > gpurun_out/r2_src_ppo.csv 2>/dev/null
ls -la gpurun_out/r2_*ppo*
