# ncu captures of the PPO kernels inside one small PPO iteration (scripts/ppo_prof.py)
P="python scripts/ppo_prof.py"
$P > gpurun_out/plain_ppo2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ppo.csv $P > gpurun_out/ncu_ppo1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"brb_ppo_grad|brb_policy_act" -s 20 -c 3 -f -o gpurun_out/prof_ppo2 $P > gpurun_out/ncu_ppo2.log 2>&1
ncu -i gpurun_out/prof_ppo2.ncu-rep --page raw --csv > gpurun_out/raw_ppo2.csv 2>/dev/null
ls -la gpurun_out/*ppo*
