P="python scripts/ppo_prof.py"
$P > gpurun_out/plain_ppo2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"brb_ppo_grad" -s 8 -c 2 -f -o gpurun_out/prof_ppo2 $P > gpurun_out/ncu_ppo2.log 2>&1
ncu -i gpurun_out/prof_ppo2.ncu-rep --page raw --csv > gpurun_out/raw_ppo2.csv 2>/dev/null
ncu -i gpurun_out/prof_ppo2.ncu-rep --page source --csv > gpurun_out/src_ppo2.csv 2>/dev/null
ls -la gpurun_out/*ppo2*
