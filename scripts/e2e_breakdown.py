"""Kernel-tuning experiment: where the host-buffer step (BalanceVecEnv(output='numpy').step) spends its time."""
import os, sys, time
os.environ["BRB_PROFILE_HOST"] = "1"
sys.path.insert(0, ".")
import numpy as np, torch
from balance_robot_b200 import make_vec
n = 65536
env = make_vec("Env01-v2", n, seed=0, output="numpy"); env.reset()
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (n, 2)).astype(np.float32) for _ in range(8)]
for k in range(300): env.step(acts[k % 8])
t0 = time.perf_counter()
K = 512
for k in range(K): env.step(acts[k % 8])
t1 = time.perf_counter()
print(f"python step(): {(t1 - t0) / K * 1e6:.1f} us per step ({n * K / (t1 - t0):.3e} env-steps/s)")
t0 = time.perf_counter()
for k in range(K): env._h_act.numpy()[...] = acts[k % 8]
t1 = time.perf_counter()
print(f"copy of the actions into the pinned buffer: {(t1 - t0) / K * 1e6:.1f} us")
