"""Kernel-tuning experiment: where the end-to-end (numpy in / numpy out) step time goes."""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
from balance_robot_b200 import make_vec, _cabi
n = 65536
env = make_vec("Env01-v2", n, seed=0, output="numpy")
env.reset()
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (n, 2)).astype(np.float32) for _ in range(8)]
for k in range(60): env.step(acts[k % 8])
K = 100
t0 = time.perf_counter()
for k in range(K): env.step(acts[k % 8])
t1 = time.perf_counter()
print(f"env.step (numpy): {(t1 - t0) / K * 1e3:.3f} ms/step -> {n * K / (t1 - t0):.3e} env-steps/s")
L = _cabi.lib(); hb = env._hbuf[0]
t0 = time.perf_counter()
for k in range(K):
    L.brb_env_step_host_compact(env._env, env._h_act.data_ptr(), hb["obs"].data_ptr(), hb["rew"].data_ptr(), hb["done"].data_ptr(),
                                env._h_ndone.data_ptr(), hb["rows"].data_ptr(), n)
t1 = time.perf_counter()
print(f"brb_env_step_host_compact only: {(t1 - t0) / K * 1e3:.3f} ms/step")
t0 = time.perf_counter()
for k in range(K):
    env._h_act.numpy()[...] = acts[k % 8]
t1 = time.perf_counter(); print(f"copy actions into pinned: {(t1 - t0) / K * 1e3:.3f} ms")
