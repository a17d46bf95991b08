"""Kernel-tuning experiment: warp-trip lane statistics of the step kernel (build with -DBRB_TRIPSTATS)."""
import os, sys, ctypes as C
os.environ["BRB_EXPERIMENT_LIB"] = "scripts/variant_tripstats.so"
sys.path.insert(0, ".")
import torch
from balance_robot_b200 import make_vec, _cabi
n = 65536
env_id = sys.argv[1] if len(sys.argv) > 1 else "Env01-v2"
env = make_vec(env_id, n, seed=0); env.reset()
gen = torch.Generator(device="cuda").manual_seed(1234)
acts = [torch.rand((n, 2), device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
L = _cabi.lib()
out = (C.c_ulonglong * 8)()
for k in range(60): env.step(acts[k % 8])
torch.cuda.synchronize(); L.brb_tripstats(out); a = list(out)
for k in range(20): env.step(acts[k % 8])
torch.cuda.synchronize(); L.brb_tripstats(out); b = list(out)
d = [y - x for x, y in zip(a, b)]
wt, lt, ws, ls = d[:4]
print("warp-trips/step/warp", wt / 20 / (n / 32), "lanes per trip", lt / wt, "warp-trips with solve frac", ws / wt, "lanes solving per solve-trip", ls / ws)
print(env.stats())
