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
out = (C.c_ulonglong * 194)()
for k in range(60): env.step(acts[k % 8])
torch.cuda.synchronize(); L.brb_tripstats(out); a = list(out)
for k in range(20): env.step(acts[k % 8])
torch.cuda.synchronize(); L.brb_tripstats(out); b = list(out)
d = [y - x for x, y in zip(a, b)]
wt, lt, ws, ls = d[:4]
print('trips by lanes in contact (share of trips):', [round(x / wt, 3) for x in d[8:41]])
print("warp-trips/step/warp", wt / 20 / (n / 32), "lanes per trip", lt / wt, "warp-trips with solve frac", ws / wt, "lanes solving per solve-trip", ls / ws)
print('robots per step by incoming group key: [no contact, <half, >=half, full step in contact]')
for k in range(18):
    print('  key', k, [round(x / 20) for x in d[48 + 4 * k:52 + 4 * k]])
tb = sum(d[176:192])
print('pyramid-row patterns of contacts at solves (share):', [round(x / tb, 4) for x in d[176:192]])
print('solves with every contact fully active:', d[193] / d[192])
print(env.stats())
