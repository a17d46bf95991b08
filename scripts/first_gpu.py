import sys, time, torch, numpy as np
sys.path.insert(0, ".")
import __graft_entry__ as g
g.smoke()
from balance_robot_b200 import make_vec, _cabi
import ctypes as C
fl = C.c_double(); ms = C.c_double()
print("fp32 probe rc", _cabi.lib().brb_fp32_peak_flops(0, C.byref(fl), C.byref(ms)), fl.value/1e12, "TFLOP/s", ms.value, "ms")
for n in (65536, 262144, 1048576):
    env = make_vec("Env01-v2", n, seed=0)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(1234)
    acts = [torch.rand((n,2), device="cuda", generator=gen)*2-1 for _ in range(8)]
    for k in range(20): env.step(acts[k%8])
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    K=30
    e0.record()
    for k in range(K): env.step(acts[k%8])
    e1.record(); torch.cuda.synchronize()
    t=e0.elapsed_time(e1)/K
    print(n, "ms/step", t, "env-steps/s %.3e"%(n/t*1e3), env.stats())
    env.close()
