"""Kernel-tuning experiment: per warp-task timeline of one step launch (build with -DBRB_TIMELINE)."""
import os, sys, ctypes as C
os.environ["BRB_EXPERIMENT_LIB"] = "scripts/variant_timeline.so"
sys.path.insert(0, ".")
import numpy as np, torch
from balance_robot_b200 import make_vec, _cabi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = make_vec("Env01-v2", n, seed=0); env.reset()
gen = torch.Generator(device="cuda").manual_seed(1234)
acts = [torch.rand((n, 2), device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
L = _cabi.lib()
buf = np.zeros(1 + 4 * 65536 + 65536, np.uint64)
for k in range(100): env.step(acts[k % 8])
L.brb_timeline(None, 1)
env.step(acts[0])
L.brb_timeline(buf.ctypes.data_as(C.c_void_p), 1)
m = int(buf[0]); r = buf[1:1 + 4 * m].reshape(m, 4)
sm = (r[:, 0] >> np.uint64(32)).astype(int); wid = (r[:, 0] & np.uint64(0xFFFFFFFF)).astype(int); first = r[:, 1].astype(np.int64)
t0 = r[:, 2].astype(np.int64); t1 = r[:, 3].astype(np.int64)
z = t0.min(); t0 = (t0 - z) / 1e3; t1 = (t1 - z) / 1e3
print(f"tasks {m}, kernel span {t1.max():.1f} us; task start: min {t0.min():.1f} median {np.median(t0):.1f} max {t0.max():.1f}")
dur = t1 - t0
order = np.argsort(first)
print("task duration (us) by position in the visit order (deciles):", [round(float(np.mean(dur[order][k * m // 10:(k + 1) * m // 10])), 1) for k in range(10)])
print("task end (us) percentiles 50/90/99/100:", [round(float(np.percentile(t1, p)), 1) for p in (50, 90, 99, 100)])
# per SM sub-partition: warp slot id mod 4 is the scheduler
key = sm * 4 + (wid % 4)
busy = {}
for k, a, b in zip(key, t0, t1): busy.setdefault(k, []).append((a, b))
ends = np.array([max(b for _, b in v) for v in busy.values()]); cnt = np.array([len(v) for v in busy.values()])
print(f"schedulers used {len(busy)}; tasks per scheduler min/mean/max {cnt.min()}/{cnt.mean():.2f}/{cnt.max()}; last task end per scheduler: percentiles 10/50/90/100 =",
      [round(float(np.percentile(ends, p)), 1) for p in (10, 50, 90, 100)])
# how many warps are resident over time
ts = np.linspace(0, t1.max(), 23)[1:-1]
print("resident warp-tasks over time:", [(round(float(t)), int(((t0 <= t) & (t1 > t)).sum())) for t in ts])
second = t0 > 5.0
print(f"tasks started after 5 us: {second.sum()}, their mean duration {dur[second].mean() if second.any() else 0:.1f} us, mean start {t0[second].mean() if second.any() else 0:.1f}")
tr = buf[1 + 4 * 65536:].view(np.uint32).reshape(-1, 2)
trips = tr[first // 32, 0].astype(int); extra = tr[first // 32, 1].astype(int)
ex = dur > 0.8 * np.percentile(dur, 60)
print("expensive tasks:", ex.sum(), "trips percentiles 10/50/90/99/100:", [int(np.percentile(trips[ex], p)) for p in (10, 50, 90, 99, 100)],
      "us per trip:", round(float((dur[ex] / trips[ex]).mean()), 3), "+-", round(float((dur[ex] / trips[ex]).std()), 3))
print("corr(duration, trips) among expensive:", round(float(np.corrcoef(dur[ex], trips[ex])[0, 1]), 3), " extra solves summed over lanes, mean:", extra[ex].mean())
top = np.argsort(-dur)[:10]
print("10 longest tasks: dur", dur[top].round(0), "trips", trips[top], "sm", sm[top])
np.save("gpurun_out/r2_timeline.npy", r)
