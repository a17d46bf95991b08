"""Kernel-tuning experiment: what would time-slicing the 250 substeps into shorter launches (with a re-sort of the visit
order between slices) buy?  Emulated with the existing kernel by compiling the model with a smaller frame_skip and
counting `250 / frame_skip` launches as one step (the task-logic prologue/epilogue runs per launch, so this is an upper
bound on the overhead)."""
import sys
sys.path.insert(0, ".")
import torch
from balance_robot_b200 import make_vec, model as model_mod

orig = model_mod.compile_model
for n in (65536, 1048576):
    for fs in (250, 125, 50, 25):
        model_mod.compile_model = lambda spec, kind, mes, **kw: orig(spec, kind, mes * (250 // fs), frame_skip=fs, **kw)
        env = make_vec("Env01-v2", n, seed=0)
        env.reset()
        gen = torch.Generator(device="cuda").manual_seed(1234)
        acts = [torch.rand((n, 2), device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
        per = 250 // fs
        for k in range(60 * per): env.step(acts[(k // per) % 8])
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        K = (40 if n < 10**6 else 10)
        s0 = env.stats()
        e0.record()
        for k in range(K * per): env.step(acts[(k // per) % 8])
        e1.record(); torch.cuda.synchronize()
        s1 = env.stats()
        t = e0.elapsed_time(e1) / K
        sub = s1["substeps"] - s0["substeps"]; con = s1["contact_substeps"] - s0["contact_substeps"]
        print(f"n={n} frame_skip={fs} ms per 250 substeps={t:.3f} ({n/t*1e3:.3e}/s) contact frac={con/sub:.3f} solves/contact={(s1['solves']-s0['solves'])/max(1,con):.3f}", flush=True)
        env.close()
