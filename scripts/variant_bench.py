"""Kernel-tuning experiment: time alternative builds (scripts/variant_*.so) of the same sources."""
import os, subprocess, sys, glob
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ".")
    import torch
    from balance_robot_b200 import make_vec
    for n in (65536, 1048576):
        env = make_vec("Env01-v2", n, seed=0)
        env.reset()
        gen = torch.Generator(device="cuda").manual_seed(1234)
        acts = [torch.rand((n, 2), device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
        for k in range(60): env.step(acts[k % 8])
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        K = 40 if n < 10**6 else 10
        e0.record()
        for k in range(K): env.step(acts[k % 8])
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / K
        st = env.stats()
        print(f"  n={n} ms/step={t:.3f} env-steps/s={n/t*1e3:.3e} solves/contact-substep={st['solves']/max(1,st['contact_substeps']):.3f} nonconv={st['nonconverged']}", flush=True)
        env.close()
else:
    libs = sorted(glob.glob("scripts/variant_*.so")) if len(sys.argv) < 2 else sys.argv[1:]
    for lib in libs:
        print(lib, flush=True)
        env = dict(os.environ, BRB_EXPERIMENT_LIB=lib)
        subprocess.run([sys.executable, __file__, "child"], env=env)
