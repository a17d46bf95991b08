"""Kernel-tuning experiment: Env03-v2 throughput for alternative builds."""
import os, subprocess, sys, glob
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ".")
    import torch
    from balance_robot_b200 import make_vec
    n = 65536
    env = make_vec("Env03-v2", n, seed=0)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(1234)
    acts = [torch.rand((n, 2), device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
    for k in range(30): env.step(acts[k % 8])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record()
    for k in range(K): env.step(acts[k % 8])
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / K
    print(f"  Env03-v2 n={n} ms/step={t:.2f} env-steps/s={n/t*1e3:.3e} nonconv={env.stats()['nonconverged']}", flush=True)
    st = env.stats(); print('   ', {k: st[k] for k in ('substeps','coupled_substeps','coupled_solves','coupled_fallbacks','block_contact_substeps','episodes')}, flush=True)
else:
    for lib in sys.argv[1:]:
        print(lib, flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, BRB_EXPERIMENT_LIB=lib))
