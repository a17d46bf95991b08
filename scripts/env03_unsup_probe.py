"""Kernel-tuning experiment: how often does Env03-v2 reach a pose whose contacts the kernel does not model?
Builds with -DBRB_PROBE_UNSUP=1/2/4 count one cause each (chassis on the floor / wheel lying flat / wheel within reach of the block)."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ".")
    import torch
    from balance_robot_b200 import make_vec
    n = 16384
    for policy in ("random", "zero"):
        env = make_vec("Env03-v2", n, seed=0); env.reset()
        gen = torch.Generator(device="cuda").manual_seed(1234)
        for k in range(400):
            a = torch.rand((n, 2), device="cuda", generator=gen) * 2 - 1 if policy == "random" else torch.zeros((n, 2), device="cuda")
            env.step(a)
        st = env.stats()
        print(f"  {policy}: env_steps {st['env_steps']} episodes {st['episodes']} unsupported {st['unsupported']} ({st['unsupported'] / st['env_steps']:.2e} of env-steps)", flush=True)
        env.close()
else:
    for lib in sys.argv[1:]:
        print(lib, flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, BRB_EXPERIMENT_LIB=lib))
