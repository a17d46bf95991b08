import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ".")
    import torch
    from balance_robot_b200 import make_vec
    n = 65536
    env = make_vec("Env03-v2", n, seed=0, wheel_block=True); env.reset()
    gen = torch.Generator(device="cuda").manual_seed(1234)
    acts = [torch.rand((n, 2), device="cuda", generator=gen) * 2 - 1 for _ in range(8)]
    for k in range(30): env.step(acts[k % 8])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(10): env.step(acts[k % 8])
    e1.record(); torch.cuda.synchronize()
    print(f"  wheel_block on: {e0.elapsed_time(e1) / 10:.2f} ms/step", flush=True)
else:
    for lib in sys.argv[1:]:
        print(lib, flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, BRB_EXPERIMENT_LIB=lib))
