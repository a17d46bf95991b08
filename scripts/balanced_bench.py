"""Throughput of the env step when every robot is balancing (the reference's own policy drives them): the heaviest workload
for the kernel — all wheels on the floor all the time — next to the random-action workload of bench.py."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from balance_robot_b200 import make_vec
from reference_policy import RobotMovePolicy
env_id = sys.argv[1] if len(sys.argv) > 1 else "Env01-v1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
env = make_vec(env_id, n, seed=0); pol = RobotMovePolicy("cuda:0")
obs = env.reset()
for _ in range(300):
    obs = env.step(pol.act(obs))[0]
s0 = env.stats()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
for a, b in ev:
    act = pol.act(obs)
    a.record(); obs = env.step(act)[0]; b.record()
torch.cuda.synchronize()
s1 = env.stats()
ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
print(f"{env_id} n={n} balanced robots: {ms:.3f} ms/step = {n / ms * 1e3:.3e} env-steps/s; contact-active fraction "
      f"{(s1['contact_substeps'] - s0['contact_substeps']) / (s1['substeps'] - s0['substeps']):.3f}, slots per contact substep "
      f"{(s1['contact_slots'] - s0['contact_slots']) / max(1, s1['contact_substeps'] - s0['contact_substeps']):.2f}, episodes ended {s1['episodes'] - s0['episodes']}")
