"""Kernel-tuning experiment: where one PPO update (10 epochs x 4 minibatches of 4M samples at 1M envs x 16 steps) spends its time."""
import sys, time
sys.path.insert(0, ".")
import torch
from balance_robot_b200 import make_vec
from balance_robot_b200.ppo import PPO, PPOConfig

def timed(fn, reps=5):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

n = 1 << 20
env = make_vec("Env01-v2", n, seed=0)
agent = PPO(env, PPOConfig(n_steps=16, seed=0), device="cuda:0")
for _ in range(2): agent.collect_rollouts()
agent.train()
total = 16 * n; mb = total // 4
gen = agent.gen
print(f"whole train(): {timed(agent.train, 2):.2f} ms")
print(f"randperm({total}): {timed(lambda: torch.randperm(total, device='cuda', generator=gen)):.3f} ms")
perm = torch.randperm(total, device="cuda", generator=gen)
adv = agent.buf["adv"].reshape(total).contiguous()
idx = perm[:mb]
def stats():
    a = adv[idx]; sd, mu = torch.std_mean(a); return torch.stack([mu, 1.0 / (sd + 1e-8)])
print(f"adv[idx] + std_mean + stack ({mb} samples): {timed(stats):.3f} ms")
print(f"adv[idx] alone: {timed(lambda: adv[idx]):.3f} ms")
