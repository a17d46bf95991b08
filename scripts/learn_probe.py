"""Experiment: does PPO with SB3-like settings learn to balance here?  Few robots, long rollouts (the reference trains ONE env
with n_steps 2048, batch 64, 10 epochs).  Prints the learning curve and a deterministic evaluation."""
import sys, time
sys.path.insert(0, ".")
import torch
from balance_robot_b200 import make_vec
from balance_robot_b200.ppo import PPO, PPOConfig, evaluate_policy
env_id = sys.argv[1] if len(sys.argv) > 1 else "Env01-v1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 256
total = int(float(sys.argv[4])) if len(sys.argv) > 4 else 2_000_000
nmb = int(sys.argv[5]) if len(sys.argv) > 5 else 32
env = make_vec(env_id, n, seed=0)
agent = PPO(env, PPOConfig(n_steps=n_steps, n_minibatches=nmb, seed=0), device="cuda:0")
t0 = time.time()
agent.learn(total, log_interval=max(1, total // (n * n_steps) // 12))
print(f"wall {time.time() - t0:.1f} s for {total} steps ({n} robots x {n_steps} steps per rollout, {nmb} minibatches)")
ev = make_vec(env_id, 64, seed=123)
m, s, lens = evaluate_policy(agent.policy, ev, 20, True, 6000)
print("deterministic eval: mean return", m, "std", s, "mean length", sum(lens) / len(lens))
