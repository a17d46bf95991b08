"""ncu target: one PPO iteration (rollout with the fused policy forward, fused update) at 262,144 robots."""
import sys
sys.path.insert(0, ".")
import torch
from balance_robot_b200 import make_vec
from balance_robot_b200.ppo import PPO, PPOConfig
env = make_vec("Env01-v2", 262144, seed=0)
agent = PPO(env, PPOConfig(n_steps=8, n_epochs=2, n_minibatches=2, seed=0), device="cuda:0")
for _ in range(2):
    agent.collect_rollouts(); agent.train()
torch.cuda.synchronize()
print("ok")
