"""Experiment: how fast does on-device PPO learn to balance?  (feeds the thresholds in tests/test_gpu_ppo.py)"""
import sys, time
sys.path.insert(0, ".")
import torch
from balance_robot_b200 import make_vec
from balance_robot_b200.ppo import PPO, PPOConfig, evaluate_policy
env_id = sys.argv[1] if len(sys.argv) > 1 else "Env01-v1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 60
env = make_vec(env_id, n, seed=0)
agent = PPO(env, PPOConfig(n_steps=32, seed=0), device="cuda:0")
t0 = time.time()
agent.learn(iters * 32 * n, log_interval=5)
print("wall", time.time() - t0)
ev = make_vec(env_id, 64, seed=123)
print("eval", evaluate_policy(agent.policy, ev, 20, True, 6000))
