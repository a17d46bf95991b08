"""BASELINE.json configs 3-5 on one GPU: rollouts with on-device policy inference and PPO rollout + update timings.
Prints one JSON line per config (device-timed with CUDA events around whole phases, after one untimed iteration)."""
import json, sys, time
sys.path.insert(0, ".")
import torch
from balance_robot_b200 import make_vec
from balance_robot_b200.ppo import PPO, PPOConfig


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); out = fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3, out


def ppo_case(tag, env_id, n, n_steps, iters=2, spinup=3):
    env = make_vec(env_id, n, seed=0)
    agent = PPO(env, PPOConfig(n_steps=n_steps, seed=0), device="cuda:0")
    for _ in range(spinup):                       # past the all-airborne start, and warms the allocator
        agent.collect_rollouts()
    agent.train()
    tr = tu = 0.0
    for _ in range(iters):
        t, _ = timed(agent.collect_rollouts); tr += t
        t, _ = timed(agent.train); tu += t
    steps = iters * n * n_steps
    print(json.dumps({"config": tag, "env": env_id, "n_envs": n, "n_steps": n_steps, "iterations": iters,
                      "rollout_env_steps_per_s": steps / tr, "rollout_s_per_iter": tr / iters, "update_s_per_iter": tu / iters,
                      "train_env_steps_per_s": steps / (tr + tu), "update_samples_per_s": steps * agent.cfg.n_epochs / tu}), flush=True)
    env.close(); del agent; torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["3", "4", "5"]
    if "3" in which:
        ppo_case("3: Env01-v3, 1M envs, on-device policy inference (PPO rollout; update timed separately)", "Env01-v3", 1 << 20, 16)
    if "4" in which:
        ppo_case("4: Env03-v2 PPO rollout + update, 65,536 envs", "Env03-v2", 65536, 32)
    if "5" in which:
        ppo_case("5 (one GPU's shard): Env01-v2 PPO, 1M envs per GPU", "Env01-v2", 1 << 20, 16)
