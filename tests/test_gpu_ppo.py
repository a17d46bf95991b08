"""-m gpu: on-device PPO against the real CUDA envs (rollout buffer, GAE and updates never leave the GPU)."""
import pathlib
import subprocess
import sys

import pytest
import torch

from balance_robot_b200 import make_vec
from balance_robot_b200.ppo import PPO, PPOConfig, evaluate_policy

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_ppo_improves_return_on_env01_v1():
    env = make_vec("Env01-v1", 2048, seed=0)
    agent = PPO(env, PPOConfig(n_steps=32, seed=0), device="cuda:0")
    first = agent.collect_rollouts()
    agent.train()
    for _ in range(24):
        last = agent.collect_rollouts()
        agent.train()
    assert agent.buf["obs"].is_cuda and agent.buf["adv"].is_cuda
    assert last["ep_rew_mean"] > 1.15 * first["ep_rew_mean"], (first, last)
    mean_r, std_r, lens = evaluate_policy(agent.policy, make_vec("Env01-v1", 64, seed=5), 10, True, 6000)
    assert mean_r > first["ep_rew_mean"]
    env.close()


def test_cli_train_writes_reference_style_artifacts(tmp_path):
    cmd = [sys.executable, str(ROOT / "sb_rl.py"), "-a", "PPO", "train", "-e", "Env01-v2", "--num-envs", "1024",
           "--total-timesteps", "200000", "--n-steps", "16"]
    res = subprocess.run(cmd, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    run = tmp_path / "models" / "Env01-v2_PPO"
    assert (run / "best_model.zip").exists()                                   # EvalCallback target, sb_rl.py:542
    assert list(run.glob("Env01-v2_PPO_cp__*_steps.zip"))                      # CheckpointCallback naming, sb_rl.py:545-550
    assert (tmp_path / "logs").is_dir() and (tmp_path / "movies").is_dir()
    # fine-tune from the saved model with -m on Env03-v2, as README.md:62 does (BASELINE.json configs[3])
    cmd2 = [sys.executable, str(ROOT / "sb_rl.py"), "-a", "PPO", "-m", str(run / "best_model.zip"), "train", "-e", "Env03-v2",
            "--num-envs", "512", "--total-timesteps", "20000", "--n-steps", "16"]
    res2 = subprocess.run(cmd2, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res2.returncode == 0, res2.stderr[-3000:]
    assert (tmp_path / "models" / "Env03-v2_PPO" / "Env03-v2_PPO_final.zip").exists()


def test_policy_rollout_throughput_path_runs():
    """configs[2]-style rollout: Env01-v3 with on-device policy inference feeding the step kernel (small size here)."""
    env = make_vec("Env01-v3", 8192, seed=1)
    pol = PPO(env, PPOConfig(n_steps=4, seed=1), device="cuda:0").policy
    obs = env.reset()
    for _ in range(10):
        a, _, _ = pol.act(obs)
        obs, r, d, info = env.step(a.clamp(-1, 1))
    assert torch.isfinite(obs).all() and torch.isfinite(r).all()
    env.close()
