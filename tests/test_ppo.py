"""On-device PPO pieces on CPU tensors: SB3-compatible policy layout, GAE against a scalar re-derivation, the clipped
update on a toy VecEnv, TimeLimit bootstrap, checkpoint round trip, and the world_size-2 gloo gradient all-reduce."""
import math
import os
import pathlib
import subprocess
import sys

import numpy as np
import pytest
import torch

from balance_robot_b200.ppo import PPO, MlpPolicy, PPOConfig, compute_gae, evaluate_policy

ROOT = pathlib.Path(__file__).resolve().parent.parent


class ToyInfos:
    def __init__(self, n):
        self.truncated = torch.zeros(n, dtype=torch.uint8)
        self.terminal_observation = torch.zeros((n, 6))
        self.episode_return = torch.zeros(n)
        self.episode_length = torch.zeros(n, dtype=torch.int32)


class ToyVecEnv:
    """reward = -|a0 - obs0|; episodes of fixed length `horizon` ending by truncation."""

    def __init__(self, n, horizon=8, seed=0):
        self.num_envs, self.horizon, self.device = n, horizon, torch.device("cpu")
        self.g = torch.Generator().manual_seed(seed)
        self.t = torch.zeros(n, dtype=torch.int32)
        self.ret = torch.zeros(n)

    def _new_obs(self):
        o = torch.zeros((self.num_envs, 6))
        o[:, 0] = torch.rand(self.num_envs, generator=self.g) - 0.5
        return o

    def reset(self):
        self.obs = self._new_obs()
        return self.obs

    def step(self, a):
        rew = -(a[:, 0] - self.obs[:, 0]).abs()
        self.t += 1
        self.ret += rew
        done = (self.t >= self.horizon).to(torch.uint8)
        infos = ToyInfos(self.num_envs)
        infos.truncated = done.clone()
        infos.terminal_observation = self.obs.clone()
        infos.episode_return, infos.episode_length = self.ret.clone(), self.t.clone()
        self.obs = self._new_obs()
        self.t[done.bool()] = 0
        self.ret[done.bool()] = 0
        return self.obs, rew, done, infos


def test_policy_layout_matches_sb3():
    p = MlpPolicy()
    keys = set(p.state_dict())
    assert keys == {"log_std", "mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias",
                    "mlp_extractor.policy_net.2.weight", "mlp_extractor.policy_net.2.bias",
                    "mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias",
                    "mlp_extractor.value_net.2.weight", "mlp_extractor.value_net.2.bias",
                    "action_net.weight", "action_net.bias", "value_net.weight", "value_net.bias"}
    assert sum(v.numel() for v in p.parameters()) == 9413
    assert torch.all(p.log_std == 0)
    w = p.action_net.weight
    assert torch.allclose(w @ w.T, 0.01 ** 2 * torch.eye(2), atol=1e-6)      # orthogonal, gain 0.01
    obs = torch.randn(5, 6)
    a, v, lp = p.act(obs, generator=torch.Generator().manual_seed(0))
    v2, lp2, ent = p.evaluate_actions(obs, a)
    assert torch.allclose(lp, lp2, atol=1e-6) and torch.allclose(v, v2)
    assert ent[0].item() == pytest.approx(2 * (0.5 + 0.5 * math.log(2 * math.pi)))
    det, _ = p.predict(obs.numpy())
    assert det.abs().max() <= 1.0


def test_gae_matches_scalar_recursion():
    rng = np.random.default_rng(0)
    T, n, g, lam = 7, 3, 0.99, 0.95
    r, v = rng.normal(size=(T, n)), rng.normal(size=(T, n))
    d = (rng.random((T, n)) < 0.3).astype(np.float64)
    lv = rng.normal(size=n)
    adv, ret = compute_gae(*(torch.tensor(x) for x in (r, v, d)), torch.tensor(lv), g, lam)
    for k in range(n):
        last = 0.0
        for t in reversed(range(T)):
            nv = lv[k] if t == T - 1 else v[t + 1, k]
            delta = r[t, k] + g * nv * (1 - d[t, k]) - v[t, k]
            last = delta + g * lam * (1 - d[t, k]) * last
            assert adv[t, k].item() == pytest.approx(last, abs=1e-12)
    assert torch.allclose(ret, adv + torch.tensor(v))


def test_ppo_learns_toy_task_and_bootstraps_truncations():
    env = ToyVecEnv(256, horizon=8, seed=1)
    agent = PPO(env, PPOConfig(n_steps=16, n_epochs=4, n_minibatches=4, learning_rate=3e-3, seed=3), device="cpu")
    first = agent.collect_rollouts()
    # every episode ends by truncation: the stored reward carries gamma * V(terminal obs)
    assert torch.any(agent.buf["dones"] > 0)
    agent.train()
    for _ in range(25):
        last = agent.collect_rollouts()
        logs = agent.train()
    assert last["ep_rew_mean"] > first["ep_rew_mean"] + 1.0
    assert 0 <= logs["clip_fraction"] <= 1 and logs["approx_kl"] >= -1e-6
    assert agent.num_timesteps == 26 * 16 * 256
    m, s, lens = evaluate_policy(agent.policy, ToyVecEnv(16, 8, seed=2), 5, True, 50)
    assert m > first["ep_rew_mean"] and all(l == 8 for l in lens)


def test_checkpoint_round_trip(tmp_path):
    env = ToyVecEnv(8)
    a = PPO(env, PPOConfig(n_steps=4, seed=1), device="cpu")
    a.collect_rollouts(); a.train()
    path = a.save(tmp_path / "Env01-v2_PPO" / "best_model")
    assert path.name == "best_model.zip"
    import zipfile
    with zipfile.ZipFile(path) as z:
        assert {"data", "policy.pth", "policy.optimizer.pth", "pytorch_variables.pth"} <= set(z.namelist())
    b = PPO.load(path, ToyVecEnv(8), device="cpu")
    for k, v in a.policy.state_dict().items():
        assert torch.equal(v, b.policy.state_dict()[k])
    assert b.num_timesteps == a.num_timesteps and b.cfg.n_steps == 4
    with pytest.raises(RuntimeError):
        PPO.load(tmp_path / "missing.zip", env)


def test_cli_surface():
    from click.testing import CliRunner
    from balance_robot_b200.sb_rl import cli
    r = CliRunner()
    assert r.invoke(cli, ["train", "-e", "Env01-v2"]).exit_code != 0                      # -a is required, as in the reference
    assert "not a Stable Baselines3 algorithm" in r.invoke(cli, ["-a", "XYZ", "train", "-e", "Env01-v2"]).output
    assert "only PPO" in r.invoke(cli, ["-a", "SAC", "train", "-e", "Env01-v2"]).output
    assert "outside the B200 hot-path scope" in r.invoke(cli, ["-a", "PPO", "convert", "-e", "Env01-v2"]).output
    res = r.invoke(cli, ["-a", "PPO", "test", "-e", "Env01-v2"])                          # default model name, missing here (sb_rl.py:147-152)
    assert isinstance(res.exception, RuntimeError) and "Could not open model file" in str(res.exception)
    out = r.invoke(cli, ["-a", "PPO", "train", "--help"]).output
    assert "--environment" in out and "--num-envs" in out


DIST_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from test_ppo import ToyVecEnv
from balance_robot_b200.ppo import PPO, PPOConfig
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
env = ToyVecEnv(64, horizon=8, seed=100 + rank)          # each rank owns its own shard of envs
agent = PPO(env, PPOConfig(n_steps=8, n_epochs=2, n_minibatches=2, seed=5), device="cpu", rank=rank, world_size=world)
stats = agent.collect_rollouts()
agent.train()
flat = torch.cat([p.detach().reshape(-1) for p in agent.policy.parameters()])
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
assert all(torch.equal(gathered[0], g) for g in gathered), "replicas diverged after all-reduced updates"
assert agent.num_timesteps == 8 * 64 * world
assert stats["episodes"] == 64 * world                    # rollout statistics are summed over ranks
if rank == 0:
    print("DIST_OK", float(flat.abs().sum()))
dist.destroy_process_group()
'''


def test_world_size_2_gloo_gradient_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(DIST_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), str(script), str(ROOT)]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    assert "DIST_OK" in res.stdout
