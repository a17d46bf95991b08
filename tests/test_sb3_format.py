"""Stable-Baselines3 model.zip interchange (reference src/sb_rl.py:519-525 `algorithm_class.load(model_file,
env=env)`, :542 best_model.zip, :545-550 checkpoints).  SB3 / gymnasium / cloudpickle are not installable here, so:
  * a zip in SB3's exact layout is written BY HAND (SB3's attribute names in `data`, cloudpickled entries present as opaque
    ":serialized:" blobs, no key of this package in it) and must load here and reproduce the fp32 forward pass of the weights;
  * the zips this package writes are taken apart the way SB3's load does: json_to_data (base64 -> pickle.loads) against stand-in
    `stable_baselines3` / `gymnasium` modules, then the attribute set BaseAlgorithm.load and PPO._setup_model read is checked."""
import base64
import io
import json
import pickle
import sys
import types
import zipfile

import numpy as np
import pytest
import torch

from balance_robot_b200 import sb3_format
from balance_robot_b200.ppo import PPO, MlpPolicy, PPOConfig

SB3_KEYS = ["mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias", "mlp_extractor.policy_net.2.weight",
            "mlp_extractor.policy_net.2.bias", "mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias",
            "mlp_extractor.value_net.2.weight", "mlp_extractor.value_net.2.bias", "action_net.weight", "action_net.bias",
            "value_net.weight", "value_net.bias", "log_std"]


class _Env:
    num_envs = 4
    device = "cpu"

    class _Sp:
        def __init__(self, lo, hi):
            self.low, self.high = np.array(lo, np.float32), np.array(hi, np.float32)
    observation_space = _Sp([-6.2831855, -6.2831855, -1, -1, -1, -1], [6.2831855, 6.2831855, 1, 1, 1, 1])
    action_space = _Sp([-1, -1], [1, 1])


def _blob(obj):
    bio = io.BytesIO()
    torch.save(obj, bio)
    return bio.getvalue()


def test_zip_in_sb3_layout_written_by_hand_loads_and_reproduces_the_policy(tmp_path):
    rng = np.random.default_rng(0)
    shapes = [(64, 6), (64,), (64, 64), (64,), (64, 6), (64,), (64, 64), (64,), (2, 64), (2,), (1, 64), (1,), (2,)]
    sd = {k: torch.tensor(rng.normal(0, 0.3, s), dtype=torch.float32) for k, s in zip(SB3_KEYS, shapes)}
    opaque = {":type:": "<class 'function'>", ":serialized:": base64.b64encode(b"\x80\x05cloudpickle-bytes-we-cannot-read").decode()}
    data = {"policy_class": dict(opaque, **{":type:": "<class 'abc.ABCMeta'>"}), "device": dict(opaque), "verbose": 1, "policy_kwargs": {},
            "num_timesteps": 123456, "_total_timesteps": 10000000000, "seed": None, "learning_rate": 0.0003, "lr_schedule": dict(opaque),
            "observation_space": dict(opaque), "action_space": dict(opaque), "n_envs": 1, "n_steps": 2048, "gamma": 0.99, "gae_lambda": 0.95,
            "ent_coef": 0.0, "vf_coef": 0.5, "max_grad_norm": 0.5, "batch_size": 64, "n_epochs": 10, "clip_range": dict(opaque),
            "clip_range_vf": None, "normalize_advantage": True, "target_kl": None, "_n_updates": 600, "use_sde": False}
    path = tmp_path / "best_model.zip"
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps(data))
        z.writestr("pytorch_variables.pth", _blob(None))
        z.writestr("policy.pth", _blob(sd))
        z.writestr("_stable_baselines3_version", "2.4.0a5")
        z.writestr("system_info.txt", "- OS: macOS\n")
    agent = PPO.load(path, _Env(), device="cpu")
    assert agent.num_timesteps == 123456
    assert agent.cfg.n_steps == 2048 and agent.cfg.n_epochs == 10 and agent.cfg.gamma == 0.99 and agent.cfg.n_minibatches == 32   # 2048 x 1 / 64
    assert agent.cfg.clip_range == 0.2 and agent.cfg.learning_rate == 0.0003          # pickled schedule -> default; float kept
    obs = torch.tensor(rng.normal(0, 1, (50, 6)), dtype=torch.float32)
    a, _ = agent.policy.predict(obs, deterministic=True)
    W = {k: v.numpy().astype(np.float64) for k, v in sd.items()}
    h = np.tanh(obs.numpy() @ W[SB3_KEYS[0]].T + W[SB3_KEYS[1]])
    h = np.tanh(h @ W[SB3_KEYS[2]].T + W[SB3_KEYS[3]])
    expect = np.clip(h @ W["action_net.weight"].T + W["action_net.bias"], -1, 1)
    np.testing.assert_allclose(a.numpy(), expect, atol=2e-6)


def test_reference_policy_weights_round_trip_through_an_sb3_layout_zip(tmp_path):
    """The reference's own trained policy (RobotMovePolicy.tflite, extracted to tests/golden/robot_move_policy.npz) as an SB3 zip:
    dequantised weights -> state dict with SB3's key names -> PPO.load -> same actions as the fp32 forward pass of those weights."""
    import pathlib
    import reference_policy as rp
    g = np.load(pathlib.Path(__file__).parent / "golden" / "robot_move_policy.npz")
    (W1, b1), (W2, b2), (W3, b3) = rp.dequantised_layers(g)
    pol = MlpPolicy()
    sd = pol.state_dict()
    sd["mlp_extractor.policy_net.0.weight"], sd["mlp_extractor.policy_net.0.bias"] = torch.tensor(W1, dtype=torch.float32), torch.tensor(b1, dtype=torch.float32)
    sd["mlp_extractor.policy_net.2.weight"], sd["mlp_extractor.policy_net.2.bias"] = torch.tensor(W2, dtype=torch.float32), torch.tensor(b2, dtype=torch.float32)
    sd["action_net.weight"], sd["action_net.bias"] = torch.tensor(W3, dtype=torch.float32), torch.tensor(b3, dtype=torch.float32)
    path = tmp_path / "Env01-v2_PPO.zip"
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps({"num_timesteps": 1, "n_steps": 2048, "n_envs": 1, "batch_size": 64}))
        z.writestr("policy.pth", _blob(sd))
    agent = PPO.load(path, _Env(), device="cpu")
    obs = np.random.default_rng(1).normal(0, 0.5, (64, 6)).astype(np.float32)
    a, _ = agent.policy.predict(torch.tensor(obs), deterministic=True)
    h = np.tanh(np.tanh(obs @ W1.T + b1) @ W2.T + b2)
    np.testing.assert_allclose(a.numpy(), np.clip(h @ W3.T + b3, -1, 1), atol=5e-6)
    # and the loaded fp32 policy agrees with the reference's int8 interpreter graph up to its quantisation noise
    a_int8 = rp.RobotMovePolicy().act(torch.tensor(obs)).numpy()
    assert np.abs(np.clip(a_int8, -1, 1) - a.numpy()).mean() < 0.05 and np.corrcoef(a_int8.ravel(), a.numpy().ravel())[0, 1] > 0.95


def test_saved_zip_has_what_sb3_load_reads(tmp_path, monkeypatch):
    agent = PPO(_Env(), PPOConfig(n_steps=16, n_minibatches=2, seed=3), device="cpu")
    agent.num_timesteps = 4242
    path = agent.save(tmp_path / "m")
    with zipfile.ZipFile(path) as z:
        names = set(z.namelist())
        raw = json.loads(z.read("data"))
        sd = torch.load(io.BytesIO(z.read("policy.pth")), weights_only=False)
        opt = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), weights_only=False)
        assert z.read("_stable_baselines3_version").decode() == sb3_format.SB3_VERSION
    assert {"data", "policy.pth", "policy.optimizer.pth", "pytorch_variables.pth", "_stable_baselines3_version", "system_info.txt"} <= names
    assert list(sd.keys())[0] == "log_std" and set(sd.keys()) == set(SB3_KEYS)           # SB3's ActorCriticPolicy.state_dict() names and order
    assert len(opt["param_groups"][0]["params"]) == len(SB3_KEYS) and opt["param_groups"][0]["eps"] == 1e-5
    # --- what BaseAlgorithm.load / PPO._setup_model read
    for key in ("policy_class", "observation_space", "action_space", "policy_kwargs", "verbose", "n_envs", "n_steps", "gamma", "gae_lambda",
                "ent_coef", "vf_coef", "max_grad_norm", "batch_size", "n_epochs", "clip_range", "clip_range_vf", "normalize_advantage",
                "target_kl", "learning_rate", "use_sde", "sde_sample_freq", "seed", "num_timesteps", "_n_updates", "tensorboard_log",
                "rollout_buffer_class", "rollout_buffer_kwargs", "_stats_window_size"):
        assert key in raw, key
    assert raw["num_timesteps"] == 4242 and raw["batch_size"] == 16 * 4 // 2 and raw["policy_kwargs"] == {}
    assert isinstance(raw["clip_range"], float) and isinstance(raw["learning_rate"], float)     # get_schedule_fn accepts floats
    # --- SB3's json_to_data: base64 -> (cloud)pickle.loads; run it against stand-in modules
    class Box:
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class ActorCriticPolicy:
        pass
    mods = {n: types.ModuleType(n) for n in ("gymnasium", "gymnasium.spaces", "gymnasium.spaces.box", "stable_baselines3", "stable_baselines3.common",
                                             "stable_baselines3.common.policies")}
    mods["gymnasium.spaces.box"].Box = Box
    mods["stable_baselines3.common.policies"].ActorCriticPolicy = ActorCriticPolicy
    for n, m in mods.items():
        monkeypatch.setitem(sys.modules, n, m)
    objs = {k: pickle.loads(base64.b64decode(v[":serialized:"])) for k, v in raw.items() if isinstance(v, dict) and ":serialized:" in v}
    assert objs["policy_class"] is ActorCriticPolicy
    osp, asp = objs["observation_space"], objs["action_space"]
    assert isinstance(osp, Box) and osp.low.dtype == np.float32 and osp.dtype is np.float32 and osp.shape is None
    np.testing.assert_array_equal(osp.low, _Env.observation_space.low)
    np.testing.assert_array_equal(asp.high, np.array([1, 1], np.float32))
    # --- and it loads back here with the same hyper-parameters and weights
    again = PPO.load(path, _Env(), device="cpu")
    assert again.cfg == agent.cfg and again.num_timesteps == 4242
    for (k1, v1), (k2, v2) in zip(agent.policy.state_dict().items(), again.policy.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
