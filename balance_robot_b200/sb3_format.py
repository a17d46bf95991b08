"""Stable-Baselines3's model.zip wire format (third party; the reference saves / loads it at src/sb_rl.py:519-525, :542, :545-550).

SB3's `save_to_zip_file` writes
    data                         JSON: every attribute of the algorithm object; values that are not JSON-serialisable are stored as
                                 {":type:": str(type), ":serialized:": base64(cloudpickle.dumps(obj)), <attr>: str(...), ...}
    policy.pth                   torch.save(policy.state_dict())
    policy.optimizer.pth         torch.save(policy.optimizer.state_dict())
    pytorch_variables.pth        torch.save(None) for PPO
    _stable_baselines3_version   text
    system_info.txt              text
and `BaseAlgorithm.load` requires data["policy_class"], ["observation_space"], ["action_space"], ["policy_kwargs"], ["verbose"],
then does `model.__dict__.update(data); model._setup_model(); model.set_parameters(params, exact_match=True)`.

SB3, gymnasium and cloudpickle are not importable in this image, so the three objects SB3 insists on unpickling (the policy class
and the two Box spaces) are written as hand-assembled pickle streams that only reference importable GLOBALS:
    policy_class        GLOBAL stable_baselines3.common.policies.ActorCriticPolicy        (how pickle stores any importable class)
    Box spaces          REDUCE gymnasium.spaces.box.Box(numpy.array([...], numpy.float32), numpy.array([...], numpy.float32),
                                                        None, numpy.float32)
which unpickle in any environment that has SB3 / gymnasium / numpy, independent of their versions' private pickle layouts.
Schedules (learning_rate, clip_range) are stored as plain floats: SB3's _setup_model passes them through get_schedule_fn.
Loading into real SB3 cannot be run here (tests/test_sb3_format.py unpickles the streams against stand-in modules and loads a
hand-written SB3-layout zip); the direction SB3 -> this package needs nothing but json + torch.load.
"""
from __future__ import annotations

import base64
import json
import struct
from typing import Any, Dict, Optional, Sequence

SB3_VERSION = "2.4.0a5"        # the gymnasium-1.0.0a1 fork the reference pins (conda-environment.yaml:20) is a 2.4 pre-release


# ---------------------------------------------------------------------------------------------- a tiny pickle assembler (protocol 2)
def _glob(module: str, name: str) -> bytes:
    return b"c" + module.encode() + b"\n" + name.encode() + b"\n"


def _float_list(vals: Sequence[float]) -> bytes:
    return b"]" + b"(" + b"".join(b"G" + struct.pack(">d", float(v)) for v in vals) + b"e"


def _np_array(vals: Sequence[float]) -> bytes:
    """numpy.array([...], numpy.float32)"""
    return _glob("numpy", "array") + b"(" + _float_list(vals) + _glob("numpy", "float32") + b"t" + b"R"


def pickle_global(module: str, name: str) -> bytes:
    return b"\x80\x02" + _glob(module, name) + b"."


def pickle_box(low: Sequence[float], high: Sequence[float]) -> bytes:
    """gymnasium.spaces.box.Box(low, high, None, numpy.float32)"""
    return (b"\x80\x02" + _glob("gymnasium.spaces.box", "Box") + b"(" + _np_array(low) + _np_array(high) + b"N" + _glob("numpy", "float32")
            + b"t" + b"R" + b".")


def _serialized(type_repr: str, blob: bytes, **attrs: str) -> Dict[str, str]:
    return {":type:": type_repr, ":serialized:": base64.b64encode(blob).decode(), **attrs}


# ---------------------------------------------------------------------------------------------- data (write)
def build_data(cfg, num_timesteps: int, n_envs: int, n_updates: int, obs_low, obs_high, act_low, act_high, batch_size: int) -> str:
    """JSON `data` entry with the attribute names of SB3's PPO (stable_baselines3/ppo/ppo.py + common/on_policy_algorithm.py +
    common/base_class.py) and this trainer's values."""
    data: Dict[str, Any] = {
        "policy_class": _serialized("<class 'abc.ABCMeta'>", pickle_global("stable_baselines3.common.policies", "ActorCriticPolicy"),
                                    __module__="stable_baselines3.common.policies"),
        "verbose": 1,
        "policy_kwargs": {},
        "num_timesteps": int(num_timesteps),
        "_total_timesteps": int(num_timesteps),
        "_num_timesteps_at_start": 0,
        "seed": int(cfg.seed),
        "action_noise": None,
        "start_time": 0,
        "learning_rate": float(cfg.learning_rate),
        "tensorboard_log": "logs",
        "_last_obs": None,
        "_last_episode_starts": None,
        "_last_original_obs": None,
        "_episode_num": 0,
        "use_sde": False,
        "sde_sample_freq": -1,
        "_current_progress_remaining": 1.0,
        "_stats_window_size": 100,
        "ep_info_buffer": None,
        "ep_success_buffer": None,
        "_n_updates": int(n_updates),
        "observation_space": _serialized("<class 'gymnasium.spaces.box.Box'>", pickle_box(obs_low, obs_high), dtype="float32",
                                         _shape=str((len(obs_low),)), low=str(list(obs_low)), high=str(list(obs_high))),
        "action_space": _serialized("<class 'gymnasium.spaces.box.Box'>", pickle_box(act_low, act_high), dtype="float32",
                                    _shape=str((len(act_low),)), low=str(list(act_low)), high=str(list(act_high))),
        "n_envs": int(n_envs),
        "n_steps": int(cfg.n_steps),
        "gamma": float(cfg.gamma),
        "gae_lambda": float(cfg.gae_lambda),
        "ent_coef": float(cfg.ent_coef),
        "vf_coef": float(cfg.vf_coef),
        "max_grad_norm": float(cfg.max_grad_norm),
        "rollout_buffer_class": None,
        "rollout_buffer_kwargs": {},
        "batch_size": int(batch_size),
        "n_epochs": int(cfg.n_epochs),
        "clip_range": float(cfg.clip_range),
        "clip_range_vf": None,
        "normalize_advantage": bool(cfg.normalize_advantage),
        "target_kl": None,
        # not SB3's: lets this package restore its own batching exactly (SB3 ignores unknown attributes: they just land in __dict__)
        "brb_n_minibatches": int(cfg.n_minibatches),
        "brb_adam_eps": float(cfg.adam_eps),
    }
    return json.dumps(data, indent=4)


# ---------------------------------------------------------------------------------------------- data (read)
def parse_data(text: str) -> Dict[str, Any]:
    """`data` of a zip written by SB3 or by build_data -> the plain values; serialised entries (classes, spaces, schedules that
    SB3 cloudpickles) are dropped — nothing here needs them."""
    raw = json.loads(text)
    out = {}
    for k, v in raw.items():
        if isinstance(v, dict) and ":serialized:" in v:
            continue
        out[k] = v
    return out


def ppo_config_fields(data: Dict[str, Any], n_envs_now: Optional[int] = None) -> Dict[str, Any]:
    """hyper-parameters of a parsed `data` entry as PPOConfig keyword arguments (legacy zips of this package keep them under
    "hyper_parameters")."""
    if "hyper_parameters" in data:
        return dict(data["hyper_parameters"])
    f: Dict[str, Any] = {}
    for k in ("n_steps", "n_epochs", "gamma", "gae_lambda", "ent_coef", "vf_coef", "max_grad_norm", "normalize_advantage", "seed"):
        if data.get(k) is not None:
            f[k] = data[k]
    for k in ("learning_rate", "clip_range"):                 # SB3 stores a pickled schedule when these were callables
        if isinstance(data.get(k), (int, float)):
            f[k] = float(data[k])
    if "brb_n_minibatches" in data:
        f["n_minibatches"] = int(data["brb_n_minibatches"])
    elif data.get("batch_size") and data.get("n_steps") and data.get("n_envs"):
        f["n_minibatches"] = max(1, int(data["n_steps"]) * int(data["n_envs"]) // int(data["batch_size"]))
    if "brb_adam_eps" in data:
        f["adam_eps"] = float(data["brb_adam_eps"])
    return f


SYSTEM_INFO = ("- OS: Linux\n- Stable-Baselines3: " + SB3_VERSION + " (layout written by balance_robot_b200, no SB3 installed)\n"
               "- PyTorch: see policy.pth\n- GPU Enabled: True\n")
