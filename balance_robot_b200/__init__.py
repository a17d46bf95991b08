"""balance_robot_b200 — B200-native batched `balance_robot` environments (Env01-v1 / v2 / v3).

`make_vec(id, num_envs)` is the batched analogue of the reference's `gym.make(id)` (sb_rl.py:500):
it returns an object with SB3's VecEnv surface whose step() is one fused sm_100a CUDA launch
(balance_robot_b200/csrc/brb_kernels.cu) behind the C-ABI in include/brb.h.
"""
from .registry import REGISTRY, spec  # noqa: F401
from .vec_env import ACTION_SPACE, OBSERVATION_SPACE, BalanceVecEnv, InfoBatch, make_vec  # noqa: F401

__all__ = ["make_vec", "BalanceVecEnv", "InfoBatch", "REGISTRY", "spec", "OBSERVATION_SPACE", "ACTION_SPACE"]
