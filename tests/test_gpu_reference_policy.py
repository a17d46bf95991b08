"""-m gpu: "trained-policy episode returns must agree statistically" (BASELINE.json north_star) with the reference's OWN policy.

The CUDA path is driven closed-loop by the int8 policy extracted from the reference's RobotMovePolicy.tflite
(tests/reference_policy.py) and compared with golden rollouts of the fp64 oracle under the same policy, seed and robots
(tests/golden/policy_rollouts.npz, made by tests/golden/make_policy_rollouts.py).  Trajectories of a closed loop with a
quantised controller decorrelate after the first rounding flip, so the comparison is on episode outcomes and reward
statistics.  Nothing here reads /root/reference.
"""
import pathlib

import numpy as np
import pytest
import torch

from balance_robot_b200 import make_vec
from reference_policy import RobotMovePolicy

pytestmark = pytest.mark.gpu
GOLD = np.load(pathlib.Path(__file__).parent / "golden" / "policy_rollouts.npz")
SEED = int(GOLD["seed"])


def drive(env_id, n, steps):
    """first-episode length / truncation flag per robot, per-step mean reward over the first `m` robots (m = golden size)."""
    m = len(GOLD[f"{env_id}_first_len"])
    env = make_vec(env_id, n, device="cuda:0", seed=SEED)
    pol = RobotMovePolicy("cuda:0")
    obs = env.reset()
    first_len = torch.zeros(n, dtype=torch.int32, device="cuda:0")
    first_trunc = torch.zeros(n, dtype=torch.uint8, device="cuda:0")
    mean_rew, mean_rew_all = [], []
    for _ in range(steps):
        obs, rew, done, info = env.step(pol.act(obs))
        ended = done.bool() & (first_len == 0)
        first_len = torch.where(ended, info.episode_length, first_len)
        first_trunc = torch.where(ended, info.truncated, first_trunc)
        mean_rew.append(rew[:m].double().mean()); mean_rew_all.append(rew.double().mean())
    st = env.stats()
    env.close()
    assert st["nonconverged"] == 0, st
    if env_id == "Env03-v2":
        # poses with a wheel within reach of the block (no wheel-block contact is generated: DESIGN.md 3b) are counted, not hidden
        print(f"Env03-v2 under the reference policy: {st['unsupported']} of {st['env_steps']} env-steps in an unsupported pose")
        assert st["unsupported"] < 0.02 * st["env_steps"], st
    else:
        assert st["unsupported"] == 0, st
    return first_len.cpu().numpy(), first_trunc.cpu().numpy(), torch.stack(mean_rew).cpu().numpy(), torch.stack(mean_rew_all).cpu().numpy()


def binom_tol(p, n1, n2, z=4.0):
    p = min(max(p, 0.02), 0.98)
    return z * np.sqrt(p * (1 - p) * (1.0 / n1 + 1.0 / n2))


def test_env01_v1_policy_balances_every_robot():
    steps = int(GOLD["Env01-v1_steps"])
    fl, _, mr, mr_all = drive("Env01-v1", 4096, steps)
    assert (GOLD["Env01-v1_first_len"] == 0).all()                 # the oracle robots never fell ...
    assert (fl == 0).all()                                          # ... and neither does any of the 4,096 here
    g = GOLD["Env01-v1_mean_reward"]
    assert g[300:].mean() > 0.95 and mr_all[300:].mean() > 0.95
    assert abs(mr[300:].mean() - g[300:].mean()) < 0.01, (mr[300:].mean(), g[300:].mean())     # same robots as the golden run


def test_env01_v2_same_robots_start_fallen_and_the_rest_recover():
    """v2 resets up to 1 rad of pitch (Q3): 13 % start beyond the 50 degree limit, the policy has to catch the others."""
    steps = int(GOLD["Env01-v2_steps"])
    gl = GOLD["Env01-v2_first_len"]
    m = len(gl)
    fl, _, mr, _ = drive("Env01-v2", 8192, steps)
    assert np.array_equal(fl[:m] == 1, gl == 1)                    # decided by the reset draws + bit-exact task logic
    p_g, p_d = (gl == 0).mean(), (fl == 0).mean()                  # first episode still running at the end
    assert abs(p_g - p_d) < binom_tol(p_g, m, len(fl)), (p_g, p_d)
    g = GOLD["Env01-v2_mean_reward"]
    assert abs(mr[300:].mean() - g[300:].mean()) < 0.03, (mr[300:].mean(), g[300:].mean())


def test_env01_v3_policy_follows_the_schedule_without_falling():
    steps = int(GOLD["Env01-v3_steps"])
    fl, _, mr, _ = drive("Env01-v3", 4096, steps)
    gl, g = GOLD["Env01-v3_first_len"], GOLD["Env01-v3_mean_reward"]
    assert abs((gl == 0).mean() - (fl == 0).mean()) < binom_tol((gl == 0).mean(), len(gl), len(fl))
    for a, b in ((200, 600), (600, 900), (900, 1100), (1100, steps)):      # one window per commanded speed
        assert abs(mr[a:b].mean() - g[a:b].mean()) < 0.03, (a, b, mr[a:b].mean(), g[a:b].mean())


def test_env03_v2_survival_under_block_impacts_matches_oracle():
    steps = int(GOLD["Env03-v2_steps"])
    fl, ft, mr, _ = drive("Env03-v2", 4096, steps)
    gl, gt = GOLD["Env03-v2_first_len"], GOLD["Env03-v2_first_trunc"]
    p_g = ((gl == 0) | (gt == 1)).mean()                            # reached the 1,200-step TimeLimit without being knocked over
    p_d = ((fl == 0) | (ft == 1)).mean()
    assert 0.3 < p_g < 0.98
    assert abs(p_g - p_d) < binom_tol(p_g, len(gl), len(fl)), (p_g, p_d)
    g = GOLD["Env03-v2_mean_reward"]
    assert abs(mr[100:].mean() - g[100:].mean()) < 0.15, (mr[100:].mean(), g[100:].mean())
