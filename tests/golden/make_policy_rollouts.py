"""Golden closed-loop rollouts of the fp64 ORACLE under the reference's trained policy (tests/reference_policy.py).

For every env: N robots seeded by the shared Philox stream (seed, env id), driven for STEPS steps by the int8 policy
extracted from the reference's RobotMovePolicy.tflite.  Stored per env: first-episode length and how it ended for every
robot, per-step mean reward, per-step count of live first episodes.  The GPU tests compare the CUDA path's statistics for
the same robots with these (tests/test_gpu_reference_policy.py); oracle-only, /root/reference is not needed.

Run:  python tests/golden/make_policy_rollouts.py        (Env03-v2 takes ~15 min on 8 cores)
"""
import os
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from balance_robot_b200 import mjcf          # noqa: E402
from oracle import ref                       # noqa: E402
from reference_policy import RobotMovePolicy  # noqa: E402

SEED = 5
CASES = {"Env01-v1": (256, 1000), "Env01-v2": (512, 600), "Env01-v3": (256, 1300), "Env03-v2": (256, 1200)}


def rollout(env_id, n, steps, nthreads):
    e3 = env_id == "Env03-v2"
    rv = ref.RefVecEnv(mjcf.parse("scene_env03.xml" if e3 else "scene_env01.xml"), env_id, n, 1200 if e3 else 6000, nthreads=nthreads)
    draws = (lambda k: ref.env03_draws(SEED, 0, n, k)) if e3 else (lambda k: ref.philox_draws(SEED, 0, n, k))
    if e3:
        rv.set_attack_side(ref.env03_attack_side(SEED, 0, n))
    pol = RobotMovePolicy()
    obs = rv.reset(draws(0)[1])
    first_len = np.zeros(n, np.int32); first_trunc = np.zeros(n, np.uint8)
    mean_rew = np.zeros(steps, np.float64); live = np.zeros(steps, np.int32)
    for k in range(1, steps + 1):
        us, ur = draws(k)
        obs, r, d, tr = rv.step(pol.act(torch.from_numpy(obs)).numpy(), us, ur)
        mean_rew[k - 1] = r.astype(np.float64).mean()
        ended = d.astype(bool) & (first_len == 0)
        first_len[ended] = rv.ep_len[ended]; first_trunc[ended] = tr[ended]
        live[k - 1] = int((first_len == 0).sum())
    rv.close()
    return first_len, first_trunc, mean_rew, live


def main():
    out = {"seed": np.int64(SEED)}
    for env_id, (n, steps) in CASES.items():
        fl, ft, mr, lv = rollout(env_id, n, steps, os.cpu_count() or 1)
        print(env_id, "robots", n, "steps", steps, "first episodes ended", int((fl > 0).sum()), "mean reward", mr.mean(), flush=True)
        out.update({f"{env_id}_first_len": fl, f"{env_id}_first_trunc": ft, f"{env_id}_mean_reward": mr, f"{env_id}_live": lv,
                    f"{env_id}_steps": np.int64(steps)})
    np.savez_compressed(pathlib.Path(__file__).resolve().parent / "policy_rollouts.npz", **out)


if __name__ == "__main__":
    main()
