"""Environment ids, mirroring the reference's gymnasium registrations (balance_robot/__init__.py:5-66).

Only the ids on the north-star path are buildable; the others raise with the reason."""
from __future__ import annotations

import dataclasses


@dataclasses.dataclass(frozen=True)
class EnvSpec:
    id: str
    kind: int                 # BRB_ENV01_* in include/brb.h
    scene: str                # MJCF under balance_robot_b200/assets
    max_episode_steps: int    # gymnasium TimeLimit
    reward_threshold: float
    entry_point: str          # the reference class this id resolves to


REGISTRY = {
    "Env01-v1": EnvSpec("Env01-v1", 0, "scene_env01.xml", 6000, 6000, "balance_robot.envs.env01_v1:Env01"),
    "Env01-v2": EnvSpec("Env01-v2", 1, "scene_env01.xml", 6000, 6000, "balance_robot.envs.env01_v2:Env01_v2"),
    "Env01-v3": EnvSpec("Env01-v3", 2, "scene_env01.xml", 6000, 6000, "balance_robot.envs.env01_v3:Env01_v3"),
    "Env03-v2": EnvSpec("Env03-v2", 3, "scene_env03.xml", 1200, 6000, "balance_robot.envs.env03_v2:Env03_v2"),
}

NOT_BUILT = {
    "Env03-v1": "out of scope (the north star names Env03-v2; v1 fires the block from random directions)",
    "Env02-v1": "out of scope (not named in the north star)",
    "Env03-v1-fail": "out of scope (mesh drop scene)",
    "Cal01": "out of scope (open-loop calibration run)",
    "EnvMove05-v1": "out of scope (TFLite policy + lidar inside the env)",
}


def spec(env_id: str) -> EnvSpec:
    if env_id in REGISTRY:
        return REGISTRY[env_id]
    if env_id in NOT_BUILT:
        raise NotImplementedError(f"{env_id}: {NOT_BUILT[env_id]}")
    raise KeyError(f"unknown environment id {env_id!r}; known: {sorted(REGISTRY)}")
