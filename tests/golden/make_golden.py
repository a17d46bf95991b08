"""Generates tests/golden/*.npz from the fp64 oracle (oracle/) — NOT from MuJoCo: the reference's physics engine is
not installable here and the reference ships no fixtures, so these vectors pin regressions of the restatement, not
parity with MuJoCo ("parity unpinned").  Re-run with:  python tests/golden/make_golden.py
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import helpers  # noqa: E402
from balance_robot_b200 import mjcf  # noqa: E402
from oracle import ref  # noqa: E402

OUT = pathlib.Path(__file__).resolve().parent


def make(env_id, n, steps, seed):
    spec = mjcf.parse("scene_env01.xml")
    rv = ref.RefVecEnv(spec, env_id, n, 6000, nthreads=8)
    _, ur = ref.philox_draws(seed, 0, n, 0)
    obs = rv.reset(ur)
    rng = np.random.default_rng(seed)
    rec = dict(obs0=obs.copy(), actions=[], obs=[], reward=[], done=[], qpos=[], qvel=[])
    q, v = rv.get_state()
    rec["qpos0"], rec["qvel0"] = q, v
    for t in range(1, steps + 1):
        act = (helpers.pd_policy(obs) + 0.2 * rng.uniform(-1, 1, (n, 2))).astype(np.float32)
        us, ur = ref.philox_draws(seed, 0, n, t)
        obs, rew, done, trunc = rv.step(act, us, ur)
        q, v = rv.get_state()
        for k, x in zip(("actions", "obs", "reward", "done", "qpos", "qvel"), (act, obs, rew, done, q, v)):
            rec[k].append(x.copy())
    rv.close()
    out = {k: np.stack(v) if isinstance(v, list) else v for k, v in rec.items()}
    out["seed"], out["env_id"] = np.int64(seed), np.array(env_id)
    np.savez_compressed(OUT / f"{env_id}_n{n}_s{steps}.npz", **out)
    print(env_id, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    make("Env01-v1", 8, 30, 101)
    make("Env01-v2", 8, 30, 102)
    make("Env01-v3", 8, 30, 103)
