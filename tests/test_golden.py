"""Committed golden trajectories (tests/golden/*.npz, made by tests/golden/make_golden.py from the fp64 oracle):
the oracle must still reproduce them bit-for-bit (regression pin of the restatement), and the kernel arithmetic
(host emulation here, the real device in test_gpu_parity.py) must track them within the 1e-5 bar."""
import pathlib

import numpy as np
import pytest

import helpers
from balance_robot_b200 import mjcf, model
from oracle import ref

GOLD = sorted((pathlib.Path(__file__).parent / "golden").glob("Env*_n*_s*.npz"))


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    env_id, seed = str(g["env_id"]), int(g["seed"])
    n, steps = g["obs0"].shape[0], g["obs"].shape[0]
    rv = ref.RefVecEnv(mjcf.parse("scene_env01.xml"), env_id, n, 6000, nthreads=4)
    _, ur = ref.philox_draws(seed, 0, n, 0)
    assert np.array_equal(rv.reset(ur), g["obs0"])
    for t in range(steps):
        us, ur = ref.philox_draws(seed, 0, n, t + 1)
        obs, rew, done, _ = rv.step(g["actions"][t], us, ur)
        q, v = rv.get_state()
        assert np.array_equal(obs, g["obs"][t]) and np.array_equal(rew, g["reward"][t]) and np.array_equal(done, g["done"][t])
        assert np.array_equal(q, g["qpos"][t]) and np.array_equal(v, g["qvel"][t])
    rv.close()


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_kernel_arithmetic_tracks_golden(path):
    g = np.load(path)
    env_id, seed = str(g["env_id"]), int(g["seed"])
    kind = {v: k for k, v in helpers.ENV_IDS.items()}[env_id]
    n, steps = g["obs0"].shape[0], g["obs"].shape[0]
    env = helpers.EmuVecEnv(model.compile_model(mjcf.parse("scene_env01.xml"), kind, 6000), n, seed=seed)
    assert np.array_equal(env.reset(), g["obs0"])
    alive = np.ones(n, bool)
    for t in range(steps):
        obs, rew, done, _ = env.step(g["actions"][t])
        alive &= ~(done.astype(bool) | g["done"][t].astype(bool))
        qd, vd, _ = env.get_state()
        eq, ev = helpers.state_errors(qd[alive], vd[alive], g["qpos"][t][alive], g["qvel"][t][alive])
        assert eq.size == 0 or (eq.max() < 1e-5 and ev.max() < 1e-5), (t, eq.max(), ev.max())
        np.testing.assert_allclose(rew[alive], g["reward"][t][alive], atol=2e-6)
    assert alive.sum() >= n // 2
    env.close()
