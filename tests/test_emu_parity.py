"""CPU-side parity of the CUDA kernel's arithmetic: balance_robot_b200/csrc/brb_kernels.cu's per-env functions compiled
for the host (tests/host_emu, -DBRB_HOST_EMU) against the fp64 oracle.  This exercises exactly the fp32 formulas,
compensated sums, active-set solver and fp64 task logic the GPU runs (up to FMA contraction), so formula bugs show up
here without a GPU; the -m gpu tests repeat the same procedures on the device through the C-ABI."""
import numpy as np
import pytest

import helpers
import parity_checks as pc
from balance_robot_b200 import mjcf, model


@pytest.fixture(scope="module")
def spec():
    return mjcf.parse("scene_env01.xml")


@pytest.mark.parametrize("kind", [0, 1])
def test_single_step_parity_closed_loop(spec, kind):
    rm = model.compile_model(spec, kind, 6000)
    env = helpers.EmuVecEnv(rm, 8, seed=3)
    wq, wv = pc.single_step_parity(env, spec, helpers.ENV_IDS[kind], 8, 3, 150)
    assert wq < 1e-6 and wv < 1e-6          # two orders inside the 1e-5 bar
    env.close()


def test_single_step_parity_random_actions_with_slip(spec):
    rm = model.compile_model(spec, 1, 6000)
    env = helpers.EmuVecEnv(rm, 16, seed=8)
    wq, wv = pc.single_step_parity(env, spec, "Env01-v2", 16, 8, 60, policy="random", max_outlier_frac=0.0)     # 16 x 60: no event in this sample
    assert wq < 1e-5 and wv < 1e-5, (wq, wv)
    st = env.stats()
    assert st[3] == 0, "active-set iteration hit its cap"
    env.close()


def test_contact_timing_outliers_are_rare(spec):
    """Larger sample of tumbling robots: with the fp64 on/off predicate the contact-timing outliers drop from ~2e-3 of
    env-steps (fp32 predicate) to ~4e-5; the bound leaves room for the sample size."""
    rm = model.compile_model(spec, 1, 6000)
    env = helpers.EmuVecEnv(rm, 128, seed=21)
    q99, outliers = pc.single_step_parity(env, spec, "Env01-v2", 128, 21, 80, policy="random", max_outlier_frac=3e-4)
    assert q99 < 1e-6, q99
    env.close()


def test_mirrored_free_run_with_auto_reset(spec):
    """configs[1] in miniature: Philox on, auto-reset on, no re-synchronisation; every env mirrored on the oracle."""
    rm = model.compile_model(spec, 1, 6000)
    n = 48
    env = helpers.EmuVecEnv(rm, n, seed=0)
    rng = np.random.default_rng(1234)
    acts = {}
    res = pc.mirrored_free_run(env, spec, "Env01-v2", n, 0, 120, lambda t: acts.setdefault(t, rng.uniform(-1, 1, (n, 2)).astype(np.float32)))
    assert res["both_done"] > 20 and res["done_mismatch"] <= 1
    assert res["compared"] > 0.9 * res["total"] and res["desync_events"] <= 0.004 * res["total"] + 2, res
    assert res["max_rew_err"] < 1e-5
    env.close()


def test_free_run_horizon(spec):
    rm = model.compile_model(spec, 0, 6000)
    env = helpers.EmuVecEnv(rm, 8, seed=4)
    h = pc.free_run_horizon(env, spec, "Env01-v1", 8, 4, 200)
    assert h >= 60, h                        # agreement until chaotic divergence (~15 steps for naive fp32)
    env.close()


@pytest.mark.parametrize("kind,steps", [(0, 40), (1, 60), (2, 230)])
def test_task_logic_bit_exact(spec, kind, steps):
    rm = model.compile_model(spec, kind, 6000)
    n = 6
    env = helpers.EmuVecEnv(rm, n, seed=13)
    checked, dones = pc.task_logic_bit_exact(env, rm.time_table, helpers.ENV_IDS[kind], n, 13, steps)
    assert checked > 0.5 * n * steps
    if kind == 1:
        assert dones > 0                     # v2 starts 12.8 % of episodes beyond the termination angle (Q3)
    env.close()


def test_replay_draws_override_philox(spec):
    rm = model.compile_model(spec, 1, 6000)
    n = 4
    a = helpers.EmuVecEnv(rm, n, seed=1)
    b = helpers.EmuVecEnv(rm, n, seed=999)          # different Philox key, same replayed draws
    rng = np.random.default_rng(0)
    ur = rng.random((n, 16))
    assert np.array_equal(a.reset(ur), b.reset(ur))
    for _ in range(5):
        u = rng.random((n, 20))
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        ra, rb = a.step(act, u), b.step(act, u)
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
    a.close(); b.close()


def test_truncation_at_time_limit(spec):
    rm = model.compile_model(spec, 0, 3)             # TimeLimit of 3 steps
    env = helpers.EmuVecEnv(rm, 2, seed=2)
    env.reset()
    for t in range(1, 4):
        obs, rew, done, trunc = env.step(np.zeros((2, 2), np.float32))
        assert done.all() == (t == 3) and trunc.all() == (t == 3)
    assert (env.epl == 3).all() and (obs[:, 1] == 0).all()
    env.close()
