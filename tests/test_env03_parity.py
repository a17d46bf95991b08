"""Env03-v2 (robot + fired block): the CUDA path's arithmetic (host emulation here, the device in test_gpu_env03.py)
against the fp64 oracle — wheel-floor contacts with position-dependent impedance, block-floor plane-box contacts,
chassis-block box-box impacts through the coupled 14-dof solve, and the remove / delay / re-fire state machine."""
import ctypes as C

import numpy as np
import pytest

import helpers
import parity_checks as pc
from balance_robot_b200 import mjcf, model
from oracle import ref


def make(n, seed, emu_cls=None):
    spec = mjcf.parse("scene_env03.xml")
    rm = model.compile_model(spec, 3, 1200)
    env = (emu_cls or helpers.EmuVecEnv)(rm, n, seed=seed)
    rv = ref.RefVecEnv(spec, "Env03-v2", n, 1200, nthreads=8)
    rv.set_attack_side(ref.env03_attack_side(seed, 0, n))
    return rm, env, rv


def env03_single_step(env, rv, n, seed, steps):
    """Device state re-synchronised to the oracle before every step (set_state); returns per-(env, step) errors for the
    robot (qpos abs, qvel rel) and the block, plus how many steps saw a chassis-block impact."""
    _, ur = ref.env03_draws(seed, 0, n, 0)
    obs = rv.reset(ur)
    o_dev = env.reset()
    pc.assert_f32_equal(o_dev, obs, 1)
    rng = np.random.default_rng(seed)
    er, eb, impacts = [], [], 0
    chk = dict(steps=0, done_mismatch=0, max_rew_err=0.0, max_obs_err=0.0, refired=0, parked=0)
    env03_single_step.last = chk
    for t in range(1, steps + 1):
        act = (helpers.pd_policy(obs) + 0.1 * rng.uniform(-1, 1, (n, 2))).astype(np.float32)
        q0, v0 = rv.get_state()
        env.set_state(q0, v0)
        for k in range(n):      # MujocoEnv.set_state on the oracle side too (mj_forward: fresh kinematics), so the pre-step reward sees the same state
            ref.lib().brb_ref_forward(C.byref(rv.model), C.byref(rv.env(k).d))
        us, ur = ref.env03_draws(seed, 0, n, t)
        obs, rew, done, trunc = rv.step(act, us, ur)
        o_dev, r_dev, d_dev, _ = env.step(act)
        q1, v1 = rv.get_state()
        qd, vd, _ = env.get_state()
        live = ~done.astype(bool) & ~d_dev.astype(bool)
        # task logic on the device against the oracle: reward (pre-step state, identical by construction), termination,
        # observation, and the remove / delay / re-fire decisions (block parked at (10, 10) or back in play)
        chk["steps"] += n
        chk["done_mismatch"] += int((done.astype(bool) != d_dev.astype(bool)).sum())
        chk["max_rew_err"] = max(chk["max_rew_err"], float(np.abs(rew - r_dev).max()))
        if live.any():
            # obs[1] is a finite difference of the pitch over 5 ms: a 1e-6 state difference shows up as 2e-4 there
            chk["max_obs_err"] = max(chk["max_obs_err"], float(np.abs(obs[live] - o_dev[live]).max()))
            park_o = (np.abs(q1[live][:, 9] - 10) < 0.5) & (np.abs(q1[live][:, 10] - 10) < 0.5)
            park_d = (np.abs(qd[live][:, 9] - 10) < 0.5) & (np.abs(qd[live][:, 10] - 10) < 0.5)
            was_parked = (np.abs(q0[live][:, 9] - 10) < 0.5) & (np.abs(q0[live][:, 10] - 10) < 0.5)
            chk["park_mismatch"] = chk.get("park_mismatch", 0) + int((park_o != park_d).sum())
            chk["parked"] += int((park_o & ~was_parked).sum())
            chk["refired"] += int((~park_o & was_parked).sum())
        impacts += sum(any(rv.env(k).d.contact[i].pair == 1 for i in range(rv.env(k).d.ncon)) for k in range(n))
        if live.any():
            eq, ev = helpers.state_errors(qd[live][:, :9], vd[live][:, :8], q1[live][:, :9], v1[live][:, :8])
            er.append(np.maximum(eq, ev))
            # block: positions can be ~10 m (parked at (10,10)), speeds 7.5 m/s: relative to max(1, |x|)
            bq = np.abs(qd[live][:, 9:] - q1[live][:, 9:]).max(1) / np.maximum(1.0, np.abs(q1[live][:, 9:]).max(1))
            bv = np.abs(vd[live][:, 8:] - v1[live][:, 8:]).max(1) / np.maximum(1.0, np.abs(v1[live][:, 8:]).max(1))
            eb.append(np.maximum(bq, bv))
    return np.concatenate(er), np.concatenate(eb), impacts


def test_reset_state_and_observation_match_oracle():
    rm, env, rv = make(16, 3)
    _, ur = ref.env03_draws(3, 0, 16, 0)
    assert np.array_equal(rv.reset(ur), env.reset())
    q, v = rv.get_state()
    qd, vd, _ = env.get_state()
    np.testing.assert_allclose(qd, q, atol=1e-14)
    np.testing.assert_allclose(vd, v, atol=1e-13)
    assert np.allclose(np.linalg.norm(v[:, 8:11], axis=1), 7.5)              # block fired at 7.5 m/s (env03_v2.py:49)
    d = np.linalg.norm(q[:, 9:11] - q[:, 0:2], axis=1)
    assert np.allclose(d, 0.3, atol=1e-9)                                    # from 0.3 m in front of / behind the robot
    env.close(); rv.close()


def test_single_step_parity_through_impacts():
    n, seed = 8, 5
    rm, env, rv = make(n, seed)
    er, eb, impacts = env03_single_step(env, rv, n, seed, 45)
    assert impacts > 20                                        # the block hit the chassis in many (env, step) pairs
    assert np.quantile(er, 0.99) < 1e-5 and (er >= 1e-5).mean() <= 0.02, (np.quantile(er, 0.99), er.max())
    assert np.quantile(eb, 0.95) < 1e-5 and (eb >= 3e-5).mean() <= 0.03, (np.quantile(eb, 0.95), eb.max())
    st = env.stats()
    assert st[3] == 0                                          # no active-set iteration cap hits
    check_task_outputs(env03_single_step.last)
    env.close(); rv.close()


def check_task_outputs(chk):
    """reward / done / obs and the block state machine of the device path against the oracle's (same pre-step state)."""
    assert chk["max_rew_err"] == 0.0, chk                      # f32 of the same fp64 expression on the same state: bit-equal
    assert chk["done_mismatch"] <= 0.002 * chk["steps"], chk   # a contact-timing outlier next to the 50 degree line
    assert chk["park_mismatch"] <= 0.01 * chk["steps"], chk    # "block slower than 0.1 m/s" is a threshold on an fp32 block state
    assert chk["max_obs_err"] < 0.05, chk


def test_block_cycle_remove_delay_refire():
    """Free-running CUDA-path arithmetic: the block comes to rest, is parked at (10, 10), and is re-fired 0.5 s later
    (env03_v1.py:39-49); episodes end by the 50 degree pitch test or the 1200-step limit."""
    n, seed = 6, 11
    rm, env, rv = make(n, seed)
    obs = env.reset()
    parked = fired = dones = 0
    was_parked = np.zeros(n, bool)
    for t in range(400):
        obs, rew, done, trunc = env.step(helpers.pd_policy(obs))
        q, v, _ = env.get_state()
        is_parked = (np.abs(q[:, 9] - 10) < 0.5) & (np.abs(q[:, 10] - 10) < 0.5)
        parked += int((is_parked & ~was_parked).sum())
        fired += int((~is_parked & was_parked & ~done.astype(bool)).sum())
        was_parked = is_parked & ~done.astype(bool)
        dones += int(done.sum())
        assert np.isfinite(q).all() and np.isfinite(v).all() and np.isfinite(obs).all()
    assert parked > 0 and fired > 0
    env.close(); rv.close()


def test_time_limit_is_1200_steps():
    from balance_robot_b200 import registry
    assert registry.spec("Env03-v2").max_episode_steps == 1200 and registry.spec("Env03-v2").kind == 3
