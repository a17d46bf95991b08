"""Env03-v2 (robot + fired block): the CUDA path's arithmetic (host emulation here, the device in test_gpu_env03.py)
against the fp64 oracle — wheel-floor contacts with position-dependent impedance, block-floor plane-box contacts,
chassis-block box-box impacts through the coupled 14-dof solve, and the remove / delay / re-fire state machine."""
import numpy as np
import pytest

import helpers
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
    assert np.array_equal(obs, o_dev)
    rng = np.random.default_rng(seed)
    er, eb, impacts = [], [], 0
    for t in range(1, steps + 1):
        act = (helpers.pd_policy(obs) + 0.1 * rng.uniform(-1, 1, (n, 2))).astype(np.float32)
        q0, v0 = rv.get_state()
        env.set_state(q0, v0)
        us, ur = ref.env03_draws(seed, 0, n, t)
        obs, rew, done, trunc = rv.step(act, us, ur)
        o_dev, r_dev, d_dev, _ = env.step(act)
        q1, v1 = rv.get_state()
        qd, vd, _ = env.get_state()
        live = ~done.astype(bool) & ~d_dev.astype(bool)
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
    env.close(); rv.close()


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
