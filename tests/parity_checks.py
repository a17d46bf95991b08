"""Parity procedures shared by the CPU (host emulation of the kernel arithmetic) and GPU test files.

`env` is anything with reset() -> obs, step(actions) -> (obs, reward, done, truncated), get_state() -> (qpos, qvel,
xquat), set_state(qpos, qvel), all numpy: tests/helpers.EmuVecEnv on the CPU, GpuAdapter (test_gpu_parity.py) on the B200.
"""
from __future__ import annotations

import numpy as np

import helpers
from oracle import ref
from pyref_env import PyRefEnv, reference_order_reset_draws

QPOS0 = np.array([0, 0, 0, 1, 0, 0, 0, 0, 0.0])
TOL = 1e-5          # BASELINE.json north_star: single-step qpos/qvel within 1e-5 relative


def single_step_parity(env, spec, env_id, n, seed, steps, policy="pd", noise=0.3, tol=TOL, max_outlier_frac=0.0):
    """Oracle runs a closed-loop trajectory; before every step the device state is set to the oracle's, both take
    the same action, post-step qpos/qvel must agree within `tol`.  Returns the worst errors seen.

    max_outlier_frac: MuJoCo's contact model switches a contact on at dist < 0 exactly, and a sliding contact carries
    O(10 N) of damping force from its first substep, so a touch-down / lift-off that lands within fp32 resolution
    (~5e-9 m) of zero can fire one substep apart in fp32 and fp64 and move a velocity by ~1e-3 (measured).
    Tumbling robots under random actions hit that about once per 500 env-steps; balanced ones essentially never."""
    rv = ref.RefVecEnv(spec, env_id, n, 6000, nthreads=8)
    _, ur = ref.philox_draws(seed, 0, n, 0)
    obs = rv.reset(ur)
    env.reset()
    rng = np.random.default_rng(seed)
    worst_q = worst_v = 0.0
    contact_steps = 0
    all_err = []
    for t in range(1, steps + 1):
        if policy == "pd":
            act = (helpers.pd_policy(obs) + noise * rng.uniform(-1, 1, (n, 2))).astype(np.float32)
        else:
            act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        q0, v0 = rv.get_state()
        env.set_state(q0, v0)
        us, ur = ref.philox_draws(seed, 0, n, t)
        obs, rew, done, trunc = rv.step(act, us, ur)
        env.step(act)
        live = ~done.astype(bool)                      # envs that were auto-reset have unrelated states now
        q1, v1 = rv.get_state()
        qd, vd, _ = env.get_state()
        # the device may have reset on its own (noisy termination draws differ by construction): compare the others
        same = live & (np.abs(qd[:, 2] - q1[:, 2]) < 1e-2)
        eq, ev = helpers.state_errors(qd[same], vd[same], q1[same], v1[same])
        if eq.size:
            worst_q, worst_v = max(worst_q, eq.max()), max(worst_v, ev.max())
            all_err.append(np.maximum(eq, ev))
        contact_steps += int(sum(rv.env(k).d.nefc > 0 for k in range(n)))
    rv.close()
    all_err = np.concatenate(all_err)
    outliers = float((all_err >= tol).mean())
    assert outliers <= max_outlier_frac, (outliers, worst_q, worst_v)
    assert np.median(all_err) < tol / 30, np.median(all_err)
    assert contact_steps > 0.3 * steps * n, "trajectory did not exercise the contact solver"
    if max_outlier_frac == 0.0:
        return worst_q, worst_v
    return float(np.quantile(all_err, 0.99)), outliers


def free_run_horizon(env, spec, env_id, n, seed, steps, tol=TOL):
    """Identical initial states, action sequence and replayed draws, no re-synchronisation: first step at which any
    env's error exceeds `tol` (chaotic divergence: the open-loop plant e-folds every ~32 steps)."""
    rv = ref.RefVecEnv(spec, env_id, n, 6000, nthreads=8)
    _, ur = ref.philox_draws(seed, 0, n, 0)
    obs = rv.reset(ur)
    o_dev = env.reset()
    assert np.array_equal(obs, o_dev), "reset observations differ"
    rng = np.random.default_rng(seed)
    horizon = steps
    alive = np.ones(n, bool)
    for t in range(1, steps + 1):
        act = (helpers.pd_policy(obs) + 0.05 * rng.uniform(-1, 1, (n, 2))).astype(np.float32)
        us, ur = ref.philox_draws(seed, 0, n, t)
        obs, rew, done, trunc = rv.step(act, us, ur)
        o_dev, r_dev, d_dev, t_dev = env.step(act)
        alive &= ~(done.astype(bool) | d_dev.astype(bool))
        q1, v1 = rv.get_state()
        qd, vd, _ = env.get_state()
        eq, ev = helpers.state_errors(qd[alive], vd[alive], q1[alive], v1[alive])
        if eq.size and (eq.max() > tol or ev.max() > tol):
            horizon = t
            break
    rv.close()
    return horizon


def task_logic_bit_exact(env, time_table, env_id, n, seed, steps, max_episode_steps=6000):
    """Reward / termination / truncation / observation / reset of the device path against the pure-Python
    transliteration of the reference env code, evaluated on the DEVICE's own states with the same Philox draws:
    must be bit-exact at the float32 outputs ("bit-exact given identical state")."""
    py = [PyRefEnv(env_id) for _ in range(n)]
    elapsed = np.zeros(n, int)
    _, ur = ref.philox_draws(seed, 0, n, 0)
    obs = env.reset()
    qpos, qvel, xquat = env.get_state()

    def sync(k):
        py[k].sim.xquat, py[k].sim.qvel, py[k].sim.time = xquat[k].copy(), qvel[k].copy(), time_table[elapsed[k]]

    def check_reset(k, u_row):
        py[k].draws = reference_order_reset_draws(env_id, u_row)
        qp = py[k].reset_draw_qpos(QPOS0)
        np.testing.assert_allclose(qp, qpos[k], atol=4e-16)
        elapsed[k] = 0
        sync(k)
        ob = py[k]._get_obs()
        np.testing.assert_array_equal(ob, obs[k])

    for k in range(n):
        check_reset(k, ur[k])
    rng = np.random.default_rng(seed + 1)
    n_done = n_checked = 0
    for t in range(1, steps + 1):
        act = (helpers.pd_policy(obs) + 0.4 * rng.uniform(-1, 1, (n, 2))).astype(np.float32)
        us, ur = ref.philox_draws(seed, 0, n, t)
        expect_r = np.zeros(n, np.float32)
        for k in range(n):
            py[k].draws = [us[k, 0]]
            sync(k)
            expect_r[k] = np.float32(py[k].pre_step(act[k])[0])
        obs, rew, done, trunc = env.step(act)
        qpos, qvel, xquat = env.get_state()
        np.testing.assert_array_equal(rew, expect_r)
        for k in range(n):
            if done[k]:
                n_done += 1
                # terminal observation and the termination decision are checked from env.tobs when available
                check_reset(k, ur[k])
                continue
            elapsed[k] += 1
            py[k].draws = [us[k, 1], us[k, 2], us[k, 3]]
            sync(k)
            ob, term = py[k].post_step()
            assert not term, "device kept running an env the reference logic terminates"
            np.testing.assert_array_equal(ob, obs[k])
            assert bool(trunc[k]) is False
            n_checked += 1
    return n_checked, n_done
