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

    Both sides draw the same Philox noise (same seed, env ids and event index), so they must also agree on which envs
    finished: the `done` flags are compared exactly and every env that neither side reset is compared (no filter on the
    state itself — a gross error cannot hide).

    max_outlier_frac: MuJoCo's contact model switches a contact on at dist < 0 exactly, and a sliding contact carries
    O(10 N) of damping force from its first substep, so a touch-down / lift-off firing one substep apart moves a velocity
    by ~1e-3.  The kernel decides that predicate in fp64 whenever the fp32 distance is within 2e-7 m of zero
    (rim_dist_fp64), which leaves only events where the fp64 distance itself lies within the in-step trajectory
    difference (~3e-10 m, from fp32 force arithmetic) of zero: measured 4e-5 of tumbling env-steps."""
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
        o_dev, r_dev, d_dev, t_dev = env.step(act)
        assert np.array_equal(done.astype(bool), d_dev.astype(bool)), (t, np.nonzero(done.astype(bool) != d_dev.astype(bool))[0])
        live = ~done.astype(bool)                      # envs that were auto-reset have unrelated states now
        q1, v1 = rv.get_state()
        qd, vd, _ = env.get_state()
        eq, ev = helpers.state_errors(qd[live], vd[live], q1[live], v1[live])
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


def mirrored_free_run(env, spec, env_id, n_mirror, seed, steps, actions_of, tol=TOL, env0=0):
    """BASELINE.json configs[1] as written: the device runs its own Philox streams (no replay, no re-synchronisation,
    auto-reset on) on a shard of any size; envs [0, n_mirror) are mirrored on the oracle with the same initial state,
    actions and draws (ref.philox_draws).  `actions_of(t)` -> float32 [n, 2] for the WHOLE shard.

    An env is compared while it is "in sync": from a reset both sides took on the same step until its error first
    exceeds `tol` (chaotic divergence / a contact-timing event) — it re-enters the comparison when both sides reset it
    on the same step again (the reset state is a function of the draws only).  Returns a dict of counts."""
    rv = ref.RefVecEnv(spec, env_id, n_mirror, 6000, nthreads=8)
    _, ur = ref.philox_draws(seed, env0, n_mirror, 0)
    obs = rv.reset(ur)
    o_dev = env.reset()
    # f32 of the same fp64 expression; CUDA's and glibc's sincos / atan2 differ in the last fp64 bit now and then, and the 24-bit
    # Philox uniforms put reset observations on f32 rounding ties often enough for that bit to show: equal to 1 f32 ulp
    np.testing.assert_allclose(o_dev[:n_mirror], obs, rtol=2.5e-7, atol=0, err_msg="reset observations differ")
    in_sync = np.ones(n_mirror, bool)
    since = np.zeros(n_mirror, int)
    out = dict(compared=0, total=0, desync_events=0, early_desync=0, done_mismatch=0, horizons=[], max_rew_err=0.0, both_done=0)
    for t in range(1, steps + 1):
        act = np.ascontiguousarray(actions_of(t), np.float32)
        us, ur = ref.philox_draws(seed, env0, n_mirror, t)
        obs, rew, done, trunc = rv.step(act[:n_mirror], us, ur)
        o_dev, r_dev, d_dev, t_dev = env.step(act)
        done, d_dev = done.astype(bool), d_dev[:n_mirror].astype(bool)
        q1, v1 = rv.get_state()
        qd, vd, _ = env.get_state()
        eq, ev = helpers.state_errors(qd[:n_mirror], vd[:n_mirror], q1, v1)
        err = np.maximum(eq, ev)
        out["total"] += n_mirror
        since += 1
        for k in range(n_mirror):
            if in_sync[k]:
                out["compared"] += 1
                if done[k] != d_dev[k]:
                    out["done_mismatch"] += 1
                    in_sync[k] = False
                elif done[k]:
                    out["both_done"] += 1
                    assert err[k] < 1e-13, ("reset states differ", t, k, err[k], qd[k], q1[k], vd[k], v1[k])          # same draws -> same reset state
                    np.testing.assert_allclose(o_dev[k], obs[k], rtol=2.5e-7, atol=0)   # sincos differs in the last fp64 bit between libm and CUDA
                    since[k] = 0
                elif err[k] >= tol:
                    out["desync_events"] += 1
                    out["horizons"].append(int(since[k]))
                    out["early_desync"] += int(since[k] < 50)
                    in_sync[k] = False
                else:
                    out["max_rew_err"] = max(out["max_rew_err"], float(abs(rew[k] - r_dev[k])))
                    assert np.abs(obs[k] - o_dev[k]).max() < 1e-2, "observation far off while the state agrees"   # obs[1] is a finite difference over 5 ms
            elif done[k] and d_dev[k]:
                in_sync[k] = True
                since[k] = 0
    rv.close()
    return out


def free_run_horizon(env, spec, env_id, n, seed, steps, tol=TOL):
    """Identical initial states, action sequence and replayed draws, no re-synchronisation: first step at which any
    env's error exceeds `tol` (chaotic divergence: the open-loop plant e-folds every ~32 steps)."""
    rv = ref.RefVecEnv(spec, env_id, n, 6000, nthreads=8)
    _, ur = ref.philox_draws(seed, 0, n, 0)
    obs = rv.reset(ur)
    o_dev = env.reset()
    assert_f32_equal(o_dev, obs, 1, "reset observations differ")
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


def assert_f32_equal(actual, desired, ulps=0, err_msg=""):
    """float32 outputs equal to `ulps` units in the last place (0 = bit-exact).  The host emulation shares glibc's atan2 / sincos with
    the reference's numpy / scipy and is held to 0; CUDA's fp64 atan2 / sincos are not correctly rounded (<= 2 ulp), and right after
    a reset the inputs are 24-bit Philox uniforms (pitch = (u - 0.5) 2.0 + (u' - 0.5) 0.05: ~30 significant bits), which puts the
    fp64 -> f32 rounding on a tie often enough for the last fp64 bit to show: the device is held to 1 ulp of the f32 output."""
    a, d = np.ascontiguousarray(actual, np.float32), np.ascontiguousarray(desired, np.float32)
    if ulps == 0:
        np.testing.assert_array_equal(a, d, err_msg=err_msg)
        return
    ia, id_ = a.view(np.int32).astype(np.int64), d.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia); id_ = np.where(id_ < 0, -(id_ & 0x7FFFFFFF), id_)
    assert np.abs(ia - id_).max(initial=0) <= ulps, (err_msg, a, d)


def task_logic_bit_exact(env, time_table, env_id, n, seed, steps, max_episode_steps=6000, ulps=0):
    """Reward / termination / truncation / observation / reset of the device path against the pure-Python
    transliteration of the reference env code, evaluated on the DEVICE's own states with the same Philox draws:
    must be bit-exact at the float32 outputs ("bit-exact given identical state"; `ulps`: see assert_f32_equal)."""
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
        assert_f32_equal(obs[k], ob, ulps)

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
        assert_f32_equal(rew, expect_r, ulps)
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
            assert_f32_equal(obs[k], ob, ulps)
            assert bool(trunc[k]) is False
            n_checked += 1
    return n_checked, n_done


# ---------------------------------------------------------------------------------------------------------------------
# fixtures produced by the UNMODIFIED reference env classes (tests/golden/make_reference_fixtures.py via tests/ref_shim)
import pathlib

REFCLS = sorted((pathlib.Path(__file__).parent / "golden").glob("refcls_*.npz"))


def replay_reference_class_fixture(env, path, tol=TOL, min_horizon=50, ulps=0):
    """`env` (host emulation or the device through the C-ABI, Philox seeded like the fixture) against a fixture recorded
    from the reference's own env class.

    resync fixture: before every step the env is put into the fixture's pre-step state with set_state (fresh kinematics on
    both sides), so the REWARD must be bit-equal to the reference class's; termination must agree; post-step state within
    `tol`; observation within the physics tolerance (obs[1] is a finite difference over 5 ms).
    free fixture: no re-synchronisation — the env must track the recorded trajectory (state within tol,
    same done flags, reward within 5e-4) for as long as an env stays in sync (chaotic divergence ends that after >= 50 steps).
    Returns counters."""
    g = np.load(path)
    resync = bool(g["resync"])
    n, steps = g["obs0"].shape[0], g["obs"].shape[0]
    nq = g["qpos0"].shape[1]
    obs = env.reset()
    np.testing.assert_allclose(obs, g["obs0"], rtol=2.5e-7, atol=0)          # sincos last-bit differences only
    qd, vd, _ = env.get_state()
    np.testing.assert_allclose(qd, g["qpos0"], atol=1e-14)
    np.testing.assert_allclose(vd, g["qvel0"], atol=1e-13)
    q_prev, v_prev = g["qpos0"], g["qvel0"]
    out = dict(rewards_bit_equal=0, compared=0, done_events=0, max_state_err=0.0, max_obs_err=0.0, refires=0, sync_steps=0, errs=[])
    in_sync = np.ones(n, bool)
    since = np.zeros(n, int)
    for t in range(steps):
        if resync:
            env.set_state(q_prev, v_prev)
        o, r, d, tr = env.step(g["actions"][t])
        d, gd = d.astype(bool), g["done"][t].astype(bool)
        qd, vd, _ = env.get_state()
        if resync:
            assert_f32_equal(r, g["reward"][t].astype(np.float32), ulps, err_msg=f"step {t}")     # BIT-equal to the reference class (device: 1 ulp)
            out["rewards_bit_equal"] += n
            assert np.array_equal(d, gd), (t, d, gd)
            assert np.array_equal(tr.astype(bool), g["truncated"][t].astype(bool))
            live = ~gd
            # robot dofs within the north-star tolerance; the Env03-v2 block (plain fp32 state, 7.5 m/s impacts) is covered by
            # tests/test_env03_parity.py with its own statistical bound
            eq, ev = helpers.state_errors(qd[live][:, :9], vd[live][:, :8], g["qpos"][t][live][:, :9], g["qvel"][t][live][:, :8])
            if eq.size:
                out["max_state_err"] = max(out["max_state_err"], float(max(eq.max(), ev.max())))
                out["errs"] += np.maximum(eq, ev).tolist()
                out["max_obs_err"] = max(out["max_obs_err"], float(np.abs(o[live] - g["obs"][t][live]).max()))
            if nq > 9 and live.any():
                parked_d = (qd[live][:, 9] == 10) & (qd[live][:, 10] == 10)
                parked_g = (g["qpos"][t][live][:, 9] == 10) & (g["qpos"][t][live][:, 10] == 10)
                assert np.array_equal(parked_d, parked_g), (t, "block remove / re-fire decision differs")
                out["refires"] += int(g["refired"][t][live].sum())
            # reset states are a function of the draws only
            if gd.any():
                np.testing.assert_allclose(qd[gd][:, :9], g["qpos"][t][gd][:, :9], atol=1e-14)
                np.testing.assert_allclose(o[gd], g["obs"][t][gd], rtol=2.5e-7, atol=0)
            out["compared"] += int(live.sum())
            out["done_events"] += int(gd.sum())
            q_prev, v_prev = g["qpos"][t], g["qvel"][t]
        else:
            since += 1
            eq, ev = helpers.state_errors(qd[:, :9], vd[:, :8], g["qpos"][t][:, :9], g["qvel"][t][:, :8])
            err = np.maximum(eq, ev)
            for k in range(n):
                if in_sync[k]:
                    if d[k] != gd[k] or (not gd[k] and err[k] >= tol):
                        assert since[k] >= min_horizon or d[k] != gd[k], (t, k, since[k], err[k])   # only chaotic divergence may end a stretch
                        in_sync[k] = False
                        continue
                    out["sync_steps"] += 1
                    assert abs(float(r[k]) - float(g["reward"][t][k])) < max(5e-4, 100 * tol), (t, k, r[k], g["reward"][t][k], err[k], since[k])      # d reward / d pitch = 0.5 dv, dv up to ~80 rad/s
                    if gd[k]:
                        out["done_events"] += 1
                        since[k] = 0
                    else:
                        out["max_state_err"] = max(out["max_state_err"], float(err[k]))
                elif d[k] and gd[k]:
                    in_sync[k], since[k] = True, 0
    return out
