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


def make(n, seed, emu_cls=None, wheel_block=False):
    spec = mjcf.parse("scene_env03.xml")
    rm = model.compile_model(spec, 3, 1200, wheel_block=wheel_block)
    env = (emu_cls or helpers.EmuVecEnv)(rm, n, seed=seed)
    flags = ref.FLAG_ACTDERIV_SKIP_CLAMPED | ref.FLAG_RPY_FROM_FIRST_ROW | (ref.FLAG_CYLINDER_BOX if wheel_block else 0)
    rv = ref.RefVecEnv(spec, "Env03-v2", n, 1200, nthreads=8, flags=flags)
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


# ------------------------------------------------------------------------------------------------ wheel-block contacts
def _cyl_box(c, a, R, L, b, E, h, margin):
    dp = C.POINTER(C.c_double)
    f = ref.lib().brb_ref_cylinder_box
    f.argtypes = [dp, dp, C.c_double, C.c_double, dp, dp, dp, C.c_double, dp, dp, dp]
    c, a, b, E, h = (np.ascontiguousarray(x, np.float64) for x in (c, a, b, E, h))
    d, n, p = C.c_double(), np.zeros(3), np.zeros(3)
    hit = f(c.ctypes.data_as(dp), a.ctypes.data_as(dp), R, L, b.ctypes.data_as(dp), E.ctypes.data_as(dp), h.ctypes.data_as(dp), margin,
            C.byref(d), n.ctypes.data_as(dp), p.ctypes.data_as(dp))
    return hit, d.value, n, p


def test_cylinder_box_collider_known_configurations():
    """The own analytic collider (oracle/brb_ref.c, restated in fp32 in brb_env03.cuh): distances of configurations with a closed form,
    normal from the cylinder to the box, nothing beyond the margin, and sep <= the true distance on random poses (it is a maximum over
    a finite set of directions of a lower bound)."""
    R, L, h, I3 = 0.034, 0.013, [0.02] * 3, np.eye(3)
    hit, d, n, p = _cyl_box([0, 0, 0], [1, 0, 0], R, L, [0, R + 0.02 + 0.001, 0], I3, h, 0.002)        # face against the tread
    assert hit and abs(d - 0.001) < 1e-12 and np.allclose(n, [0, 1, 0]) and np.allclose(p, [0, R + 0.0005, 0], atol=1e-12)
    hit, d, n, p = _cyl_box([0, 0, 0], [1, 0, 0], R, L, [L + 0.02 - 0.0005, 0.01, 0], I3, h, 0.002)    # face against the cap, 0.5 mm deep
    assert hit and abs(d + 0.0005) < 1e-12 and np.allclose(n, [1, 0, 0]) and abs(p[0] - (L - 0.00025)) < 1e-12
    assert not _cyl_box([0, 0, 0], [1, 0, 0], R, L, [0, R + 0.02 + 0.0021, 0], I3, h, 0.002)[0]         # just outside the margin
    c45 = np.array([[np.cos(np.pi / 4), np.sin(np.pi / 4), 0], [-np.sin(np.pi / 4), np.cos(np.pi / 4), 0], [0, 0, 1]])
    hit, d, n, p = _cyl_box([0, 0, 0], [0, 0, 1], R, L, [R + 0.02 * np.sqrt(2) + 0.001, 0, 0], c45, h, 0.002)   # vertical edge against the tread
    assert hit and abs(d - 0.001) < 1e-9 and np.allclose(n, [1, 0, 0], atol=1e-9)
    # vertex against the tread: body diagonal along +y
    from scipy.spatial.transform import Rotation as Rot
    v = np.array([1.0, 1.0, 1.0]) / np.sqrt(3)
    rot = Rot.align_vectors([[0, -1, 0]], [v])[0].as_matrix()          # box axes: rows of rot.T map local -> world
    hit, d, n, p = _cyl_box([0, 0, 0], [1, 0, 0], R, L, [0, R + 0.02 * np.sqrt(3) + 0.0015, 0], rot.T, h, 0.002)
    assert hit and abs(d - 0.0015) < 1e-9 and np.allclose(n, [0, 1, 0], atol=1e-6) and np.allclose(p, [0, R + 0.00075, 0], atol=1e-6)
    # random poses near contact: the result is a lower bound of the true signed distance (a maximum over a finite set of directions);
    # separated shapes: exact (compared with a dense sampling of directions); 1.5 mm deep: within 1 mm of it
    rng = np.random.default_rng(0)
    dirs = rng.normal(size=(40000, 3)); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    a = np.array([1.0, 0, 0])
    sampled = lambda delta, E, m=40000: ((dirs[:m] @ delta) - L * np.abs(dirs[:m] @ a) - R * np.sqrt(np.maximum(0, 1 - (dirs[:m] @ a) ** 2))
                                         - (np.abs(dirs[:m] @ E.T) * 0.02).sum(1)).max()
    for trial in range(24):
        E = Rot.random(random_state=int(rng.integers(1 << 30))).as_matrix()
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        target = 0.001 if trial % 2 == 0 else -0.0015
        lo, hi = 0.0, 0.2
        for _ in range(40):
            mid = 0.5 * (lo + hi)
            if sampled(u * mid, E, 3000) > target: hi = mid
            else: lo = mid
        delta = u * hi
        brute = sampled(delta, E)
        hit, d, n, p = _cyl_box([0, 0, 0], a, R, L, delta, E, h, 1.0)
        assert hit and abs(np.linalg.norm(n) - 1) < 1e-12 and n @ delta > 0
        if target > 0:
            assert brute - 1e-5 <= d <= brute + 6e-4, (d, brute)      # exact to ~10 um for separated shapes (the random sampling of directions resolves ~0.5 mm)
        else:
            assert brute - 1e-3 <= d <= brute + 6e-4, (d, brute)


def wheel_block_scenario(env, rv, n, seed, steps):
    """Every env starts with its block 0.8 mm from a wheel (cap, tread, edge-on, random orientation) and closing at 0.4 m/s; device state
    re-synchronised to the oracle before every step.  Returns robot / block errors per (env, step) and the number of wheel-block contacts."""
    from scipy.spatial.transform import Rotation as Rot
    _, ur = ref.env03_draws(seed, 0, n, 0)
    rv.reset(ur)
    env.reset()
    m = rv.model
    rng = np.random.default_rng(seed)
    for k in range(n):
        d = rv.env(k).d
        ref.lib().brb_ref_forward(C.byref(m), C.byref(d))
        g = 1 + (k % 2)                                            # l_wheel_geom / r_wheel_geom
        cw = np.array(d.geom_xpos[g][:]); Rw = np.array(d.geom_xmat[g][:]).reshape(3, 3); ax = Rw[:, 2]
        out = ax if ax @ (cw - np.array(d.qpos[0:3])) > 0 else -ax
        direction = [out, np.array([0.0, 1.0, 0.0]), np.array([0.0, -1.0, 0.0]), np.array([0.0, 0.6, 0.8])][(k // 2) % 4]
        if (k // 2) % 4:
            direction = direction - (direction @ ax) * ax
            direction /= np.linalg.norm(direction)
        rot = [Rot.identity(), Rot.from_euler("z", 45, degrees=True), Rot.random(random_state=int(rng.integers(1 << 30)))][(k // 8) % 3]
        E = rot.as_matrix().T                                      # rows = box axes
        lo, hi = 0.0, 0.2
        for _ in range(60):                                        # distance of the block centre along `direction` that leaves a 0.8 mm gap
            mid = 0.5 * (lo + hi)
            hit, dist, _, _ = _cyl_box(cw, ax, m.geom_size[g][0], m.geom_size[g][1], cw + mid * direction, E, list(m.geom_size[4][:]), 10.0)
            if dist > 0.0008: hi = mid
            else: lo = mid
        bp = cw + hi * direction
        qx, qy, qz, qw = rot.as_quat()
        for j, val in enumerate(list(bp) + [qw, qx, qy, qz]): d.qpos[9 + j] = float(val)
        for j, val in enumerate(list(-0.4 * direction) + [0.0, 0.0, 0.0]): d.qvel[8 + j] = float(val)
    er, eb, contacts = [], [], 0
    for t in range(1, steps + 1):
        act = np.zeros((n, 2), np.float32)
        q0, v0 = rv.get_state()
        env.set_state(q0, v0)
        for k in range(n): ref.lib().brb_ref_forward(C.byref(m), C.byref(rv.env(k).d))
        us, ur = ref.env03_draws(seed, 0, n, t)
        obs, rew, done, trunc = rv.step(act, us, ur)
        o_dev, r_dev, d_dev, _ = env.step(act)
        q1, v1 = rv.get_state()
        qd, vd, _ = env.get_state()
        live = ~done.astype(bool) & ~d_dev.astype(bool)
        contacts += sum(any(rv.env(k).d.contact[i].pair in (3, 5) for i in range(rv.env(k).d.ncon)) for k in range(n))
        if live.any():
            eq, ev = helpers.state_errors(qd[live][:, :9], vd[live][:, :8], q1[live][:, :9], v1[live][:, :8])
            er.append(np.maximum(eq, ev))
            bq = np.abs(qd[live][:, 9:] - q1[live][:, 9:]).max(1) / np.maximum(1.0, np.abs(q1[live][:, 9:]).max(1))
            bv = np.abs(vd[live][:, 8:] - v1[live][:, 8:]).max(1) / np.maximum(1.0, np.abs(v1[live][:, 8:]).max(1))
            eb.append(np.maximum(bq, bv))
    return np.concatenate(er), np.concatenate(eb), contacts


def test_wheel_block_contacts_match_oracle():
    n, seed = 24, 17
    rm, env, rv = make(n, seed, wheel_block=True)
    er, eb, contacts = wheel_block_scenario(env, rv, n, seed, 8)
    assert contacts >= 5 * n                                   # every block reached its wheel and stayed in contact for steps
    # Typical agreement is 3e-7 (robot) / 2e-6 (block).  The tails are configurations in which the collider itself is discontinuous --
    # a block face flat against the cap or the tread, two candidate directions with equal separation -- where fp32 and fp64 pick
    # differently for a substep; the adversarial set-up here puts a third of the envs exactly there.
    assert np.median(er) < 2e-6 and np.quantile(er, 0.80) < 1e-5 and (er >= 1e-3).mean() <= 0.04, (np.quantile(er, [0.5, 0.8, 0.95]), er.max(), contacts)
    assert np.median(eb) < 5e-6 and np.quantile(eb, 0.75) < 1e-5 and (eb >= 1e-3).mean() <= 0.07, (np.quantile(eb, [0.5, 0.75, 0.95]), eb.max(), contacts)
    assert env.stats()[3] == 0 and env.stats()[4] == 0        # nothing non-converged, nothing unsupported
    env.close(); rv.close()
