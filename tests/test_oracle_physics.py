"""Pins for the fp64 oracle's mj_step restatement that need no MuJoCo:
closed-form trajectories, conservation laws, the static contact equilibrium and the solver's KKT conditions."""
import ctypes as C

import numpy as np
import pytest

from balance_robot_b200 import mjcf, model
from oracle import ref


@pytest.fixture(scope="module")
def ctx():
    spec = mjcf.parse("scene_env01.xml")
    m = ref.model_from_spec(spec)
    return spec, m, ref.lib()


def rand_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def test_free_fall_is_exact_for_semi_implicit_stepping(ctx):
    _, m, L = ctx
    d = ref.new_data(m)
    n = 2500                                  # 50 ms: still above the floor (Q11: contact after ~64 ms)
    L.brb_ref_step(C.byref(m), C.byref(d), n)
    h, g = 2e-5, 9.81
    assert d.nefc == 0
    assert d.qpos[2] == pytest.approx(-0.5 * g * h * h * n * (n + 1), rel=1e-12)
    assert d.qvel[2] == pytest.approx(-g * h * n, rel=1e-12)
    assert d.time == pytest.approx(n * h, rel=1e-12)


def test_first_contact_time_matches_geometry(ctx):
    _, m, L = ctx
    d = ref.new_data(m)
    steps = 0
    while d.nefc == 0 and steps < 5000:
        L.brb_ref_step(C.byref(m), C.byref(d), 1)
        steps += 1
    # wheel bottoms start 0.02 above the floor: t = sqrt(2*0.02/9.81) = 63.9 ms -> ~3193 substeps
    assert abs(steps - np.sqrt(2 * 0.02 / 9.81) / 2e-5) < 3


def _flight(h, n):
    spec = mjcf.parse("scene_env01.xml")
    spec.timestep = h
    m = ref.model_from_spec(spec)
    L = ref.lib()
    rng = np.random.default_rng(1)
    d = ref.new_data(m)
    d.qpos[2] = 1.0                                        # high above the floor
    d.qpos[3:7] = rand_quat(rng)
    d.qvel[0:8] = [0.3, -0.2, 0.1, 2.0, -1.0, 3.0, 20.0, -10.0]
    d.ctrl[0] = d.ctrl[1] = 1e3                            # clipped to 78.54 -> servo saturated at +0.65 N m all along

    def momentum():
        L.brb_ref_kinematics(C.byref(m), C.byref(d))
        L.brb_ref_mass_matrix(C.byref(m), C.byref(d))
        M = ref.arr(d.qM).reshape(16, 16)[:8, :8]
        return (M @ ref.arr(d.qvel, 8))[:3]                # generalized momentum of the world-frame linear dofs

    e0 = L.brb_ref_energy(C.byref(m), C.byref(d), None, None)
    p0 = momentum().copy()
    work = 0.0
    for _ in range(n):
        w0 = ref.arr(d.qvel, 8)[6:8].copy()
        L.brb_ref_step(C.byref(m), C.byref(d), 1)
        wm = 0.5 * (w0 + ref.arr(d.qvel, 8)[6:8])
        assert abs(d.actuator_force[0]) == 0.65 and abs(d.actuator_force[1]) == 0.65
        work += float(((0.65 - 0.01 * wm) * wm).sum()) * h  # motor + hinge damping power on the relative wheel speed
    e1 = L.brb_ref_energy(C.byref(m), C.byref(d), None, None)
    q = ref.arr(d.qpos, 9)[3:7]
    mt = sum(m.body_mass[b] for b in range(4))
    return momentum() - p0, (e1 - e0 - work) / abs(e0), abs(np.linalg.norm(q) - 1), mt


def test_momentum_and_energy_in_flight():
    """Linear momentum and the energy balance hold up to the integrator's FIRST-ORDER error: halving h halves the
    residual (a wrong bias / gyroscopic term would leave an h-independent one)."""
    T = 0.004                                              # wheels reach ~65 rad/s, still below the 78.54 ctrl clip
    dp1, de1, qn1, mt = _flight(2e-5, 200)
    dp2, de2, qn2, _ = _flight(1e-5, 400)
    assert dp1[2] == pytest.approx(-mt * 9.81 * T, rel=1e-3)
    for k in range(2):
        assert abs(dp1[k]) < 3e-6
        assert 1.7 < dp1[k] / dp2[k] < 2.3
    assert abs(de1) < 1e-4 and 1.7 < de1 / de2 < 2.3
    assert qn1 < 1e-13 and qn2 < 1e-13


def test_mass_matrix_equals_body_frame_closed_form_at_random_orientation(ctx):
    spec, m, L = ctx
    rm = model.compile_model(spec, 0, 6000)
    rng = np.random.default_rng(2)
    for _ in range(5):
        d = ref.new_data(m)
        d.qpos[0:3] = rng.normal(size=3)
        d.qpos[3:7] = rand_quat(rng)
        d.qpos[7:9] = rng.normal(size=2) * 3
        L.brb_ref_kinematics(C.byref(m), C.byref(d))
        L.brb_ref_mass_matrix(C.byref(m), C.byref(d))
        M = ref.arr(d.qM).reshape(16, 16)[:8, :8]
        R = ref.arr(d.xmat)[1].reshape(3, 3)
        T = np.eye(8)
        T[:3, :3] = R.T
        np.testing.assert_allclose(M, T.T @ rm.M_b @ T, atol=1e-15)


def test_bias_equals_gyrostat_closed_form(ctx):
    """The CUDA kernel's chassis-frame bias against the oracle's projected Newton-Euler recursion."""
    spec, m, L = ctx
    rm = model.compile_model(spec, 0, 6000)
    rng = np.random.default_rng(3)
    mass, cz, (Ixx, Iyy, Izz), Ia, g = rm.mass, rm.com_z, rm.inertia_origin, rm.wheel_axial_inertia, 9.81
    for _ in range(5):
        d = ref.new_data(m)
        d.qpos[3:7] = rand_quat(rng)
        d.qvel[0:8] = rng.normal(size=8) * [1, 1, 1, 3, 3, 3, 40, 40]
        L.brb_ref_kinematics(C.byref(m), C.byref(d))
        L.brb_ref_bias(C.byref(m), C.byref(d))
        bias = ref.arr(d.qfrc_bias, 8)
        R = ref.arr(d.xmat)[1].reshape(3, 3)
        n = R[2]                                         # world z in the chassis frame
        w = ref.arr(d.qvel, 8)[3:6]
        sL, sR = d.qvel[6], d.qvel[7]
        c = np.array([0, 0, cz])
        lin_b = mass * np.cross(w, np.cross(w, c)) + mass * g * n
        Lm = np.array([Ixx * w[0] + Ia * (sR - sL), Iyy * w[1], Izz * w[2]])
        ang_b = np.cross(w, Lm) + mass * g * np.cross(c, n)
        expect = np.concatenate([R @ lin_b, ang_b, [0, 0]])
        np.testing.assert_allclose(bias, expect, atol=1e-12)


def test_static_equilibrium_pins_contact_softness(ctx):
    spec, m, L = ctx
    rm = model.compile_model(spec, 0, 6000)
    d = ref.new_data(m)
    L.brb_ref_step(C.byref(m), C.byref(d), 250 * 200)    # 1 s: drops 2 cm and settles
    assert d.ncon == 4 and d.nefc == 16
    forces = ref.arr(d.efc_force, 16)
    assert forces.sum() == pytest.approx(rm.mass * 9.81, rel=1e-6)          # total normal force = m g
    # analytic sink: 16 rows each carry D*K*imp*|dist| (A.1)
    sink = -rm.mass * 9.81 / (16 * rm.contact["D"] * rm.contact["K"] * rm.contact["imp"])
    assert d.contact[0].dist == pytest.approx(sink, rel=1e-5)
    assert sink == pytest.approx(-0.2493e-3, rel=1e-3)
    assert np.allclose(ref.arr(d.efc_R, 16), rm.contact["R"])
    # contact frame for a z-up plane: t1 = +y, t2 = -x (A.6)
    np.testing.assert_allclose(ref.arr(d.contact[0].frame), [0, 0, 1, 0, 1, 0, -1, 0, 0], atol=1e-15)


def test_solver_kkt_conditions_under_slip(ctx):
    _, m, L = ctx
    rng = np.random.default_rng(4)
    d = ref.new_data(m)
    L.brb_ref_step(C.byref(m), C.byref(d), 250 * 30)
    worst = 0.0
    for k in range(60):
        d.ctrl[0] = d.qvel[6] + 4 * rng.uniform(-1, 1)
        d.ctrl[1] = d.qvel[7] + 4 * rng.uniform(-1, 1)
        for _ in range(25):
            L.brb_ref_step(C.byref(m), C.byref(d), 10)
            L.brb_ref_forward(C.byref(m), C.byref(d))
            ne = d.nefc
            if ne == 0:
                continue
            M = ref.arr(d.qM).reshape(16, 16)[:8, :8]
            J = ref.arr(d.efc_J).reshape(ref.MAXEFC, ref.MAXNV)[:ne, :8]
            f = ref.arr(d.efc_force, ne)
            a, a_s = ref.arr(d.qacc, 8), ref.arr(d.qacc_smooth, 8)
            jar = J @ a - ref.arr(d.efc_aref, ne)
            D = ref.arr(d.efc_D, ne)
            np.testing.assert_allclose(f, -D * np.minimum(0, jar), rtol=1e-9, atol=1e-12)     # f_i = -D min(0, jar)
            res = M @ (a - a_s) - J.T @ f
            worst = max(worst, np.abs(res).max())
    assert worst < 1e-10                                                                       # M (a - a_s) = J' f


def test_wheel_spin_up_in_air_matches_linear_ode(ctx):
    """Cal01-like (reference cal01.py:19-20,42): clamped 0.65 N m between wheel and chassis, 0.01 w damping."""
    spec, m, L = ctx
    rm = model.compile_model(spec, 0, 6000)
    d = ref.new_data(m)
    d.qpos[2] = 0.15
    d.ctrl[0] = d.ctrl[1] = 20.0
    n = 60
    L.brb_ref_step(C.byref(m), C.byref(d), n)
    # both wheels get +0.65 about their own hinge axes (-x and +x): net reaction on the chassis about x is zero
    h = 2e-5
    Mi = rm.M_b_inv
    w = np.zeros(8)
    for _ in range(n):
        f = np.zeros(8)
        f[6] = 0.65 - 0.01 * w[6]
        f[7] = 0.65 - 0.01 * w[7]
        Mh = rm.M_b.copy()
        Mh[6, 6] += h * 0.01
        Mh[7, 7] += h * 0.01                       # clamped -> actuator derivative skipped (A.9)
        w = w + h * np.linalg.solve(Mh, f)
    assert abs(4 * (20 - d.qvel[6])) > 0.65        # still saturated at the end of the window
    assert d.qvel[6] == pytest.approx(w[6], rel=1e-9)
    assert d.qvel[3] == pytest.approx(0.0, abs=1e-9)


def test_zero_action_robot_falls_and_pd_controller_balances():
    import helpers
    spec = mjcf.parse("scene_env01.xml")
    n = 8
    for policy, expect_alive in (("zero", False), ("pd", True)):
        rv = ref.RefVecEnv(spec, "Env01-v1", n, 6000, nthreads=8)
        _, ur = ref.philox_draws(5, 0, n, 0)
        o = rv.reset(ur)
        ended = np.zeros(n, bool)
        first_done = np.full(n, 10 ** 9)
        for k in range(1, 301):
            a = np.zeros((n, 2), np.float32) if policy == "zero" else helpers.pd_policy(o)
            us, ur = ref.philox_draws(5, 0, n, k)
            o, r, dn, tr = rv.step(a, us, ur)
            first_done = np.where(dn.astype(bool) & ~ended, k, first_done)
            ended |= dn.astype(bool)
        if expect_alive:
            assert not ended.any()
        else:
            assert ended.all() and first_done.max() <= 220      # falls past 50 deg within ~1 s (8c item 7)
        rv.close()
