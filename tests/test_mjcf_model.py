"""Host model compiler: MJCF subset -> ModelSpec -> constant block, against SURVEY.md A.1 hand-derived numbers,
against the oracle's independent mj_setConst restatement, and (in this container) against the reference XML."""
import ctypes as C
import pathlib

import numpy as np
import pytest

from balance_robot_b200 import mjcf, model, registry
from oracle import ref

REF_ENVS = pathlib.Path("/root/reference/src/balance_robot/envs")


@pytest.fixture(scope="module")
def spec():
    return mjcf.parse("scene_env01.xml")


def test_sizes_and_masses(spec):
    assert (spec.nq, spec.nv, spec.nu) == (9, 8, 2)
    assert [b.name for b in spec.bodies] == ["world", "robot_body", "l_wheel", "r_wheel"]
    # Q8: inertiafromgeom ignores <inertial>; density 1000
    assert spec.bodies[1].mass == pytest.approx(0.63270, abs=1e-5)
    assert spec.bodies[2].mass == pytest.approx(0.0944237, abs=1e-7)
    chassis_I = np.diag(spec.bodies[1].inertia)
    np.testing.assert_allclose(chassis_I, [1.6139e-3, 2.0690e-3, 5.9943e-4], rtol=2e-4)
    wheel_I = spec.bodies[2].inertia
    assert wheel_I[0, 0] == pytest.approx(5.45769e-5, rel=1e-5)      # axial = body x after the geom quat
    assert wheel_I[1, 1] == pytest.approx(3.26077e-5, rel=1e-5)
    assert spec.timestep == 2e-5 and spec.integrator == "implicitfast"


def test_pairs(spec):
    names = [(spec.geoms[p.geom1].name, spec.geoms[p.geom2].name, p.explicit) for p in spec.pairs]
    assert ("floor", "l_wheel_geom", True) in names and ("floor", "r_wheel_geom", True) in names
    assert ("floor", "chassis_geom", False) in names           # dynamic pair, default parameters
    p = spec.pairs[0]
    assert p.friction[:3] == (0.9, 0.9, 0.1) and p.solref == (0.02, 0.5) and p.solimp[:3] == (0.5, 0.5, 0.002)
    dropped = {(a, b): why for a, b, why in spec.dropped_pairs}
    assert "parent-child" in dropped[("chassis_geom", "l_wheel_geom")]


def test_env03_scene_mixing():
    s = mjcf.parse("scene_env03.xml")
    assert (s.nq, s.nv) == (16, 14)
    assert s.bodies[s.body_id("block")].mass == pytest.approx(0.064)       # Q8: 0.2 in <inertial> is ignored
    floor_block = [p for p in s.pairs if {s.geoms[p.geom1].name, s.geoms[p.geom2].name} == {"floor", "block_geom"}][0]
    assert floor_block.solref == pytest.approx((0.0125, 0.95)) and floor_block.margin == 0.002   # A.6 mixing
    floor_wheel = [p for p in s.pairs if {s.geoms[p.geom1].name, s.geoms[p.geom2].name} == {"floor", "l_wheel_geom"}][0]
    assert floor_wheel.friction[0] == 1.0 and not floor_wheel.explicit                            # Q9


def test_constant_block_matches_survey_and_oracle(spec):
    rm = model.compile_model(spec, 1, 6000)
    m = ref.model_from_spec(spec)
    # SURVEY.md A.1
    assert rm.mass == pytest.approx(0.8215474, abs=1e-7)
    assert rm.com_z == pytest.approx(0.084444, abs=1e-6)
    assert rm.meaninertia == pytest.approx(0.310538, abs=1e-6)
    assert rm.contact["D"] == pytest.approx(0.1010276, rel=1e-6)
    assert rm.contact["R"] == pytest.approx(9.89828, rel=1e-6)
    assert rm.contact["K"] == pytest.approx(40000.0) and rm.contact["B"] == pytest.approx(200.0)
    # oracle (independent C implementation of mj_setConst)
    assert rm.meaninertia == pytest.approx(m.meaninertia, rel=1e-13)
    iw = ref.arr(m.body_invweight0, 4)
    assert rm.invweight0["robot_body"][0] == pytest.approx(iw[1, 0], rel=1e-12)
    assert rm.invweight0["l_wheel"][0] == pytest.approx(iw[2, 0], rel=1e-12)
    assert rm.invweight0["l_wheel"][1] == pytest.approx(iw[2, 1], rel=1e-12)
    # M_b == oracle M(qpos0)
    d = ref.new_data(m)
    ref.lib().brb_ref_forward(C.byref(m), C.byref(d))
    M = ref.arr(d.qM).reshape(ref.MAXNV, ref.MAXNV)[:8, :8]
    np.testing.assert_allclose(rm.M_b, M, atol=1e-15)
    np.testing.assert_allclose(rm.M_b_inv @ rm.M_b, np.eye(8), atol=1e-9)


def test_time_table_is_sequential_sum(spec):
    rm = model.compile_model(spec, 1, 6000)
    t = 0.0
    for _ in range(500):
        t += 2e-5
    assert rm.time_table[2] == t and rm.time_table[0] == 0.0
    assert rm.time_table[200] > 1.0 and rm.time_table[199] <= 1.0     # Q10: first schedule switch after 200 steps
    assert len(rm.time_table) == 6002


def test_unsupported_models_fail_loudly():
    with pytest.raises(model.UnsupportedModel):
        model.compile_model(mjcf.parse("scene_env03.xml"), 0, 1200)       # the block scene only compiles as Env03-v2
    with pytest.raises(model.UnsupportedModel):
        model.compile_model(mjcf.parse("scene_env01.xml"), 3, 1200)
    with pytest.raises(NotImplementedError):
        registry.spec("Env03-v1")
    with pytest.raises(KeyError):
        registry.spec("Nope-v0")


def test_registry_matches_reference_ids():
    assert registry.spec("Env01-v2").max_episode_steps == 6000
    assert {k: v.kind for k, v in registry.REGISTRY.items()} == {"Env01-v1": 0, "Env01-v2": 1, "Env01-v3": 2, "Env03-v2": 3}
    assert registry.spec("Env03-v2").max_episode_steps == 1200


@pytest.mark.skipif(not REF_ENVS.exists(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("ours,theirs", [("scene_env01.xml", "env01_v1.xml"), ("scene_env03.xml", "env03_v1.xml")])
def test_assets_equal_reference_mjcf(ours, theirs):
    a, b = mjcf.parse(ours), mjcf.parse(REF_ENVS / theirs)
    assert (a.nq, a.nv, a.nu, a.timestep, a.gravity) == (b.nq, b.nv, b.nu, b.timestep, b.gravity)
    for x, y in zip(a.bodies, b.bodies):
        assert (x.parent, x.pos, x.quat, x.mass, x.ipos) == (y.parent, y.pos, y.quat, y.mass, y.ipos)
        if x.inertia is not None:
            np.testing.assert_array_equal(x.inertia, y.inertia)
    for x, y in zip(a.joints, b.joints):
        assert (x.name, x.type, x.axis, x.damping) == (y.name, y.type, y.axis, y.damping)
    for x, y in zip(a.geoms, b.geoms):
        assert (x.type, x.body, x.size, x.pos, x.quat, x.margin, x.solref) == (y.type, y.body, y.size, y.pos, y.quat, y.margin, y.solref)
    for x, y in zip(a.actuators, b.actuators):
        assert (x.name, x.joint, x.kv, x.ctrlrange, x.forcerange) == (y.name, y.joint, y.kv, y.ctrlrange, y.forcerange)
    key = lambda s, p: (s.geoms[p.geom1].body, s.geoms[p.geom2].body, p.friction, p.solref, p.solimp, p.margin, p.explicit)
    assert sorted(key(a, p) for p in a.pairs) == sorted(key(b, p) for p in b.pairs)
