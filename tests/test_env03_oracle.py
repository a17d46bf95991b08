"""Env03-v2 on the oracle side: MJCF compile of the block scene, box-box collider
sanity, and the C env-level oracle against the Python transliteration of envs/env03_v1.py + envs/env03_v2.py."""
import ctypes as C

import numpy as np
import pytest

import helpers
from balance_robot_b200 import mjcf
from oracle import ref
from pyref_env03 import PyRefEnv03


@pytest.fixture(scope="module")
def ctx():
    spec = mjcf.parse("scene_env03.xml")
    return spec, ref.model_from_spec(spec), ref.lib()


def test_block_constants(ctx):
    spec, m, L = ctx
    assert m.nq == 16 and m.nv == 14
    iw = ref.arr(m.body_invweight0, 5)
    assert iw[4, 0] == pytest.approx(1 / 0.064) and iw[4, 1] == pytest.approx(1 / (0.064 * 2 * 0.02 ** 2 / 3))


def test_block_rests_on_floor_inside_its_margin(ctx):
    _, m, L = ctx
    d = ref.new_data(m)
    d.qpos[9:12] = [1.0, 1.0, 0.05]
    L.brb_ref_step(C.byref(m), C.byref(d), 250 * 60)
    block = [d.contact[i] for i in range(d.ncon) if d.contact[i].pair == 6]
    assert len(block) == 4 and all(0 < c.dist < 0.002 for c in block)          # margin 0.002: soft contact before touching
    f = sum(d.efc_force[c.efc_address + r] for c in block for r in range(4))
    assert f == pytest.approx(0.064 * 9.81, rel=1e-4)


def test_block_hits_chassis_and_transfers_momentum(ctx):
    _, m, L = ctx
    d = ref.new_data(m)
    d.qpos[9:12] = [10, 10, 0]
    L.brb_ref_step(C.byref(m), C.byref(d), 250 * 60)
    d.qpos[9:12] = [0.005, 0.3, 0.14]
    d.qvel[8:11] = [0, -7.5, 0]
    hit = False
    for _ in range(12):
        L.brb_ref_step(C.byref(m), C.byref(d), 250)
        hit |= any(d.contact[i].pair == 1 for i in range(d.ncon))
    assert hit
    assert d.qvel[9] > -1.0                 # block stopped / bounced
    assert abs(d.qvel[3]) > 3.0             # chassis got a pitch-rate kick
    assert np.isfinite(ref.arr(d.qpos, 16)).all()


def test_env03_v2_oracle_equals_python_transliteration(ctx):
    spec, m, L = ctx
    n, steps, seed = 3, 260, 31
    rv = ref.RefVecEnv(spec, "Env03-v2", n, 1200, nthreads=3)
    side = ref.env03_attack_side(seed, 0, n)
    rv.set_attack_side(side)
    py = [PyRefEnv03(m, bool(side[k])) for k in range(n)]
    us, ur = ref.env03_draws(seed, 0, n, 0)
    obs = rv.reset(ur)
    for k in range(n):
        np.testing.assert_array_equal(py[k].reset(ur[k][:24]), obs[k])
        C.memmove(C.byref(py[k].d), C.byref(rv.env(k).d), C.sizeof(ref.RefData))
    rng = np.random.default_rng(3)
    fired = removed = dones = 0
    for t in range(1, steps + 1):
        act = (helpers.pd_policy(obs) + 0.2 * rng.uniform(-1, 1, (n, 2))).astype(np.float32)
        us, ur = ref.env03_draws(seed, 0, n, t)
        timers_before = [rv.env(k).has_block_timer for k in range(n)]
        obs, rew, done, trunc = rv.step(act, us, ur)
        for k in range(n):
            ob, r, term = py[k].step(act[k], us[k][:5])
            assert np.float32(r) == pytest.approx(rew[k], rel=1e-6, abs=1e-7)
            if done[k]:
                dones += 1
                assert term or trunc[k]
                np.testing.assert_array_equal(py[k].reset(ur[k][:24]), obs[k])
                C.memmove(C.byref(py[k].d), C.byref(rv.env(k).d), C.sizeof(ref.RefData))
                continue
            np.testing.assert_allclose(ob, obs[k], rtol=1e-6, atol=1e-7)
            assert not term
            e = rv.env(k)
            np.testing.assert_allclose(ref.arr(py[k].d.qpos, 16), ref.arr(e.d.qpos, 16), rtol=1e-10, atol=1e-12)
            # block launch velocity: np.linalg.norm (BLAS dot) vs sqrt(x*x+y*y+z*z) may differ in the last ulp
            np.testing.assert_allclose(ref.arr(py[k].d.qvel, 14), ref.arr(e.d.qvel, 14), rtol=1e-10, atol=1e-12)
            removed += (not timers_before[k]) and bool(e.has_block_timer)
            fired += bool(timers_before[k]) and not e.has_block_timer
            # re-synchronise (a 1-ulp launch-velocity difference is amplified by later collisions): single-step comparison
            C.memmove(C.byref(py[k].d), C.byref(e.d), C.sizeof(ref.RefData))
            py[k].block_delay_time_start = e.block_delay_time_start if e.has_block_timer else None
    assert removed > 0 and fired > 0          # the remove -> 0.5 s delay -> re-fire cycle happened
    rv.close()
