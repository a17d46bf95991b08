"""Task logic pinned to the reference's OWN code: tests/golden/refcls_*.npz were recorded by
running the unmodified classes of /root/reference/src/balance_robot/envs through tests/ref_shim (stand-in mujoco /
gymnasium modules, oracle physics, injected draws; generator: tests/golden/make_reference_fixtures.py).

  * the C oracle's env layer (oracle/brb_ref_env.c) must replay every fixture BIT for bit — observation, reward,
    termination, reset state, block state machine;
  * the pure-Python transliteration (tests/pyref_env.py), which the -m gpu tests evaluate on the device's own states,
    must give bit-equal reward / observation / termination on the fixture's states;
  * the CUDA kernel's arithmetic (host emulation here; the device itself in tests/test_gpu_parity.py) must give the
    reference class's reward bit for bit from the same pre-step state, the same termination decisions, and post-step
    states / observations within the physics tolerance;
  * when /root/reference is present (this container), the fixtures are regenerated live and must equal the committed ones.
"""
import ctypes as C
import pathlib
import sys

import numpy as np
import pytest

import helpers
import parity_checks as pc
from balance_robot_b200 import mjcf, model
from oracle import ref
from pyref_env import PyRefEnv

KIND = {"Env01-v1": 0, "Env01-v2": 1, "Env01-v3": 2, "Env03-v2": 3}
SCENE = {0: "scene_env01.xml", 1: "scene_env01.xml", 2: "scene_env01.xml", 3: "scene_env03.xml"}
IDS = [p.stem for p in pc.REFCLS]


def test_fixtures_are_committed():
    assert len(pc.REFCLS) == 8, "tests/golden/refcls_*.npz missing: run tests/golden/make_reference_fixtures.py where /root/reference exists"


@pytest.mark.parametrize("path", pc.REFCLS, ids=IDS)
def test_oracle_env_layer_replays_reference_class_fixture_bit_for_bit(path):
    g = np.load(path)
    env_id, seed, resync = str(g["env_id"]), int(g["seed"]), bool(g["resync"])
    kind = KIND[env_id]
    n, steps = g["obs0"].shape[0], g["obs"].shape[0]
    spec = mjcf.parse(SCENE[kind])
    rv = ref.RefVecEnv(spec, env_id, n, 1200 if kind == 3 else 6000, nthreads=1)
    if kind == 3:
        rv.set_attack_side(g["attack_side_front"])
        draws = lambda t: ref.env03_draws(seed, 0, n, t)
    else:
        draws = lambda t: ref.philox_draws(seed, 0, n, t)
    obs = rv.reset(draws(0)[1])
    assert np.array_equal(obs, g["obs0"])
    q, v = rv.get_state()
    assert np.array_equal(q, g["qpos0"]) and np.array_equal(v, g["qvel0"])
    for t in range(steps):
        if resync:      # MujocoEnv.set_state(qpos, qvel) with the env's own state = mj_forward
            for k in range(n):
                ref.lib().brb_ref_forward(C.byref(rv.model), C.byref(rv.env(k).d))
        us, ur = draws(t + 1)
        obs, rew, done, trunc = rv.step(g["actions"][t], us, ur)
        assert np.array_equal(obs, g["obs"][t]), t
        assert np.array_equal(rew, g["reward"][t].astype(np.float32)), t
        assert np.array_equal(done, g["done"][t]) and np.array_equal(trunc, g["truncated"][t]), t
        d = done.astype(bool)
        assert np.array_equal(rv.terminal_obs[d], g["terminal_obs"][t][d])
        q, v = rv.get_state()
        assert np.array_equal(q, g["qpos"][t]) and np.array_equal(v, g["qvel"][t]), t
    rv.close()


@pytest.mark.parametrize("path", [p for p in pc.REFCLS if "Env03" not in p.stem and "free" in p.stem], ids=lambda p: p.stem)
def test_python_transliteration_equals_reference_classes_on_fixture_states(path):
    """tests/pyref_env.py is what test_task_logic_bit_exact evaluates on the DEVICE's states; here it is pinned to the
    reference classes: same (stale xquat, qvel, time), same draws -> bit-equal reward, observation, termination."""
    g = np.load(path)
    env_id, seed = str(g["env_id"]), int(g["seed"])
    n, steps = g["obs0"].shape[0], g["obs"].shape[0]
    from pyref_env import reference_order_reset_draws
    py = [PyRefEnv(env_id) for _ in range(n)]
    xq_prev = g["qpos0"][:, 3:7] / np.linalg.norm(g["qpos0"][:, 3:7], axis=1, keepdims=True)      # set_state -> fresh kinematics
    v_prev, t_prev = g["qvel0"], np.zeros(n)
    _, ur = ref.philox_draws(seed, 0, n, 0)
    for k in range(n):
        py[k].draws = reference_order_reset_draws(env_id, ur[k])
        np.testing.assert_allclose(py[k].reset_draw_qpos(pc.QPOS0), g["qpos0"][k], atol=4e-16)
        py[k].sim.xquat, py[k].sim.qvel, py[k].sim.time = xq_prev[k], v_prev[k], 0.0
        assert np.array_equal(py[k]._get_obs(), g["obs0"][k])
    xq_prev = xq_prev.copy()
    for t in range(steps):
        us, ur = ref.philox_draws(seed, 0, n, t + 1)
        for k in range(n):
            e = py[k]
            e.sim.xquat, e.sim.qvel, e.sim.time = xq_prev[k], v_prev[k], t_prev[k]
            e.draws = [us[k, 0]]
            assert np.float32(e.pre_step(g["actions"][t][k])[0]) == np.float32(g["reward"][t][k]), (t, k)
            e.sim.xquat, e.sim.time = g["xquat"][t][k], g["time"][t][k]
            e.draws = [us[k, 1], us[k, 2], us[k, 3]]
            if g["done"][t][k]:
                # the recorded qvel is the reset state's; the terminal observation is checked through the oracle replay above
                e.draws = reference_order_reset_draws(env_id, ur[k])
                e.reset_draw_qpos(pc.QPOS0)
                qn = g["qpos"][t][k][3:7]
                e.sim.xquat, e.sim.qvel, e.sim.time = qn / np.linalg.norm(qn), g["qvel"][t][k], 0.0
                assert np.array_equal(e._get_obs(), g["obs"][t][k]), (t, k)
                xq_prev[k], t_prev[k] = e.sim.xquat, 0.0
            else:
                e.sim.qvel = g["qvel"][t][k]
                ob, term = e.post_step()
                assert not term and np.array_equal(ob, g["obs"][t][k]), (t, k)
                xq_prev[k], t_prev[k] = g["xquat"][t][k], g["time"][t][k]
        v_prev = g["qvel"][t]


@pytest.mark.parametrize("path", pc.REFCLS, ids=IDS)
def test_kernel_arithmetic_against_reference_class_fixture(path):
    g = np.load(path)
    env_id, seed = str(g["env_id"]), int(g["seed"])
    kind = KIND[env_id]
    rm = model.compile_model(mjcf.parse(SCENE[kind]), kind, 1200 if kind == 3 else 6000)
    env = helpers.EmuVecEnv(rm, g["obs0"].shape[0], seed=seed)
    # Env03-v2 free run: block impacts (7.5 m/s, plain fp32 block state) amplify differences within a few steps; its
    # step-by-step bound is the resync fixture and tests/test_env03_parity.py
    out = pc.replay_reference_class_fixture(env, path, **(dict(tol=1e-4, min_horizon=0) if kind == 3 else {}))
    if bool(g["resync"]):
        assert out["rewards_bit_equal"] == g["reward"].size and out["compared"] > 0.8 * g["reward"].size
        errs = np.array(out.pop("errs"))
        if kind == 3:       # block impacts: contact-timing outliers (tests/test_env03_parity.py states the same bound)
            assert np.quantile(errs, 0.98) < 1e-5 and out["refires"] > 0, (np.quantile(errs, 0.98), out)
        else:
            assert out["max_state_err"] < 1e-6 and out["max_obs_err"] < 5e-3, out
    else:
        assert out["sync_steps"] > (0.2 if kind == 3 else 0.5) * g["reward"].size, out
    env.close()


@pytest.mark.skipif(not pathlib.Path("/root/reference/src/balance_robot").exists(), reason="the reference tree exists only in the build container")
def test_committed_fixtures_equal_a_live_run_of_the_reference_classes():
    sys.path.insert(0, str(pathlib.Path(__file__).parent / "golden"))
    import make_reference_fixtures as mk
    for env_id, n, steps, seed, noise in mk.CASES[1:3]:
        out, _ = mk.generate_quiet(env_id, n, 40, seed, 0, noise, save=False)
        g = np.load(pathlib.Path(__file__).parent / "golden" / f"refcls_{env_id}_free.npz")
        for key in ("obs0", "qpos0"):
            assert np.array_equal(out[key], g[key])
        for key in ("actions", "obs", "reward", "done", "qpos", "qvel", "xquat"):
            assert np.array_equal(out[key], g[key][:40]), key
