"""Behavioural pin against real MuJoCo: the reference's OWN trained policy drives the restated physics.

RobotMovePolicy.tflite (reference: src/balance_robot/envs/RobotMovePolicy.tflite, used by RobotMoveBaseEnv.py:81-210) was
trained by the reference author with PPO against the real `mujoco.mj_step` on this robot model.  It is the only artefact in
the reference produced by the real engine.  A controller tuned to one plant balances another only if mass, inertia, motor
model, contact and time step are close, so "it balances here, a zero action does not" is evidence — behavioural, not
numerical — that the oracle (and the kernel arithmetic) reproduce the reference's dynamics.  CPU only.
"""
import numpy as np
import pytest
import torch

import helpers
from balance_robot_b200 import mjcf, model
from oracle import ref
from reference_policy import FIXTURE, OP_FULLY_CONNECTED, OP_TANH, RobotMovePolicy

FALL = 50 * np.pi / 180


@pytest.fixture(scope="module")
def spec():
    return mjcf.parse("scene_env01.xml")


@pytest.fixture(scope="module")
def policy():
    return RobotMovePolicy()


def test_fixture_is_the_sb3_mlp_policy():
    z = np.load(FIXTURE)
    codes = [int(z[f"op{k}_code"]) for k in range(int(z["n_ops"]))]
    assert codes.count(OP_FULLY_CONNECTED) == 7 and codes.count(OP_TANH) == 4      # pi 6-64-64-2 (mean + action head), vf 6-64-64-1
    assert list(z["t0_shape"]) == [1, 6] and str(z["t0_dtype"]) == "int8"
    assert [list(z[f"t{k}_shape"]) for k in (10, 8, 6)] == [[64, 6], [64, 64], [2, 64]]
    assert list(z["outputs"]) == [21, 27, 35]


def test_integer_interpreter_matches_float_network(policy):
    """Independent evaluation with de-quantised weights in float: the int8 path may differ by activation rounding only."""
    z = np.load(FIXTURE)
    deq = lambda k: z[f"t{k}_data"].astype(np.float64) * z[f"t{k}_scale"].astype(np.float64).reshape(-1, *([1] * (z[f"t{k}_data"].ndim - 1)))
    g = torch.Generator().manual_seed(0)
    obs = (torch.rand((512, 6), generator=g) * 2 - 1) * torch.tensor([1.5, 2.0, 1.0, 1.0, 1.0, 1.0])
    x = (policy.quantize_obs(obs).numpy() - policy.in_zero) * policy.in_scale
    h = np.tanh(x @ deq(10).T + deq(9))
    h = np.tanh(h @ deq(8).T + deq(7))
    y = h @ deq(6).T + deq(4)
    y = np.clip(y, float(policy.out_scale) * (-128 - policy.out_zero), float(policy.out_scale) * (127 - policy.out_zero))   # int8 output range
    got = policy.act(obs).numpy()
    assert np.abs(got - y).max() < 0.1, np.abs(got - y).max()         # output LSB 0.016; hidden activations are int8 too
    assert np.abs(got - y).mean() < 0.012


def _drive(rv, policy, seed, n, steps, draws=None):
    draws = draws or (lambda k: ref.philox_draws(seed, 0, n, k))
    obs = rv.reset(draws(0)[1])
    rew, fell = [], np.zeros(n, bool)
    for k in range(1, steps + 1):
        us, ur = draws(k)
        obs, r, d, tr = rv.step(policy.act(torch.from_numpy(obs)).numpy(), us, ur)
        rew.append(r.mean())
        fell |= d.astype(bool) & ~tr.astype(bool)
    return np.asarray(rew), fell, obs


def test_reference_policy_balances_the_oracle_robot(spec, policy):
    n = 8
    rv = ref.RefVecEnv(spec, "Env01-v1", n, 6000, nthreads=4)
    rew, fell, _ = _drive(rv, policy, 3, n, 700)
    assert not fell.any()
    assert rew[300:].mean() > 0.95, rew[300:].mean()        # reward 1 = upright, still, not yawing (RobotBaseEnv.py:190-219)


def test_zero_action_does_not_balance(spec):
    """The control for the test above: without the policy every robot is past 50 degrees within 2 s."""
    n = 8
    rv = ref.RefVecEnv(spec, "Env01-v1", n, 6000, nthreads=4)
    obs = rv.reset(ref.philox_draws(3, 0, n, 0)[1])
    fell = np.zeros(n, bool)
    for k in range(1, 401):
        us, ur = ref.philox_draws(3, 0, n, k)
        _, _, d, _ = rv.step(np.zeros((n, 2), np.float32), us, ur)
        fell |= d.astype(bool)
    assert fell.all()


def test_reference_policy_follows_the_v3_speed_schedule(spec, policy):
    """Env01-v3 commands +d, -d, 2d, 3d wheel speed at t = 1, 3, 4.5, 5.5 s (env01_v3.py:27-37), |d| in [10, 20] rad/s."""
    n = 8
    rv = ref.RefVecEnv(spec, "Env01-v3", n, 6000, nthreads=4)
    rew, fell, obs = _drive(rv, policy, 3, n, 1300)
    assert not fell.any()
    err = np.abs(obs[:, 4]) * 170.0 / 4.0                   # |target - wheel speed| one second after the last switch, rad/s
    target = np.array([abs(rv.env(k).target_wheel_speed) for k in range(n)])
    assert (target >= 30).all() and (err < 0.4 * target).all(), (err, target)


def test_reference_policy_balances_the_kernel_arithmetic(spec, policy):
    """Same drive on the host emulation of the CUDA kernel's fp32 arithmetic (tests/host_emu), next to the oracle."""
    n, steps, seed = 8, 500, 3
    rm = model.compile_model(spec, 0, 6000)
    emu = helpers.EmuVecEnv(rm, n, seed=seed)
    rv = ref.RefVecEnv(spec, "Env01-v1", n, 6000, nthreads=4)
    obs_e = emu.reset()
    obs_r = rv.reset(ref.philox_draws(seed, 0, n, 0)[1])
    assert np.array_equal(obs_e, obs_r)
    re, rr = [], []
    for k in range(1, steps + 1):
        obs_e, r1, d1, _ = emu.step(policy.act(torch.from_numpy(obs_e)).numpy())
        us, ur = ref.philox_draws(seed, 0, n, k)
        obs_r, r2, d2, _ = rv.step(policy.act(torch.from_numpy(obs_r)).numpy(), us, ur)
        assert not d1.any() and not d2.any()
        re.append(r1.mean()); rr.append(r2.mean())
    # closed loop with an int8 policy: trajectories decorrelate once a rounding flips, the statistics must not
    assert abs(np.mean(re[250:]) - np.mean(rr[250:])) < 0.01, (np.mean(re[250:]), np.mean(rr[250:]))
    emu.close()
