"""-m gpu: the CUDA path (through the C-ABI, balance_robot_b200.make_vec) against the fp64 oracle and the committed golden
vectors, plus size-independent properties at BASELINE.json's full sizes.  Nothing here reads /root/reference."""
import pathlib

import numpy as np
import pytest
import torch

import helpers
import parity_checks as pc
from balance_robot_b200 import make_vec, mjcf, model

pytestmark = pytest.mark.gpu
GOLD = sorted((pathlib.Path(__file__).parent / "golden").glob("Env*_n*_s*.npz"))


class GpuAdapter:
    """numpy-in / numpy-out view of BalanceVecEnv for the shared parity procedures."""

    def __init__(self, env_id, n, seed, **kw):
        self.env = make_vec(env_id, n, device="cuda:0", seed=seed, **kw)
        self.n = n

    def reset(self, replay=None):
        r = None if replay is None else torch.as_tensor(replay)
        return self.env.reset(r).cpu().numpy()

    def step(self, act, replay=None):
        r = None if replay is None else torch.as_tensor(replay)
        o, rw, d, info = self.env.step(torch.as_tensor(np.asarray(act, np.float32)).cuda(), r)
        self.tobs, self.epr, self.epl = info.terminal_observation.cpu().numpy(), info.episode_return.cpu().numpy(), info.episode_length.cpu().numpy()
        return o.cpu().numpy().copy(), rw.cpu().numpy().copy(), d.cpu().numpy().copy(), info.truncated.cpu().numpy().copy()

    def get_state(self):
        return tuple(x.cpu().numpy() for x in self.env.get_state())

    def set_state(self, qpos, qvel):
        self.env.set_state(qpos, qvel)

    def stats(self):
        return self.env.stats()

    def close(self):
        self.env.close()


@pytest.fixture(scope="module")
def spec():
    return mjcf.parse("scene_env01.xml")


def test_native_library_is_loaded():
    from balance_robot_b200 import _cabi
    assert _cabi.lib().brb_version() >= 100
    with open("/proc/self/maps") as f:
        assert "libbrb_cuda.so" in f.read()


@pytest.mark.parametrize("kind", [0, 1])
def test_single_step_parity_closed_loop(spec, kind):
    env = GpuAdapter(helpers.ENV_IDS[kind], 32, 3)
    wq, wv = pc.single_step_parity(env, spec, helpers.ENV_IDS[kind], 32, 3, 150)
    assert wq < 1e-5 and wv < 1e-5          # the north-star tolerance (typical: 1e-8 .. 1e-6)
    env.close()


def test_single_step_parity_1000_step_horizon(spec):
    """north_star: single-step qpos/qvel within 1e-5 over 1,000-step horizons (closed-loop PD trajectory)."""
    env = GpuAdapter("Env01-v1", 16, 5)
    wq, wv = pc.single_step_parity(env, spec, "Env01-v1", 16, 5, 1000, noise=0.2)
    assert wq < 1e-5 and wv < 1e-5
    env.close()


def test_single_step_parity_random_actions_with_slip(spec):
    env = GpuAdapter("Env01-v2", 64, 8)
    # 3,770 tumbling env-steps; at most ONE contact-timing event (this seed has the touch-down at fp64 dist = -1.7e-11 m that
    # DESIGN.md §5 dissects: the in-step trajectory difference, 3e-10 m, decides its sign).  Was 1 % before the fp64 predicate.
    q99, outliers = pc.single_step_parity(env, spec, "Env01-v2", 64, 8, 60, policy="random", max_outlier_frac=3e-4)
    assert q99 < 1e-6, q99
    assert env.stats()["nonconverged"] == 0
    env.close()


def test_contact_timing_outliers_are_rare(spec):
    """51,200 env-steps of tumbling robots (random actions, v2's +-1 rad reset pitch): the contact on/off predicate is decided
    in fp64 near dist = 0 (rim_dist_fp64), so what remains are touch-downs whose fp64 distance lies within the in-step
    trajectory difference (~3e-10 m) of zero: measured ~4e-5 of env-steps (was 2e-3 with the fp32 predicate)."""
    env = GpuAdapter("Env01-v2", 512, 21)
    q99, outliers = pc.single_step_parity(env, spec, "Env01-v2", 512, 21, 100, policy="random", max_outlier_frac=2e-4)
    assert q99 < 1e-6, q99
    env.close()


def test_config1_shard_with_256_envs_mirrored_on_the_oracle(spec):
    """BASELINE.json configs[1] as written: 65,536-env shard (512 CTAs, sorted visit order, in-kernel auto-reset, Philox noise
    on, actions U(-1,1)^2 from torch.Generator(cuda).manual_seed(1234)); envs [0, 256) mirrored on the oracle with the same
    initial state, actions and draws, free-running (no re-synchronisation) for 100 steps."""
    n, m = 65536, 256
    env = GpuAdapter("Env01-v2", n, 0)
    g = torch.Generator(device="cuda").manual_seed(1234)
    res = pc.mirrored_free_run(env, spec, "Env01-v2", m, 0, 100, lambda t: (torch.rand((n, 2), device="cuda", generator=g) * 2 - 1).cpu().numpy())
    assert res["both_done"] > 200                                  # auto-resets taken on the same step on both sides
    assert res["done_mismatch"] <= 2, res                          # only a noisy pitch within 1e-8 rad of 50 degrees can differ
    assert res["compared"] > 0.95 * res["total"], res              # almost every env-step was inside a synchronised stretch
    # a free-running env leaves the comparison when its error passes 1e-5: contact-timing events become likelier as the
    # (un-resynchronised) state difference grows from 1e-8, and tumbling robots amplify it faster than the balanced plant's
    # e-fold of 32 steps.  Measured: 22 such events in 25,600 env-steps, 98.9 % of all env-steps inside synchronised stretches.
    assert res["desync_events"] <= 0.002 * res["total"], res
    assert res["max_rew_err"] < 1e-5
    assert env.stats()["nonconverged"] == 0 and env.stats()["unsupported"] == 0
    env.close()


def test_free_run_horizon(spec):
    env = GpuAdapter("Env01-v1", 32, 4)
    h = pc.free_run_horizon(env, spec, "Env01-v1", 32, 4, 200)
    assert h >= 50, h
    env.close()


@pytest.mark.parametrize("kind,steps", [(0, 40), (1, 60), (2, 230)])
def test_task_logic_bit_exact(spec, kind, steps):
    rm = model.compile_model(spec, kind, 6000)
    env = GpuAdapter(helpers.ENV_IDS[kind], 8, 13)
    checked, dones = pc.task_logic_bit_exact(env, rm.time_table, helpers.ENV_IDS[kind], 8, 13, steps, ulps=1)    # 1 f32 ulp: CUDA atan2 vs glibc (assert_f32_equal)
    assert checked > 0.5 * 8 * steps
    env.close()


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_tracks_golden(path):
    g = np.load(path)
    env_id, seed = str(g["env_id"]), int(g["seed"])
    n, steps = g["obs0"].shape[0], g["obs"].shape[0]
    env = GpuAdapter(env_id, n, seed)
    pc.assert_f32_equal(env.reset(), g["obs0"], 1)
    alive = np.ones(n, bool)
    for t in range(steps):
        obs, rew, done, _ = env.step(g["actions"][t])
        alive &= ~(done.astype(bool) | g["done"][t].astype(bool))
        qd, vd, _ = env.get_state()
        eq, ev = helpers.state_errors(qd[alive], vd[alive], g["qpos"][t][alive], g["qvel"][t][alive])
        assert eq.size == 0 or (eq.max() < 1e-5 and ev.max() < 1e-5), (t, eq.max(), ev.max())
    env.close()


@pytest.mark.parametrize("path", pc.REFCLS, ids=[p.stem for p in pc.REFCLS])
def test_device_against_reference_class_fixture(path):
    """tests/golden/refcls_*.npz were recorded from the UNMODIFIED reference env classes (tests/ref_shim +
    tests/golden/make_reference_fixtures.py, oracle physics).  resync fixtures: reward equal to the reference class's to 1 f32 ulp (bit-equal on the host emulation, tests/test_reference_classes.py),
    same termination / truncation / block remove + re-fire decisions, post-step state within 1e-5; free fixtures: the device
    tracks the recorded trajectory until chaotic divergence (>= 50 steps)."""
    g = np.load(path)
    env_id, kind3 = str(g["env_id"]), str(g["env_id"]) == "Env03-v2"
    env = GpuAdapter(env_id, g["obs0"].shape[0], int(g["seed"]))
    out = pc.replay_reference_class_fixture(env, path, ulps=1, **(dict(tol=1e-4, min_horizon=0) if kind3 else {}))
    if bool(g["resync"]):
        errs = np.array(out.pop("errs"))
        assert out["rewards_bit_equal"] == g["reward"].size and out["compared"] > 0.8 * g["reward"].size
        if kind3:
            assert np.quantile(errs, 0.98) < 1e-5 and out["refires"] > 0, (np.quantile(errs, 0.98), out)
        else:
            assert out["max_state_err"] < 1e-5 and out["max_obs_err"] < 5e-3, out
    else:
        assert out["sync_steps"] > (0.2 if kind3 else 0.5) * g["reward"].size, out
    env.close()


def test_device_equals_host_emulation_of_the_same_source(spec):
    """The kernel on the GPU and the same source compiled for the host differ only by FMA contraction / rsqrt rounding."""
    rm = model.compile_model(spec, 1, 6000)
    n = 16
    gpu, emu = GpuAdapter("Env01-v2", n, 21), helpers.EmuVecEnv(rm, n, seed=21)
    pc.assert_f32_equal(gpu.reset(), emu.reset(), 1)
    rng = np.random.default_rng(1)
    for t in range(10):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        og, rg, dg, _ = gpu.step(act)
        oe, re_, de, _ = emu.step(act)
        assert np.array_equal(dg, de)
        qg, vg, _ = gpu.get_state()
        qe, ve, _ = emu.get_state()
        eq, ev = helpers.state_errors(qg, vg, qe, ve)
        assert eq.max() < 1e-6 and ev.max() < 1e-6
    gpu.close(); emu.close()


def test_replay_and_shard_invariance():
    """Philox streams are keyed by global env id: envs [64,128) of one big shard == a second shard with offset 64."""
    n = 128
    a = make_vec("Env01-v2", n, seed=5)
    b = make_vec("Env01-v2", 64, seed=5, env_id_offset=64)
    oa, ob = a.reset(), b.reset()
    assert torch.equal(oa[64:], ob)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(20):
        act = torch.rand((n, 2), device="cuda", generator=g) * 2 - 1
        ra = a.step(act)
        rb = b.step(act[64:].contiguous())
        for x, y in zip(ra[:3], rb[:3]):
            assert torch.equal(x[64:], y)
    a.close(); b.close()


def test_numpy_mode_matches_torch_mode_and_sb3_info_contract():
    n = 1000          # not a multiple of the compaction kernels' 256-thread CTAs
    a = make_vec("Env01-v2", n, seed=9)
    b = make_vec("Env01-v2", n, seed=9, output="numpy")
    oa, ob = a.reset(), b.reset()
    assert np.array_equal(oa.cpu().numpy(), ob) and ob.dtype == np.float32 and ob.shape == (n, 6)
    rng = np.random.default_rng(0)
    seen_done = 0
    for _ in range(15):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        o1, r1, d1, i1 = a.step(torch.as_tensor(act).cuda())
        o2, r2, d2, i2 = b.step(act)
        assert np.array_equal(o1.cpu().numpy(), o2) and np.array_equal(r1.cpu().numpy(), r2)
        assert np.array_equal(d1.cpu().numpy().astype(bool), d2) and d2.dtype == bool and isinstance(i2, list)
        for k in np.flatnonzero(d2):
            seen_done += 1
            assert set(i2[k]) == {"terminal_observation", "TimeLimit.truncated", "episode"}
            assert i2[k]["terminal_observation"].shape == (6,) and i2[k]["episode"]["l"] >= 1
            assert i1[int(k)]["episode"]["l"] == i2[k]["episode"]["l"] and i1[int(k)]["episode"]["r"] == i2[k]["episode"]["r"]
            assert np.array_equal(i1[int(k)]["terminal_observation"], i2[k]["terminal_observation"])
            assert i1[int(k)]["TimeLimit.truncated"] == i2[k]["TimeLimit.truncated"]
        assert all(i2[k] == {} for k in np.flatnonzero(~d2)[:5])
    assert seen_done > 0          # Q3: 12.8 % of v2 episodes end at their first step
    a.close(); b.close()


def test_host_step_compact_equals_full_host_step():
    """brb_env_step_host (every per-episode array in full) and brb_env_step_host_compact (finished envs only) on twin envs."""
    import ctypes as C
    from balance_robot_b200 import _cabi
    n = 70000
    a = make_vec("Env01-v2", n, seed=4, output="numpy")
    b = make_vec("Env01-v2", n, seed=4, output="numpy")
    a.reset(); b.reset()
    L = _cabi.lib()
    obs = np.zeros((n, 6), np.float32); rew = np.zeros(n, np.float32); done = np.zeros(n, np.uint8); trunc = np.zeros(n, np.uint8)
    tobs = np.zeros((n, 6), np.float32); epr = np.zeros(n, np.float32); epl = np.zeros(n, np.int32)
    rng = np.random.default_rng(1)
    total = 0
    for k in range(20):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        _cabi.check(L.brb_env_step_host(a._env, act.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, trunc.ctypes.data,
                                        tobs.ctypes.data, epr.ctypes.data, epl.ctypes.data), "brb_env_step_host")
        o2, r2, d2, infos = b.step(act)
        assert np.array_equal(obs, o2) and np.array_equal(rew, r2) and np.array_equal(done.astype(bool), d2)
        idx = np.flatnonzero(done)
        assert np.array_equal(idx, infos._idx)
        total += len(idx)
        for i in idx[:: max(1, len(idx) // 50)]:
            d = infos[int(i)]
            assert np.array_equal(d["terminal_observation"], tobs[i]) and d["episode"] == {"r": float(epr[i]), "l": int(epl[i]), "t": d["episode"]["t"]}
            assert d["TimeLimit.truncated"] == bool(trunc[i])
    assert total > 1000
    a.close(); b.close()


def test_host_step_compact_with_pageable_buffers_and_any_layout():
    """The compact host call must not depend on HOW the caller allocated its buffers: pinned block laid out as
    brb_env_host_layout says (one copy, rows written by the device), or separate pageable numpy arrays (three copies, rows
    in a second copy) -- same bytes out.  Also the capacity error: more finished envs than max_rows -> BRB_EINVAL."""
    import ctypes as C
    from balance_robot_b200 import _cabi
    n = 5000                                   # not a multiple of 64: the staging offsets carry alignment padding
    a = make_vec("Env01-v2", n, seed=9, output="numpy")
    b = make_vec("Env01-v2", n, seed=9, output="numpy")
    a.reset(); b.reset()
    L = _cabi.lib()
    offs, total = (C.c_int64 * 3)(), C.c_int64()
    _cabi.check(L.brb_env_host_layout(a._env, C.byref(offs), C.byref(total)), "brb_env_host_layout")
    assert offs[0] == 0 and offs[1] >= 24 * n and offs[2] >= offs[1] + 4 * n and total.value == offs[2] + n
    obs = np.zeros((n, 6), np.float32); rew = np.zeros(n, np.float32); done = np.zeros(n, np.uint8)
    rows = np.zeros((n, _cabi.DONE_ROW_WORDS), np.float32); nd = np.zeros(1, np.int32)
    rng = np.random.default_rng(2)
    seen = 0
    for k in range(40):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        _cabi.check(L.brb_env_step_host_compact(a._env, act.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data,
                                                nd.ctypes.data, rows.ctypes.data, n), "brb_env_step_host_compact")
        o2, r2, d2, infos = b.step(act)
        assert np.array_equal(obs, o2) and np.array_equal(rew, r2) and np.array_equal(done.astype(bool), d2)
        assert int(nd[0]) == len(infos._idx) == int(done.sum())
        assert np.array_equal(rows[:int(nd[0])].view(np.int32)[:, 0], infos._idx)
        if nd[0]:
            i = int(infos._idx[0])
            assert np.array_equal(infos[i]["terminal_observation"], rows[0, 1:7])
        seen += int(nd[0])
    assert seen > 100
    # capacity error (pageable and pinned row buffers alike): run until some env finishes with room for no row at all
    for k in range(200):
        act = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        rc = L.brb_env_step_host_compact(a._env, act.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, nd.ctypes.data, None, 0)
        if nd[0] > 0:
            assert rc == -22, rc          # BRB_EINVAL
            break
        assert rc == 0
    else:
        raise AssertionError("no episode finished")
    a.close(); b.close()


def test_full_size_properties():
    """BASELINE.json configs[1] size (65,536 envs): determinism, unit quaternions, finite state, episode accounting."""
    n = 65536
    outs = []
    for rep in range(2):
        env = make_vec("Env01-v2", n, seed=0)
        env.reset()
        g = torch.Generator(device="cuda").manual_seed(1234)
        tot_done = 0
        for _ in range(30):
            act = torch.rand((n, 2), device="cuda", generator=g) * 2 - 1
            o, r, d, info = env.step(act)
            tot_done += int(d.sum())
        qpos, qvel, xq = env.get_state()
        st = env.stats()
        assert torch.isfinite(qpos).all() and torch.isfinite(qvel).all() and torch.isfinite(o).all()
        assert (qpos[:, 3:7].norm(dim=1) - 1).abs().max() < 1e-6
        assert (xq.norm(dim=1) - 1).abs().max() < 1e-12
        assert st["episodes"] == tot_done and st["env_steps"] == 30 * n and st["substeps"] == 30 * n * 250
        assert st["nonconverged"] < 1e-5 * st["contact_substeps"]
        assert 0.10 < tot_done / (30 * n) * 30 < 3.0
        outs.append((qpos.clone(), o.clone()))
        env.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])   # bitwise reproducible


def test_truncation_and_monitor_on_device(spec):
    env = make_vec("Env01-v1", 64, seed=2)
    env.reset()
    # drive elapsed to the limit quickly is impractical (6000 steps); check the counter and the episode bookkeeping instead
    z = torch.zeros((64, 2), device="cuda")
    total = torch.zeros(64, device="cuda", dtype=torch.float64)
    for t in range(1, 8):
        o, r, d, info = env.step(z)
        total += r.double()
        assert (env.elapsed_steps().cpu().numpy() == np.where(d.cpu().numpy() > 0, 0, t)).all() or d.any()
    live = (d == 0).cpu()
    assert torch.allclose(info.episode_return.cpu()[live].double(), total.cpu()[live], rtol=1e-5)
    env.close()


def test_episode_statistics_agree_with_oracle(spec):
    """north_star: episode returns must agree statistically.  Same stochastic policy (weak PD + uniform exploration noise), independent noise
    streams: episode length / return distributions of the CUDA path (thousands of episodes) vs the fp64 oracle."""
    from oracle import ref
    rng = np.random.default_rng(0)

    def run(step_fn, reset_fn, n, steps):
        obs = reset_fn()
        lens, rets = [], []
        for t in range(steps):
            act = np.clip(0.03 * helpers.pd_policy(obs) + rng.uniform(-1, 1, (n, 2)), -1, 1).astype(np.float32)   # weak, noisy policy: episodes of ~40 steps
            obs, done, epr, epl = step_fn(act, t)
            lens += epl[done].tolist()
            rets += epr[done].tolist()
        return np.array(lens), np.array(rets)

    n_o, n_g, steps = 256, 8192, 200
    rv = ref.RefVecEnv(spec, "Env01-v2", n_o, 6000, nthreads=8)

    def o_reset():
        return rv.reset(ref.philox_draws(77, 0, n_o, 0)[1])

    def o_step(act, t):
        us, ur = ref.philox_draws(77, 0, n_o, t + 1)
        obs, rew, done, trunc = rv.step(act, us, ur)
        return obs, done.astype(bool), rv.ep_return.copy(), rv.ep_len.copy()
    lo, ro = run(o_step, o_reset, n_o, steps)
    env = GpuAdapter("Env01-v2", n_g, 78)

    def g_step(act, t):
        obs, rew, done, trunc = env.step(act)
        return obs, done.astype(bool), env.epr.copy(), env.epl.copy()
    lg, rg = run(g_step, env.reset, n_g, steps)
    assert len(lo) > 300 and len(lg) > 10000
    # one-step episodes (Q3: ~12.8 % start beyond 50 degrees)
    assert abs((lo == 1).mean() - (lg == 1).mean()) < 0.05
    assert abs((lg == 1).mean() - 0.128) < 0.03
    # means within 4 standard errors of the (smaller) oracle sample
    for a, b in ((lo, lg), (ro, rg)):
        se = a.std() / np.sqrt(len(a)) + b.std() / np.sqrt(len(b))
        assert abs(a.mean() - b.mean()) < 4 * se + 1e-9, (a.mean(), b.mean(), se)
    # quartiles of the episode-length distribution
    qo, qg = np.quantile(lo, [0.25, 0.5, 0.75]), np.quantile(lg, [0.25, 0.5, 0.75])
    assert np.all(np.abs(qo - qg) <= np.maximum(3, 0.25 * qg)), (qo, qg)
    env.close(); rv.close()


def test_seed_rekeys_the_streams_and_reset_draws_new_start_states(spec):
    """VecEnv.reset() called again must draw NEW start states (the reference's RNG streams keep advancing across resets; an
    eval loop that resets before every evaluation would otherwise judge the policy on the same few trajectories forever):
    the reset epoch is folded into the Philox block index.  seed() re-keys the streams and restarts the epochs."""
    from oracle import ref
    n = 64
    a = make_vec("Env01-v2", n, seed=1)
    o1 = a.reset().clone()
    o1b = a.reset().clone()
    assert not torch.equal(o1, o1b)              # second reset: epoch 1
    rv = ref.RefVecEnv(spec, "Env01-v2", n, 6000)
    assert np.allclose(rv.reset(ref.philox_draws(1, 0, n, 0)[1]), o1.cpu().numpy(), rtol=2.5e-7, atol=0)
    assert np.allclose(rv.reset(ref.philox_blocks(1, 0, n, 0, 1 + 16, 4)), o1b.cpu().numpy(), rtol=2.5e-7, atol=0)   # blocks 17..20 of event 0
    a.seed(2)
    o2 = a.reset().clone()
    assert not torch.equal(o1, o2)
    b = make_vec("Env01-v2", n, seed=2)
    assert torch.equal(o2, b.reset())
    a.close(); b.close(); rv.close()


def test_time_limit_truncation_on_device():
    """gymnasium TimeLimit(6000) (balance_robot/__init__.py:15): robots balanced by the PD controller reach the limit; the step
    that hits it reports done with TimeLimit.truncated, the episode record has l = 6000, and the env auto-resets."""
    n = 64
    env = make_vec("Env01-v1", n, seed=4)
    obs = env.reset()
    reached = torch.zeros(n, dtype=torch.bool, device="cuda")
    gains = torch.tensor([12.0, 0.6, 0.35 * 0.034 * 170.0 / 4.0], device="cuda")
    for t in range(1, 6001):
        u = (gains[0] * obs[:, 0] * 0.25 + gains[1] * obs[:, 1] - gains[2] * obs[:, 4]).clamp(-1, 1)     # helpers.pd_policy on the device
        obs, r, d, info = env.step(torch.stack([-u, u], 1))
        if t < 6000:
            assert not bool(info.truncated.any())
    trunc = info.truncated.bool()
    assert int(trunc.sum()) >= 2, int(trunc.sum())               # some robots stayed up for the whole 30 s
    assert bool((d.bool() | ~trunc).all())
    assert bool((info.episode_length[trunc] == 6000).all())
    assert bool((info.episode_return[trunc] > 4000).all())       # ~1 per step while upright (RobotBaseEnv._get_reward)
    assert bool((env.elapsed_steps()[trunc] == 0).all())         # auto-reset
    assert bool((obs[trunc][:, 1] == 0).all())                   # Q6 on the reset observation
    rec = info[int(torch.nonzero(trunc)[0])]
    assert rec["TimeLimit.truncated"] is True and rec["episode"]["l"] == 6000
    env.close()


def test_every_kernel_at_sizes_off_the_cta_and_tile_grid():
    """compute-sanitizer is closed on this pool, so the bounds of every kernel are exercised with sizes that are not
    multiples of the CTA (128), compaction (256) or PPO tile (128) sizes, next to canary-guarded output buffers."""
    from balance_robot_b200.ppo import PPO, PPOConfig
    for env_id, n in (("Env01-v1", 77), ("Env01-v2", 333), ("Env01-v3", 130), ("Env03-v2", 70)):
        env = make_vec(env_id, n, seed=1)
        env.reset()
        for _ in range(3):
            obs, r, d, info = env.step(torch.rand((n, 2), device="cuda") * 2 - 1)
        assert torch.isfinite(obs).all() and obs.shape == (n, 6) and r.shape == (n,) and d.shape == (n,)
        q = env.get_state()
        env.set_state(q[0], q[1])
        assert env.elapsed_steps().shape == (n,) and env.stats()["env_steps"] == 3 * n
        env.close()
    envh = make_vec("Env01-v2", 300, seed=2, output="numpy")
    envh.reset()
    rows = [hb["rows"] for hb in envh._hbuf]
    for k in range(4):
        for rws in rows:
            rws.fill_(-77.0)                                  # canary behind the compacted finished-episode rows
        o, r, d, infos = envh.step(np.random.default_rng(k).uniform(-1, 1, (300, 2)).astype(np.float32))
        nd = int(d.sum())
        assert all(infos[int(i)]["episode"]["l"] >= 1 for i in np.flatnonzero(d))
        assert (envh._hbuf[envh._flip]["rows"][nd:] == -77.0).all()
    envh.close()
    env = make_vec("Env01-v2", 200, seed=3)
    agent = PPO(env, PPOConfig(n_steps=5, n_epochs=1, n_minibatches=3, seed=0), device="cuda:0")      # minibatches of 333 samples
    agent.collect_rollouts()
    out = agent.train()
    assert all(np.isfinite(v) for v in out.values())
    assert torch.isfinite(agent.policy.pack_params()).all()
    env.close()
