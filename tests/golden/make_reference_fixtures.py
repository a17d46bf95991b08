"""Generates tests/golden/refcls_*.npz by running the UNMODIFIED reference env classes
(/root/reference/src/balance_robot/envs/{env01_v1,env01_v2,env01_v3,env03_v2}.py on RobotBaseEnv.py) through
tests/ref_shim: stand-in `mujoco` / `gymnasium` modules whose mj_step is the fp64 oracle and whose random streams are
injected.  What these vectors pin is therefore the TASK LOGIC — reward (pre-step, stale kinematics), observation,
termination, reset_model incl. the quaternion-order bug, the v3 schedule, the Env03-v2 block remove / delay / re-fire
state machine and the order in which every random number is consumed — to the reference's own code, executed here.
The physics underneath is the oracle's restatement: `mj_step` parity with MuJoCo stays unpinned.

Around each env the script restates what gymnasium's TimeLimit and SB3's DummyVecEnv do (both third party): elapsed-step
truncation, terminal observation kept, reset on done.

Two fixtures per env id:
  refcls_<id>_free.npz    free-running episodes (the reward sees kinematics one substep stale)
  refcls_<id>_resync.npz  env.set_state(qpos, qvel) (-> mj_forward, fresh kinematics) before every step: a device path can
                          be put into the identical pre-step state with set_state, so its reward must be BIT-equal and its
                          post-step state / observation within the physics tolerance

Needs /root/reference (this container only).  Re-run:  python tests/golden/make_reference_fixtures.py
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import helpers  # noqa: E402
import ref_shim  # noqa: E402
from oracle import ref  # noqa: E402

OUT = pathlib.Path(__file__).resolve().parent
GAIN = np.array([1.0, 0.05, 1.0, 0.0])
MAX_STEPS = {"Env01-v1": 6000, "Env01-v2": 6000, "Env01-v3": 6000, "Env03-v2": 1200}   # balance_robot/__init__.py:5-52


def draws_for(env_id, seed, n, event):
    return ref.env03_draws(seed, 0, n, event) if env_id == "Env03-v2" else ref.philox_draws(seed, 0, n, event)


def push_reset_draws(env_id, u):
    """slot layout -> the two streams, in the order the reference consumes them."""
    Q = ref_shim.QUEUES
    u = list(u)
    if env_id == "Env03-v2":       # env03_v1.py:60-83: 16 np_random jitters, 3 rotation draws, set_block_pos_vel (env03_v2.py:42-53: 5 draws)
        Q.gym_u += u[0:16]
        Q.global_u += u[16:24]
    elif env_id == "Env01-v3":     # env01_v3.py:44,52 then env01_v1.py:40-49
        Q.gym_u += [u[12], u[13]] + u[0:9]
        Q.global_u += u[9:12]
    elif env_id == "Env01-v2":     # env01_v2.py:53-62, then _get_obs: get_pitch (RobotBaseEnv.py:224), get_pitch_dot_alt (:145)
        Q.gym_u += u[0:9]
        Q.global_u += u[9:12] + [u[12], u[13]]
    else:                          # env01_v1.py:40-49
        Q.gym_u += u[0:9]
        Q.global_u += u[9:12]


def generate(env_id, n, steps, seed, resync, noise, save=True):
    Q = ref_shim.QUEUES
    envs = []
    side = ref.env03_attack_side(seed, 0, n) if env_id == "Env03-v2" else None
    for k in range(n):
        if env_id == "Env03-v2":
            # Env03_v2.__init__: attack_side_front = np.random.random() > 0.5 (env03_v2.py:22); ref.env03_attack_side is `u > 0.5` of this draw
            Q.global_u.append(float(ref.philox_blocks(seed, k, 1, ref.ATTACK_SIDE_EVENT, 0, 1)[0, 0]))
        envs.append(ref_shim.make(env_id))
        Q.assert_drained()
        if side is not None:
            assert bool(envs[k].attack_side_front) == bool(side[k])
    us, ur = draws_for(env_id, seed, n, 0)
    obs = np.zeros((n, 6), np.float32)
    for k, e in enumerate(envs):
        push_reset_draws(env_id, ur[k])
        obs[k], _ = e.reset()
        Q.assert_drained()
    nq, nv = envs[0].model.nq, envs[0].model.nv
    elapsed = np.zeros(n, int)
    rec = dict(obs0=obs.copy(), qpos0=np.stack([e.data.qpos.copy() for e in envs]), qvel0=np.stack([e.data.qvel.copy() for e in envs]),
               actions=[], obs=[], reward=[], done=[], truncated=[], terminal_obs=[], qpos=[], qvel=[], xquat=[], time=[], refired=[])
    rng = np.random.default_rng(seed)
    for t in range(1, steps + 1):
        # every other robot gets a crippled controller so that episodes end (termination, reset) inside the fixture
        act = np.clip(GAIN[np.arange(n) % 4, None] * helpers.pd_policy(obs) + noise * rng.uniform(-1, 1, (n, 2)), -1, 1).astype(np.float32)
        us, ur = draws_for(env_id, seed, n, t)
        rew = np.zeros(n); done = np.zeros(n, np.uint8); trunc = np.zeros(n, np.uint8); tobs = np.zeros((n, 6), np.float32)
        xq = np.zeros((n, 4)); tm = np.zeros(n); refired = np.zeros(n, np.uint8)
        for k, e in enumerate(envs):
            if resync:
                e.set_state(e.data.qpos.copy(), e.data.qvel.copy())
            if env_id == "Env01-v2":
                Q.global_u += list(us[k][:4])          # reward, termination, obs pitch, obs pitch-rate
            elif env_id == "Env03-v2":
                Q.global_u += list(us[k][:5])          # consumed only if the block is re-fired this step
            ob, r, terminated, truncated, info = e.step(act[k])
            if env_id == "Env03-v2":
                refired[k] = len(Q.global_u) == 0
                Q.global_u.clear()
            Q.assert_drained()
            assert truncated is False and info == {}
            elapsed[k] += 1
            tr = elapsed[k] >= MAX_STEPS[env_id]       # gymnasium TimeLimit
            rew[k], xq[k], tm[k] = r, e.data.body("robot_body").xquat, e.data.time
            if terminated or tr:                        # DummyVecEnv: keep the terminal observation, reset
                done[k], trunc[k], tobs[k] = 1, int(tr and not terminated), ob
                push_reset_draws(env_id, ur[k])
                ob, _ = e.reset()
                Q.assert_drained()
                elapsed[k] = 0
            obs[k] = ob
        for key, x in zip(("actions", "obs", "reward", "done", "truncated", "terminal_obs", "qpos", "qvel", "xquat", "time", "refired"),
                          (act, obs, rew, done, trunc, tobs, np.stack([e.data.qpos.copy() for e in envs]),
                           np.stack([e.data.qvel.copy() for e in envs]), xq, tm, refired)):
            rec[key].append(np.array(x, copy=True))
    out = {k: np.stack(v) if isinstance(v, list) else v for k, v in rec.items()}
    out["seed"], out["env_id"], out["resync"] = np.int64(seed), np.array(env_id), np.int64(resync)
    if side is not None:
        out["attack_side_front"] = np.asarray(side, np.uint8)
    name = OUT / f"refcls_{env_id}_{'resync' if resync else 'free'}.npz"
    if save:
        np.savez_compressed(name, **out)
    print(name.name, "episodes finished:", int(out["done"].sum()), "re-fires:", int(out["refired"].sum()), {k: v.shape for k, v in out.items() if getattr(v, "ndim", 0) > 1})
    return out


CASES = (("Env01-v1", 4, 200, 201, 0.5), ("Env01-v2", 8, 120, 202, 0.5), ("Env01-v3", 4, 260, 203, 0.3), ("Env03-v2", 4, 260, 204, 0.2))


def generate_quiet(*args, **kw):
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):      # env01_v3.py:53 prints the pitch offset at every reset
        ref_shim.install()
        out = generate(*args, **kw)
    return out, buf.getvalue().strip().splitlines()[-1]


if __name__ == "__main__":
    for env_id, n, steps, seed, noise in CASES:
        for resync in (0, 1):
            print(generate_quiet(env_id, n, steps, seed + 10 * resync, resync, noise)[1])
