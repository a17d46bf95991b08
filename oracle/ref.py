"""ctypes front end of the fp64 CPU ORACLE (oracle/brb_ref.c, oracle/brb_ref_env.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  Nothing under balance_robot_b200/ imports this module.  PARITY UNPINNED (see brb_ref.h).
"""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
LIB_PATH = HERE / "libbrb_ref.so"

MAXBODY, MAXJNT, MAXNQ, MAXNV, MAXGEOM, MAXPAIR, MAXU, MAXCON, MAXEFC = 6, 6, 20, 16, 8, 16, 4, 32, 128
FLAG_ACTDERIV_SKIP_CLAMPED, FLAG_RPY_FROM_FIRST_ROW = 1, 2
FLAG_CYLINDER_BOX = 4            # wheel-block contacts through the own analytic cylinder-box collider (brb_ref.c)
ENV_KINDS = {"Env01-v1": 0, "Env01-v2": 1, "Env01-v3": 2, "Env03-v2": 3}

d, i = C.c_double, C.c_int


class RefModel(C.Structure):
    _fields_ = [
        ("nq", i), ("nv", i), ("nu", i), ("nbody", i), ("njnt", i), ("ngeom", i), ("npair", i), ("flags", i),
        ("timestep", d), ("gravity", d * 3),
        ("body_parent", i * MAXBODY), ("body_jnt", i * MAXBODY),
        ("body_pos", d * 3 * MAXBODY), ("body_quat", d * 4 * MAXBODY),
        ("body_mass", d * MAXBODY), ("body_ipos", d * 3 * MAXBODY), ("body_inertia", d * 9 * MAXBODY),
        ("jnt_type", i * MAXJNT), ("jnt_body", i * MAXJNT), ("jnt_qposadr", i * MAXJNT), ("jnt_dofadr", i * MAXJNT),
        ("jnt_axis", d * 3 * MAXJNT), ("jnt_pos", d * 3 * MAXJNT), ("jnt_damping", d * MAXJNT),
        ("geom_type", i * MAXGEOM), ("geom_body", i * MAXGEOM),
        ("geom_size", d * 3 * MAXGEOM), ("geom_pos", d * 3 * MAXGEOM), ("geom_quat", d * 4 * MAXGEOM),
        ("pair_geom1", i * MAXPAIR), ("pair_geom2", i * MAXPAIR), ("pair_condim", i * MAXPAIR),
        ("pair_friction", d * 5 * MAXPAIR), ("pair_solref", d * 2 * MAXPAIR), ("pair_solimp", d * 5 * MAXPAIR),
        ("pair_margin", d * MAXPAIR), ("pair_gap", d * MAXPAIR),
        ("act_jnt", i * MAXU), ("act_ctrllimited", i * MAXU), ("act_forcelimited", i * MAXU),
        ("act_kv", d * MAXU), ("act_gear", d * MAXU), ("act_ctrlrange", d * 2 * MAXU), ("act_forcerange", d * 2 * MAXU),
        ("qpos0", d * MAXNQ), ("body_invweight0", d * 2 * MAXBODY), ("meaninertia", d),
        ("solver_tolerance", d),
    ]


class RefContact(C.Structure):
    _fields_ = [
        ("dist", d), ("pos", d * 3), ("frame", d * 9), ("includemargin", d), ("friction", d * 5),
        ("solref", d * 2), ("solimp", d * 5),
        ("pair", i), ("dim", i), ("body1", i), ("body2", i), ("efc_address", i), ("exclude", i),
    ]


class RefData(C.Structure):
    _fields_ = [
        ("qpos", d * MAXNQ), ("qvel", d * MAXNV), ("qacc_warmstart", d * MAXNV), ("ctrl", d * MAXU), ("time", d),
        ("xpos", d * 3 * MAXBODY), ("xquat", d * 4 * MAXBODY), ("xmat", d * 9 * MAXBODY), ("xipos", d * 3 * MAXBODY),
        ("xanchor", d * 3 * MAXJNT), ("xaxis", d * 3 * MAXJNT),
        ("geom_xpos", d * 3 * MAXGEOM), ("geom_xmat", d * 9 * MAXGEOM),
        ("qM", d * (MAXNV * MAXNV)),
        ("qfrc_bias", d * MAXNV), ("qfrc_passive", d * MAXNV), ("qfrc_actuator", d * MAXNV), ("actuator_force", d * MAXU),
        ("qfrc_smooth", d * MAXNV), ("qacc_smooth", d * MAXNV), ("qacc", d * MAXNV), ("qfrc_constraint", d * MAXNV),
        ("ncon", i), ("nefc", i), ("solver_niter", i),
        ("contact", RefContact * MAXCON),
        ("efc_J", d * (MAXEFC * MAXNV)), ("efc_pos", d * MAXEFC), ("efc_margin", d * MAXEFC), ("efc_aref", d * MAXEFC),
        ("efc_R", d * MAXEFC), ("efc_D", d * MAXEFC), ("efc_force", d * MAXEFC), ("efc_vel", d * MAXEFC),
        ("stat_substeps", C.c_longlong), ("stat_contact_substeps", C.c_longlong),
        ("stat_newton_iters", C.c_longlong), ("stat_efc_rows", C.c_longlong),
    ]


class RefEnv(C.Structure):
    _fields_ = [
        ("kind", i), ("max_episode_steps", i), ("elapsed_steps", i), ("has_last", i),
        ("last_time", d), ("last_pitch", d), ("target_wheel_speed", d), ("target_yaw", d),
        ("delay_target_speed", d), ("pitch_offset", d),
        ("has_block_timer", i), ("attack_side_front", i), ("block_delay_time_start", d), ("block_delay", d),
        ("d", RefData),
    ]


def build(force: bool = False) -> pathlib.Path:
    """Compile the oracle with gcc (recipe: oracle/Makefile)."""
    srcs = [HERE / n for n in ("brb_ref.c", "brb_ref_env.c", "brb_ref.h", "brb_ref_env.h")]
    if force or not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(HERE), "-B", "libbrb_ref.so"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        assert L.brb_ref_sizeof_model() == C.sizeof(RefModel), (L.brb_ref_sizeof_model(), C.sizeof(RefModel))
        assert L.brb_ref_sizeof_data() == C.sizeof(RefData), (L.brb_ref_sizeof_data(), C.sizeof(RefData))
        assert L.brb_ref_sizeof_contact() == C.sizeof(RefContact)
        assert L.brb_ref_sizeof_env() == C.sizeof(RefEnv), (L.brb_ref_sizeof_env(), C.sizeof(RefEnv))
        L.brb_ref_energy.restype = C.c_double
        L.brb_ref_env_yaw.restype = C.c_double
        L.brb_ref_vec_env.restype = C.POINTER(RefEnv)
        L.brb_ref_vec_env.argtypes = [C.c_void_p, C.c_int]
        L.brb_ref_vec_create.argtypes = [C.POINTER(RefModel), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.brb_ref_vec_destroy.argtypes = [C.c_void_p]
        L.brb_ref_vec_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.brb_ref_vec_step.argtypes = [C.c_void_p] + [C.c_void_p] * 10 + [C.c_int]
        L.brb_ref_philox_draws.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def model_from_spec(spec, flags: int = FLAG_ACTDERIV_SKIP_CLAMPED | FLAG_RPY_FROM_FIRST_ROW) -> RefModel:
    """ModelSpec (balance_robot_b200.mjcf.parse output) -> finalized oracle model."""
    m = RefModel()
    m.nq, m.nv, m.nu = spec.nq, spec.nv, spec.nu
    m.nbody, m.njnt, m.ngeom, m.npair = len(spec.bodies), len(spec.joints), len(spec.geoms), len(spec.pairs)
    assert m.nbody <= MAXBODY and m.njnt <= MAXJNT and m.ngeom <= MAXGEOM and m.npair <= MAXPAIR
    assert m.nq <= MAXNQ and m.nv <= MAXNV and m.nu <= MAXU
    m.flags = flags
    m.timestep = spec.timestep
    m.gravity[:] = spec.gravity
    for b, body in enumerate(spec.bodies):
        m.body_parent[b] = body.parent
        m.body_jnt[b] = body.joint
        m.body_pos[b][:] = body.pos
        m.body_quat[b][:] = body.quat
        m.body_mass[b] = body.mass
        m.body_ipos[b][:] = body.ipos
        m.body_inertia[b][:] = (np.zeros(9) if body.inertia is None else np.asarray(body.inertia).ravel())
    for j, jn in enumerate(spec.joints):
        m.jnt_type[j], m.jnt_body[j], m.jnt_qposadr[j], m.jnt_dofadr[j] = jn.type, jn.body, jn.qposadr, jn.dofadr
        m.jnt_axis[j][:] = jn.axis
        m.jnt_pos[j][:] = jn.pos
        m.jnt_damping[j] = jn.damping
    for g, ge in enumerate(spec.geoms):
        m.geom_type[g], m.geom_body[g] = ge.type, ge.body
        m.geom_size[g][:] = ge.size
        m.geom_pos[g][:] = ge.pos
        m.geom_quat[g][:] = ge.quat
    for p, pr in enumerate(spec.pairs):
        m.pair_geom1[p], m.pair_geom2[p], m.pair_condim[p] = pr.geom1, pr.geom2, pr.condim
        m.pair_friction[p][:] = pr.friction
        m.pair_solref[p][:] = pr.solref
        m.pair_solimp[p][:] = pr.solimp
        m.pair_margin[p], m.pair_gap[p] = pr.margin, pr.gap
    for u, a in enumerate(spec.actuators):
        m.act_jnt[u], m.act_ctrllimited[u], m.act_forcelimited[u] = a.joint, int(a.ctrllimited), int(a.forcelimited)
        m.act_kv[u], m.act_gear[u] = a.kv, a.gear
        m.act_ctrlrange[u][:] = a.ctrlrange
        m.act_forcerange[u][:] = a.forcerange
    rc = lib().brb_ref_model_finalize(C.byref(m))
    if rc != 0:
        raise RuntimeError(f"brb_ref_model_finalize failed: {rc}")
    return m


def new_data(m: RefModel) -> RefData:
    dd = RefData()
    lib().brb_ref_reset_data(C.byref(m), C.byref(dd))
    return dd


def arr(field, n=None) -> np.ndarray:
    """numpy copy of a ctypes array field."""
    a = np.ctypeslib.as_array(field).copy()
    return a if n is None else a[:n]


def philox_draws(seed: int, env0: int, n: int, event: int):
    """(u_step[n,4], u_reset[n,16]) of the counter-based stream shared with the CUDA path."""
    us = np.empty((n, 4), np.float64)
    ur = np.empty((n, 16), np.float64)
    lib().brb_ref_philox_draws(seed, env0, n, event, us.ctypes.data, ur.ctypes.data)
    return us, ur


def philox_blocks(seed: int, env0: int, n: int, event: int, first_block: int, nblocks: int) -> np.ndarray:
    out = np.empty((n, 4 * nblocks), np.float64)
    lib().brb_ref_philox_blocks(C.c_uint64(seed), C.c_uint64(env0), n, C.c_uint32(event), C.c_uint32(first_block), nblocks,
                                C.c_void_p(out.ctypes.data))
    return out


ATTACK_SIDE_EVENT = 0xFFFFFFFF


def env03_draws(seed: int, env0: int, n: int, event: int):
    """(u_step[n,8], u_reset[n,32]) for Env03-v2: reset rows = Philox blocks 1..8, re-fire rows = blocks 9..10."""
    return philox_blocks(seed, env0, n, event, 9, 2), philox_blocks(seed, env0, n, event, 1, 8)


def env03_attack_side(seed: int, env0: int, n: int) -> np.ndarray:
    """attack_side_front = np.random.random() > 0.5, drawn once per env instance (env03_v2.py:22)."""
    return philox_blocks(seed, env0, n, ATTACK_SIDE_EVENT, 0, 1)[:, 0] > 0.5


class RefVecEnv:
    """Vectorised oracle env (DummyVecEnv + TimeLimit + Monitor semantics), draws injected per call."""

    def __init__(self, spec, env_id: str, n: int, max_episode_steps: int, nthreads: int = 1, flags=None):
        self.model = model_from_spec(spec) if flags is None else model_from_spec(spec, flags)
        self.n, self.nthreads, self.kind = n, nthreads, ENV_KINDS[env_id]
        self.reset_stride, self.step_stride = (32, 8) if self.kind == 3 else (16, 4)
        self._h = C.c_void_p()
        rc = lib().brb_ref_vec_create(C.byref(self.model), self.kind, max_episode_steps, n, C.byref(self._h))
        if rc != 0:
            raise MemoryError(rc)
        self.obs = np.zeros((n, 6), np.float32)
        self.reward = np.zeros(n, np.float32)
        self.done = np.zeros(n, np.uint8)
        self.truncated = np.zeros(n, np.uint8)
        self.terminal_obs = np.zeros((n, 6), np.float32)
        self.ep_return = np.zeros(n, np.float32)
        self.ep_len = np.zeros(n, np.int32)

    def set_attack_side(self, front) -> None:
        for k in range(self.n):
            lib().brb_ref_env_set_attack_side(C.byref(self.env(k)), int(bool(front[k])))

    def env(self, k: int) -> RefEnv:
        return lib().brb_ref_vec_env(self._h, k).contents

    def reset(self, u_reset: np.ndarray) -> np.ndarray:
        u_reset = np.ascontiguousarray(u_reset, np.float64)
        assert u_reset.shape == (self.n, self.reset_stride)
        lib().brb_ref_vec_reset(self._h, u_reset.ctypes.data, self.obs.ctypes.data, self.nthreads)
        return self.obs.copy()

    def step(self, actions: np.ndarray, u_step: np.ndarray, u_reset: np.ndarray):
        actions = np.ascontiguousarray(actions, np.float32)
        u_step = np.ascontiguousarray(u_step, np.float64)
        u_reset = np.ascontiguousarray(u_reset, np.float64)
        assert actions.shape == (self.n, 2) and u_step.shape == (self.n, self.step_stride) and u_reset.shape == (self.n, self.reset_stride)
        lib().brb_ref_vec_step(self._h, actions.ctypes.data, u_step.ctypes.data, u_reset.ctypes.data,
                               self.obs.ctypes.data, self.reward.ctypes.data, self.done.ctypes.data,
                               self.truncated.ctypes.data, self.terminal_obs.ctypes.data,
                               self.ep_return.ctypes.data, self.ep_len.ctypes.data, self.nthreads)
        return self.obs.copy(), self.reward.copy(), self.done.copy(), self.truncated.copy()

    def get_state(self):
        nq, nv = self.model.nq, self.model.nv
        qpos = np.stack([arr(self.env(k).d.qpos, nq) for k in range(self.n)])
        qvel = np.stack([arr(self.env(k).d.qvel, nv) for k in range(self.n)])
        return qpos, qvel

    def close(self):
        if self._h:
            lib().brb_ref_vec_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
