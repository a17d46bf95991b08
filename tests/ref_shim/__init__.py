"""TEST-ONLY stand-ins for the third-party modules the reference env classes import (`mujoco`, `gymnasium`), so that the
UNMODIFIED classes under /root/reference/src/balance_robot/envs run in this container, where neither is installable:

  mujoco.mj_step / mj_forward / mj_resetData / mj_rnePostConstraint   -> the fp64 oracle (oracle/brb_ref.c)
  mujoco.MjModel.from_xml_path                                         -> balance_robot_b200.mjcf.parse on the REFERENCE's own XML
  gymnasium.envs.mujoco.MujocoEnv                                      -> the few members the env classes touch
      (envs/RobotBaseEnv.py:56-65: __init__(model_path, frame_skip, observation_space, ...); model.nq, data.body(n).xquat /
       .xpos, data.joint(n).qpos / .qvel, data.actuator(n).ctrl, data.time, init_qpos / init_qvel, set_state -> mj_forward,
       reset -> mj_resetData + reset_model, np_random.uniform, frame_skip, render_mode, unwrapped.mujoco_renderer.viewer)
  gymnasium.spaces.Box, gymnasium.utils.EzPickle, gymnasium.envs.registration.register

Random numbers are INJECTED: `np.random.random` (the global legacy stream the reference uses for the
rotation, the v2 noise and the block) and `self.np_random.uniform` (the gymnasium-seeded Generator) pop from two queues
the caller fills, so a fixture records exactly which uniform went where.

Only tests/golden/make_reference_fixtures.py and tests/test_reference_classes.py use this; it needs /root/reference and
therefore never runs on the GPU box (the fixtures it generates do).  The physics underneath is the oracle, NOT MuJoCo:
what this pins is the task logic (reward / observation / termination / reset / block state machine / RNG call order)
to the reference's own code; `mj_step` itself stays unpinned.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import sys
import types

import numpy as np

REFERENCE_SRC = pathlib.Path("/root/reference/src")
ROOT = pathlib.Path(__file__).resolve().parents[2]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def available() -> bool:
    return (REFERENCE_SRC / "balance_robot" / "envs" / "RobotBaseEnv.py").exists()


class DrawQueues:
    """The two random streams of the reference, replaced by caller-filled FIFOs."""

    def __init__(self):
        self.global_u = []       # np.random.random()
        self.gym_u = []          # self.np_random.uniform(...)

    def random(self):
        return self.global_u.pop(0)

    def uniform(self, low=0.0, high=1.0, size=None):
        # numpy Generator.uniform: low + (high - low) * U[0, 1)
        if size is None:
            return low + (high - low) * self.gym_u.pop(0)
        n = int(np.prod(size))
        u = np.array([self.gym_u.pop(0) for _ in range(n)])
        return (low + (high - low) * u).reshape(size)

    def assert_drained(self):
        assert not self.global_u and not self.gym_u, (len(self.global_u), len(self.gym_u))


QUEUES = DrawQueues()
REGISTERED = {}
_installed = False


def install():
    """Put the stand-in modules into sys.modules, patch np.random.random, add the reference to sys.path."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("/root/reference is not present: the reference classes cannot run here")
    from balance_robot_b200 import mjcf
    from oracle import ref
    L = ref.lib()

    class _Body:
        def __init__(self, data, bid):
            self._d, self._b = data, bid

        @property
        def xquat(self):
            return self._d._xquat[self._b]

        @property
        def xpos(self):
            return self._d._xpos[self._b]

    class _Joint:
        def __init__(self, data, j):
            spec = data._model.spec
            jn = spec.joints[j]
            nq, nv = (7, 6) if jn.type == mjcf.JNT_FREE else (1, 1)
            self.qpos = data.qpos[jn.qposadr:jn.qposadr + nq]      # numpy views onto mjData
            self.qvel = data.qvel[jn.dofadr:jn.dofadr + nv]

    class _Actuator:
        def __init__(self, data, u):
            self._d, self._u = data, u

        @property
        def ctrl(self):
            return self._d._ctrl[self._u:self._u + 1]

        @ctrl.setter
        def ctrl(self, v):
            self._d._ctrl[self._u] = np.asarray(v, np.float64).ravel()[0]

    class MjModel:
        def __init__(self, spec):
            self.spec = spec
            self._m = ref.model_from_spec(spec)
            self.nq, self.nv, self.nu = spec.nq, spec.nv, spec.nu

        @classmethod
        def from_xml_path(cls, path):
            return cls(mjcf.parse(path))

    class MjData:
        def __init__(self, model):
            self._model = model
            self._raw = ref.new_data(model._m)
            as_arr = np.ctypeslib.as_array
            self.qpos = as_arr(self._raw.qpos)[:model.nq]
            self.qvel = as_arr(self._raw.qvel)[:model.nv]
            self._ctrl = as_arr(self._raw.ctrl)
            self._xquat = as_arr(self._raw.xquat)
            self._xpos = as_arr(self._raw.xpos)

        @property
        def time(self):
            return self._raw.time

        @property
        def ctrl(self):
            return self._ctrl[:self._model.nu]

        def body(self, name):
            return _Body(self, self._model.spec.body_id(name))

        def joint(self, name):
            return _Joint(self, self._model.spec.joint_id(name))

        def actuator(self, name):
            return _Actuator(self, next(k for k, a in enumerate(self._model.spec.actuators) if a.name == name))

    mujoco = types.ModuleType("mujoco")
    mujoco.MjModel, mujoco.MjData = MjModel, MjData
    mujoco.mj_step = lambda m, d, nstep=1: L.brb_ref_step(C.byref(m._m), C.byref(d._raw), int(nstep))
    mujoco.mj_forward = lambda m, d: L.brb_ref_forward(C.byref(m._m), C.byref(d._raw))
    mujoco.mj_resetData = lambda m, d: L.brb_ref_reset_data(C.byref(m._m), C.byref(d._raw))
    mujoco.mj_rnePostConstraint = lambda m, d: None       # fills cacc / cfrc_* only; no env code reads them
    mujoco.mjtGridPos = types.SimpleNamespace(mjGRID_TOPRIGHT=1)

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.dtype = np.asarray(low, dtype), np.asarray(high, dtype), dtype
            self.shape = self.low.shape if shape is None else tuple(shape)

    class EzPickle:
        def __init__(self, *args, **kwargs):
            self._ezpickle_args, self._ezpickle_kwargs = args, kwargs

    class _Renderer:
        viewer = None

    class MujocoEnv:
        """gymnasium.envs.mujoco.MujocoEnv, reduced to what the balance_robot env classes use."""

        def __init__(self, model_path, frame_skip, observation_space=None, render_mode=None, width=480, height=480,
                     camera_id=None, camera_name=None, default_camera_config=None, **kwargs):
            self.fullpath = model_path
            self.model = MjModel.from_xml_path(model_path)
            self.data = MjData(self.model)
            self.init_qpos = self.data.qpos.ravel().copy()
            self.init_qvel = self.data.qvel.ravel().copy()
            self.frame_skip = frame_skip
            self.observation_space = observation_space
            self.render_mode = render_mode
            self._set_action_space()
            self.np_random = QUEUES
            self.mujoco_renderer = _Renderer()

        @property
        def unwrapped(self):
            return self

        @property
        def dt(self):
            return self.model.spec.timestep * self.frame_skip

        def set_state(self, qpos, qvel):
            assert qpos.shape == (self.model.nq,) and qvel.shape == (self.model.nv,)
            self.data.qpos[:] = np.copy(qpos)
            self.data.qvel[:] = np.copy(qvel)
            mujoco.mj_forward(self.model, self.data)

        def reset(self, *, seed=None, options=None):
            mujoco.mj_resetData(self.model, self.data)
            ob = self.reset_model()
            return ob, {}

        def render(self):
            return None

        def close(self):
            pass

    def register(id, entry_point, max_episode_steps=None, reward_threshold=None, **kw):
        REGISTERED[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps, reward_threshold=reward_threshold)

    gym = types.ModuleType("gymnasium")
    gym.utils = types.ModuleType("gymnasium.utils")
    gym.utils.EzPickle = EzPickle
    gym.spaces = types.ModuleType("gymnasium.spaces")
    gym.spaces.Box = Box
    gym.envs = types.ModuleType("gymnasium.envs")
    gym.envs.mujoco = types.ModuleType("gymnasium.envs.mujoco")
    gym.envs.mujoco.MujocoEnv = MujocoEnv
    gym.envs.registration = types.ModuleType("gymnasium.envs.registration")
    gym.envs.registration.register = register
    gym.envs.registration.registry = REGISTERED
    gym.envs.registration.make = gym.envs.registration.spec = gym.envs.registration.pprint_registry = None
    for mod in (mujoco, gym, gym.utils, gym.spaces, gym.envs, gym.envs.mujoco, gym.envs.registration):
        assert mod.__name__ not in sys.modules, f"{mod.__name__} is importable here: use the real one instead of the shim"
        sys.modules[mod.__name__] = mod
    np.random.random = QUEUES.random
    sys.path.insert(0, str(REFERENCE_SRC))
    _installed = True


def make(env_id: str):
    """Instantiate the unmodified reference class registered under `env_id` (entry point from balance_robot/__init__.py).
    Env03_v2.__init__ draws attack_side_front from np.random.random(): push that uniform on QUEUES.global_u first."""
    import importlib
    install()
    import balance_robot  # noqa: F401  (runs the reference's register() calls)
    mod, cls = REGISTERED[env_id]["entry_point"].split(":")
    return getattr(importlib.import_module(mod), cls)()
