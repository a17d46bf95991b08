"""Shared test plumbing: oracle handles, the host emulation of the kernel arithmetic, PD controller."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
EMU_DIR = ROOT / "tests" / "host_emu"
EMU_LIB = EMU_DIR / "libbrb_emu.so"
KERNEL_SRC = ROOT / "balance_robot_b200" / "csrc" / "brb_kernels.cu"

ENV_IDS = {0: "Env01-v1", 1: "Env01-v2", 2: "Env01-v3", 3: "Env03-v2"}


def build_emu() -> C.CDLL:
    deps = [EMU_DIR / "emu.cpp", EMU_DIR / "emu_shim.h", KERNEL_SRC, ROOT / "include" / "brb.h",
            ROOT / "balance_robot_b200" / "csrc" / "brb_internal.h", ROOT / "balance_robot_b200" / "csrc" / "brb_chol8.inc", ROOT / "balance_robot_b200" / "csrc" / "brb_chol6.inc",
            ROOT / "balance_robot_b200" / "csrc" / "brb_env03.cuh", ROOT / "balance_robot_b200" / "csrc" / "brb_schur6.inc"]
    if not EMU_LIB.exists() or any(d.stat().st_mtime > EMU_LIB.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-DBRB_HOST_EMU", "-I", str(EMU_DIR), "-mfma",
                        "-ffp-contract=fast", "-x", "c++", str(EMU_DIR / "emu.cpp"), "-o", str(EMU_LIB)], check=True)
    E = C.CDLL(str(EMU_LIB))
    E.emu_create.restype = C.c_void_p
    E.emu_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_ulonglong, C.c_longlong]
    for name in ("emu_reset", "emu_step", "emu_get_state", "emu_set_state", "emu_get_stats", "emu_destroy"):
        getattr(E, name).restype = None
    E.emu_reset.argtypes = [C.c_void_p] * 3
    E.emu_step.argtypes = [C.c_void_p] * 10
    E.emu_get_state.argtypes = [C.c_void_p] * 4
    E.emu_set_state.argtypes = [C.c_void_p] * 3
    E.emu_get_stats.argtypes = [C.c_void_p] * 2
    E.emu_destroy.argtypes = [C.c_void_p]
    return E


class EmuVecEnv:
    """TEST-ONLY: runs the CUDA kernel's per-env functions on the host (tests/host_emu)."""

    def __init__(self, robot_model, n, seed=0, env0=0):
        self.E = build_emu()
        self.n = n
        self.nq, self.nv = robot_model.consts.nq, robot_model.consts.nv
        tt = np.ascontiguousarray(robot_model.time_table)
        self.h = C.c_void_p(self.E.emu_create(C.addressof(robot_model.consts), tt.ctypes.data, len(tt), n, seed, env0))
        self.obs = np.zeros((n, 6), np.float32); self.rew = np.zeros(n, np.float32)
        self.done = np.zeros(n, np.uint8); self.trunc = np.zeros(n, np.uint8)
        self.tobs = np.zeros((n, 6), np.float32); self.epr = np.zeros(n, np.float32); self.epl = np.zeros(n, np.int32)

    def reset(self, replay=None):
        r = None if replay is None else np.ascontiguousarray(replay, np.float64)
        self.E.emu_reset(self.h, self.obs.ctypes.data, None if r is None else r.ctypes.data)
        return self.obs.copy()

    def step(self, act, replay=None):
        a = np.ascontiguousarray(act, np.float32)
        r = None if replay is None else np.ascontiguousarray(replay, np.float64)
        self.E.emu_step(self.h, a.ctypes.data, self.obs.ctypes.data, self.rew.ctypes.data, self.done.ctypes.data,
                        self.trunc.ctypes.data, self.tobs.ctypes.data, self.epr.ctypes.data, self.epl.ctypes.data,
                        None if r is None else r.ctypes.data)
        return self.obs.copy(), self.rew.copy(), self.done.copy(), self.trunc.copy()

    def get_state(self):
        qp = np.zeros((self.n, self.nq)); qv = np.zeros((self.n, self.nv)); xq = np.zeros((self.n, 4))
        self.E.emu_get_state(self.h, qp.ctypes.data, qv.ctypes.data, xq.ctypes.data)
        return qp, qv, xq

    def set_state(self, qpos, qvel):
        qp = np.ascontiguousarray(qpos, np.float64); qv = np.ascontiguousarray(qvel, np.float64)
        self.E.emu_set_state(self.h, qp.ctypes.data, qv.ctypes.data)

    def stats(self):
        out = (C.c_ulonglong * 12)()
        self.E.emu_get_stats(self.h, out)
        return list(out)

    def close(self):
        if self.h:
            self.E.emu_destroy(self.h)
            self.h = None


def pd_policy(obs: np.ndarray, gain_p=12.0, gain_d=0.6, gain_v=0.35) -> np.ndarray:
    """Hand-tuned stabilising controller on the observation (pitch, pitch-rate, wheel speeds):
    drive the wheels under the fall.  Left wheel axis is -x, so it gets the opposite sign (RobotBaseEnv.py:163-165)."""
    pitch = obs[:, 0] * 0.25
    pitch_dot = obs[:, 1]
    wheel_speed = -obs[:, 4] * 170.0 / 4.0      # obs[4] = (0 - wheel_speed)/170*4 for target 0
    u = gain_p * pitch + gain_d * pitch_dot + gain_v * wheel_speed * 0.034
    u = np.clip(u, -1, 1)
    return np.stack([-u, u], 1).astype(np.float32)


def state_errors(qpos, qvel, qpos_ref, qvel_ref):
    """max abs qpos error and qvel error relative to max(1, |qvel|_inf) per env."""
    eq = np.abs(qpos - qpos_ref).max(1)
    ev = np.abs(qvel - qvel_ref).max(1) / np.maximum(1.0, np.abs(qvel_ref).max(1))
    return eq, ev
