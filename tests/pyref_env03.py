"""Pure-Python transliteration of the reference's Env03 / Env03_v2 classes (envs/env03_v1.py:17-113, envs/env03_v2.py:14-59)
running on the oracle's PHYSICS primitives (brb_ref_reset_data / brb_ref_forward / brb_ref_step stand in for
mj_resetData / mj_forward / mj_step).  Used to pin the C env-level oracle for Env03-v2: both must produce identical
trajectories bit for bit given the same injected draws."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
from scipy.spatial.transform import Rotation

from oracle import ref
from pyref_env import PyRefEnv


class PyRefEnv03(PyRefEnv):
    def __init__(self, model, attack_side_front: bool):
        super().__init__("Env01-v1")          # RobotBaseEnv getters / reward / obs without noise or offsets
        self.m, self.L = model, ref.lib()
        self.d = ref.RefData()
        self.block_delay_time_start = None
        self.block_delay = 0.5
        self.attack_side_front = attack_side_front

    def _sync(self):
        self.sim.xquat = ref.arr(self.d.xquat)[1].copy()
        self.sim.qvel = ref.arr(self.d.qvel, 14).copy()
        self.sim.time = self.d.time

    def get_yaw(self):
        quat = ref.arr(self.d.xquat)[1]
        if quat[0] == 0:
            return 0
        return Rotation.from_quat([quat[1], quat[2], quat[3], quat[0]]).as_euler('xyz', degrees=False)[2]

    def set_block_pos_vel(self):
        robot_pos = ref.arr(self.d.xpos)[1]
        block_attack_angle = -self.get_yaw()
        if not self.attack_side_front:
            block_attack_angle += math.pi
        block_x_pos = 0.3 * math.sin(block_attack_angle) + robot_pos[0]
        block_y_pos = 0.3 * math.cos(block_attack_angle) + robot_pos[1]
        block_pos = np.array([block_x_pos, block_y_pos, 0.15])
        block_target_pos = np.array([(self._u() - 0.5) * 0.02 + robot_pos[0], 0 + robot_pos[1], self._u() * 0.025 + 0.13])
        block_vel_vector = block_target_pos - block_pos
        block_vel_vector = 7.5 * (block_vel_vector / np.linalg.norm(block_vel_vector))
        x_rot = self._u() * 2 * math.pi
        y_rot = self._u() * 2 * math.pi
        z_rot = self._u() * 2 * math.pi
        block_rotation = Rotation.from_euler('xyz', [x_rot, y_rot, z_rot])
        for k in range(3):
            self.d.qpos[9 + k] = block_pos[k]
            self.d.qvel[8 + k] = block_vel_vector[k]
        for k, v in enumerate(block_rotation.as_quat()):
            self.d.qpos[12 + k] = v

    def reset(self, draws):
        self.draws = list(draws)
        self.L.brb_ref_reset_data(C.byref(self.m), C.byref(self.d))
        qpos = ref.arr(self.m.qpos0, 16) + np.array([-0.01 + (0.01 - -0.01) * self._u() for _ in range(16)])
        qpos[2] = 0
        x_rot = (self._u() - 0.5) * 2 * math.pi
        y_rot = (self._u() - 0.5) * 0.4
        z_rot = (self._u() - 0.5) * 0.4
        qpos[3:7] = Rotation.from_euler('xyz', [x_rot, y_rot, z_rot]).as_quat()
        for k in range(16):
            self.d.qpos[k] = qpos[k]
        for k in range(14):
            self.d.qvel[k] = 0.0
        self.L.brb_ref_forward(C.byref(self.m), C.byref(self.d))
        self.set_block_pos_vel()
        self.block_delay_time_start = None
        self._sync()
        return self._get_obs()

    def step(self, a, draws):
        self.draws = list(draws)
        self._sync()
        reward = self._get_reward()
        # data.joint(...).qvel[0] is a numpy float64 scalar in the reference (float64 + float32 -> float64)
        self.d.ctrl[0] = np.float64(self.d.qvel[6]) + a[0] * 4.0
        self.d.ctrl[1] = np.float64(self.d.qvel[7]) + a[1] * 4.0
        self.L.brb_ref_step(C.byref(self.m), C.byref(self.d), 250)
        block_vel_vec = np.array(ref.arr(self.d.qvel, 14)[8:11])
        if np.linalg.norm(block_vel_vec) < 0.1 and self.block_delay_time_start is None:
            self.d.qpos[9], self.d.qpos[10], self.d.qpos[11] = 10, 10, 0
            self.block_delay_time_start = self.d.time
        if self.block_delay_time_start is not None and (self.d.time - self.block_delay_time_start) > self.block_delay:
            self.set_block_pos_vel()
            self.block_delay_time_start = None
        self._sync()
        terminated = np.abs(self.get_pitch()) > (50 * math.pi / 180)
        ob = self._get_obs()
        return ob, reward, bool(terminated)
