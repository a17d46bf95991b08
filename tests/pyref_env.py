"""Pure-Python transliteration of the reference task logic (numpy + scipy), used as an independent check of
BOTH the C oracle (oracle/brb_ref_env.c) and the CUDA path's fp64 task logic.

It follows reference RobotBaseEnv.py:127-246, env01_v1.py:15-58, env01_v2.py:16-71, env01_v3.py:13-96 statement by
statement, but reads its physics from a plain `Sim` record (xquat, qvel, time) instead of mujoco.MjData, and takes
its random draws from an injected queue instead of np.random / self.np_random.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.spatial.transform import Rotation

PITCH_MAX, PITCH_DOT_MAX, WHEEL_SPEED_MAX, WHEEL_SPEED_DELTA_MAX, YAW_MAX = 0.25, 1, 170.0, 4.0, 45.0


class Sim:
    def __init__(self):
        self.xquat = np.array([1.0, 0, 0, 0])
        self.qvel = np.zeros(8)
        self.time = 0.0


class PyRefEnv:
    def __init__(self, kind: str):
        assert kind in ("Env01-v1", "Env01-v2", "Env01-v3")
        self.kind = kind
        self.sim = Sim()
        self.last_time = None
        self.last_pitch = None
        self.target_wheel_speed = 0.0
        self.target_yaw = 0.0
        self.delay_target_speed = 0.0
        self.pitch_offset = 0.0
        self.draws = []

    def _u(self):
        return self.draws.pop(0)

    # --- getters
    def _base_pitch(self):
        quat = self.sim.xquat
        if quat[0] == 0:
            return 0
        rotation = Rotation.from_quat([quat[1], quat[2], quat[3], quat[0]])
        return rotation.as_euler('xyz', degrees=False)[0]

    def get_pitch(self):
        p = self._base_pitch()
        if self.kind == "Env01-v2":
            p += (self._u() - 0.5) * 0.05
        elif self.kind == "Env01-v3":
            p = p + self.pitch_offset
        return p

    def get_pitch_dot_alt(self):
        pitch = self.get_pitch()
        ts = self.sim.time
        pitch_dot = 0
        if self.last_time is not None and self.last_pitch is not None:
            dt = ts - self.last_time
            if dt > 0.0:
                pitch_dot = (pitch - self.last_pitch) / dt
        self.last_time = ts
        self.last_pitch = pitch
        return pitch_dot

    def get_wheel_velocities(self):
        return self.sim.qvel[6], self.sim.qvel[7]

    def get_wheel_yaw(self):
        vel_l, vel_r = self.get_wheel_velocities()
        return vel_l - (-1 * vel_r)

    def get_wheel_speed(self):
        vel_l, vel_r = self.get_wheel_velocities()
        return (vel_l + (-1 * vel_r)) / 2

    def get_yaw_dot(self):
        return self.sim.qvel[5]

    # --- reward / obs
    def _get_reward(self):
        if self.kind == "Env01-v3":
            reward = 0.6
            pitch = self.get_pitch()
            wheel_speed = self.get_wheel_speed()
            dv = self.target_wheel_speed - wheel_speed
            reward -= (abs(pitch) * 0.05)
            MAX_DV = 40.0
            max_dv = np.clip(dv, -MAX_DV, MAX_DV)
            dv_n = max_dv / MAX_DV
            dv_s = abs(dv_n)
            reward -= (0.15 * dv_s)
            if self.target_wheel_speed > 0 and self.target_wheel_speed > wheel_speed:
                reward += (-1.0 * pitch) * 10.0 * dv_s
            elif self.target_wheel_speed < 0 and self.target_wheel_speed < wheel_speed:
                reward += (1.0 * pitch) * 10.0 * dv_s
            elif self.target_wheel_speed > 0 and self.target_wheel_speed < wheel_speed:
                reward += (1.0 * pitch) * 10.0 * dv_s
            elif self.target_wheel_speed < 0 and self.target_wheel_speed > wheel_speed:
                reward += (-1.0 * pitch) * 10.0 * dv_s
            dyd = self.target_yaw - self.get_wheel_yaw()
            reward -= (0.007 * abs(dyd))
            return reward
        reward = 1.0
        vel_l, vel_r = self.sim.qvel[6], self.sim.qvel[7]
        average_wheel_speed = (vel_l * -1 + vel_r) / 2.0
        dv = 0 - average_wheel_speed
        dyd = 0 - self.get_yaw_dot()
        reward -= (0.025 * abs(dyd))
        pitch = self.get_pitch()
        reward -= (abs(pitch))
        reward += pitch * dv * 0.5
        return reward

    def _get_obs(self):
        pitch = self.get_pitch()
        pitch_dot = self.get_pitch_dot_alt()
        wheel_vel_l, wheel_vel_r = self.get_wheel_velocities()
        return np.array([
            pitch / PITCH_MAX,
            pitch_dot / PITCH_DOT_MAX,
            wheel_vel_l / WHEEL_SPEED_MAX * 4,
            wheel_vel_r / WHEEL_SPEED_MAX * 4,
            (self.target_wheel_speed - self.get_wheel_speed()) / WHEEL_SPEED_MAX * 4,
            (self.target_yaw - self.get_wheel_yaw()) / YAW_MAX * 3,
        ], dtype=np.float32).ravel()

    # --- step split in the two halves around mj_step
    def pre_step(self, a):
        """returns (reward, ctrl) — everything Env01.step does before mujoco.mj_step."""
        if self.kind == "Env01-v3":
            t = self.sim.time
            if t > 5.5:
                self.target_wheel_speed = 3.0 * self.delay_target_speed
            elif t > 4.5:
                self.target_wheel_speed = 2.0 * self.delay_target_speed
            elif t > 3.0:
                self.target_wheel_speed = -1.0 * self.delay_target_speed
            elif t > 1.0:
                self.target_wheel_speed = self.delay_target_speed
        reward = self._get_reward()
        vel_l = self.sim.qvel[6] + a[0] * WHEEL_SPEED_DELTA_MAX
        vel_r = self.sim.qvel[7] + a[1] * WHEEL_SPEED_DELTA_MAX
        return reward, (vel_l, vel_r)

    def post_step(self):
        """returns (obs, terminated) — everything after mj_step."""
        terminated = np.abs(self.get_pitch()) > (50 * math.pi / 180)
        ob = self._get_obs()
        return ob, bool(terminated)

    def reset_draw_qpos(self, qpos0):
        """reset_model up to set_state: returns the qpos it would set.  Draw order = reference call order."""
        if self.kind == "Env01-v3":
            self.target_wheel_speed = 0
            self.target_yaw = 0
            self.delay_target_speed = -10.0 + (10 - -10.0) * self._u()
            if self.delay_target_speed > 0:
                self.delay_target_speed += 10
            else:
                self.delay_target_speed -= 10
            self.pitch_offset = -0.0349066 + (0.0349066 - -0.0349066) * self._u()
        qpos = qpos0 + np.array([-0.01 + (0.01 - -0.01) * self._u() for _ in range(9)])
        qpos[2] = 0
        x_rot = (self._u() - 0.5) * 2 * math.pi
        if self.kind == "Env01-v2":
            y_rot = (self._u() - 0.5) * 0.2
            z_rot = (self._u() - 0.5) * 2.0
        else:
            y_rot = (self._u() - 0.5) * 0.4
            z_rot = (self._u() - 0.5) * 0.4
        rotation = Rotation.from_euler('xyz', [x_rot, y_rot, z_rot])
        qpos[3:7] = rotation.as_quat()
        return qpos


def reference_order_reset_draws(kind: str, u_reset):
    """Map the slot layout (oracle/brb_ref_env.c header) to the reference's sequential draw order."""
    u = list(u_reset)
    if kind == "Env01-v3":
        return [u[12], u[13]] + u[0:12]
    if kind == "Env01-v2":
        return u[0:12] + [u[12], u[13]]
    return u[0:12]
