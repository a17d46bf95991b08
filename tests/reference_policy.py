"""TEST INFRASTRUCTURE: integer interpreter for the reference's trained policy (tests/golden/robot_move_policy.npz).

The fixture is the operator graph of /root/reference/src/balance_robot/envs/RobotMovePolicy.tflite (extracted by
tests/golden/make_policy_fixture.py).  RobotMoveBaseEnv._step_wheel_speeds (RobotMoveBaseEnv.py:180-210) drives the robot
with it: quantise the 6 observations to int8 (round(obs / scale) + zero_point, clipped to [-128, 127]), invoke the
interpreter, dequantise output #1 (two wheel-speed deltas), ctrl = wheel qvel + 4 * output.  The env's own step applies
exactly ctrl = qvel + 4 * action (env01_v1.py:18-23), so `act()` below is the action to feed to Env01-*.

The arithmetic restates TFLite's int8 reference kernels: FULLY_CONNECTED = int32 accumulate of (x - zx) * w + bias, then
MultiplyByQuantizedMultiplier (gemmlowp SaturatingRoundingDoublingHighMul + RoundingDivideByPOT) per output channel;
TANH = 256-entry look-up table built from float tanh.  Only the operators on the path to output #1 are evaluated.
"""
from __future__ import annotations

import math
import pathlib

import numpy as np
import torch

FIXTURE = pathlib.Path(__file__).resolve().parent / "golden" / "robot_move_policy.npz"
OP_FULLY_CONNECTED, OP_TANH = 9, 28


def _quantize_multiplier(m: float):
    """TFLite QuantizeMultiplier: m = q * 2^shift with q in [0.5, 1) as a Q31 integer."""
    if m == 0.0:
        return 0, 0
    q, shift = math.frexp(m)
    qf = int(round(q * (1 << 31)))
    if qf == (1 << 31):
        qf //= 2
        shift += 1
    return qf, shift


def _mul_by_quantized_multiplier(acc: torch.Tensor, qf: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    """acc int64 [N, C] (int32 range), qf / shift int64 [C]."""
    left = torch.clamp(shift, min=0)
    right = torch.clamp(-shift, min=0)
    ab = (acc << left) * qf
    nudge = torch.where(ab >= 0, torch.full_like(ab, 1 << 30), torch.full_like(ab, 1 - (1 << 30)))
    s = ab + nudge
    hi = torch.where(s >= 0, s >> 31, -((-s) >> 31))          # C++ division truncates toward zero
    mask = (torch.ones_like(right) << right) - 1
    rem = hi & mask
    thr = (mask >> 1) + (hi < 0).to(torch.int64)
    return (hi >> right) + (rem > thr).to(torch.int64)


class RobotMovePolicy:
    def __init__(self, device="cpu"):
        z = np.load(FIXTURE)
        self.device = torch.device(device)
        self.inp = int(z["inputs"][0])
        self.out = int(z["outputs"][1])            # "the second output is the one that includes the actions" (RobotMoveBaseEnv.py:93-96)
        self.scale = {k: z[f"t{k}_scale"].astype(np.float64) for k in range(int(z["n_tensors"]))}
        self.zero = {k: z[f"t{k}_zero"] for k in range(int(z["n_tensors"]))}
        self.in_scale, self.in_zero = float(z[f"t{self.inp}_scale"][0]), int(z[f"t{self.inp}_zero"][0])   # float32 scale, as get_input_details() reports it
        self.out_scale, self.out_zero = np.float32(z[f"t{self.out}_scale"][0]), int(z[f"t{self.out}_zero"][0])
        # keep only the operators the action output depends on
        ops = [(int(z[f"op{k}_code"]), [int(v) for v in z[f"op{k}_in"]], [int(v) for v in z[f"op{k}_out"]]) for k in range(int(z["n_ops"]))]
        need, keep = {self.out}, []
        for code, ins, outs in reversed(ops):
            if need & set(outs):
                keep.append((code, ins, outs))
                need |= {t for t in ins if f"t{t}_data" not in z}
        self.ops = []
        for code, ins, outs in reversed(keep):
            if code == OP_FULLY_CONNECTED:
                x, w, b = ins
                mult = [_quantize_multiplier(float(self.scale[x][0]) * float(sw) / float(self.scale[outs[0]][0])) for sw in
                        np.broadcast_to(self.scale[w], (z[f"t{w}_data"].shape[0],))]
                self.ops.append(("fc", x, outs[0],
                                 torch.tensor(z[f"t{w}_data"].astype(np.int64), device=self.device),
                                 torch.tensor(z[f"t{b}_data"].astype(np.int64), device=self.device),
                                 torch.tensor([m[0] for m in mult], dtype=torch.int64, device=self.device),
                                 torch.tensor([m[1] for m in mult], dtype=torch.int64, device=self.device),
                                 int(self.zero[x][0]), int(self.zero[outs[0]][0])))
            elif code == OP_TANH:
                x, o = ins[0], outs[0]
                v = np.arange(-128, 128)
                t = np.tanh((np.float32(self.scale[x][0]) * (v - int(self.zero[x][0])).astype(np.float32)).astype(np.float32))
                lut = np.clip(np.round(t.astype(np.float32) * np.float32(1.0 / np.float32(self.scale[o][0]))) + int(self.zero[o][0]), -128, 127)
                self.ops.append(("tanh", x, o, torch.tensor(lut.astype(np.int64), device=self.device)))
            else:
                raise NotImplementedError(f"TFLite builtin operator {code} on the action path")

    def quantize_obs(self, obs: torch.Tensor) -> torch.Tensor:
        q = torch.round(obs.to(torch.float64) / self.in_scale) + self.in_zero      # np.round = round-half-even = torch.round
        return torch.clamp(q, -128, 127).to(torch.int64)

    def act(self, obs: torch.Tensor) -> torch.Tensor:
        """obs float32 [N, 6] -> action float32 [N, 2] (the dequantised interpreter output)."""
        val = {self.inp: self.quantize_obs(obs.to(self.device))}
        for op in self.ops:
            if op[0] == "fc":
                _, x, o, w, b, qf, sh, zx, zo = op
                acc = (val[x] - zx).to(torch.float64) @ w.to(torch.float64).T        # |acc| < 2^24: exact in fp64
                acc = acc.to(torch.int64) + b
                val[o] = torch.clamp(_mul_by_quantized_multiplier(acc, qf, sh) + zo, -128, 127)
            else:
                _, x, o, lut = op
                val[o] = lut[val[x] + 128]
        q = val[self.out]
        return (float(self.out_scale) * (q - self.out_zero).to(torch.float32)).to(torch.float32)


def dequantised_layers(z=None):
    """[(W, b)] of the three FULLY_CONNECTED layers on the action path as float arrays (weight = int8 x per-channel scale,
    bias = int32 x its scale): the fp32 policy the int8 graph was quantised from, up to the rounding of the weights."""
    z = np.load(FIXTURE) if z is None else z
    pol = RobotMovePolicy()
    out = []
    ops = [(int(z[f"op{k}_code"]), [int(v) for v in z[f"op{k}_in"]], [int(v) for v in z[f"op{k}_out"]]) for k in range(int(z["n_ops"]))]
    need, keep = {pol.out}, []
    for code, ins, outs in reversed(ops):
        if need & set(outs):
            keep.append((code, ins, outs))
            need |= {t for t in ins if f"t{t}_data" not in z}
    for code, ins, outs in reversed(keep):
        if code == OP_FULLY_CONNECTED:
            x, w, b = ins
            W = z[f"t{w}_data"].astype(np.float64) * np.broadcast_to(z[f"t{w}_scale"].astype(np.float64), (z[f"t{w}_data"].shape[0],))[:, None]
            B = z[f"t{b}_data"].astype(np.float64) * np.broadcast_to(z[f"t{b}_scale"].astype(np.float64), z[f"t{b}_data"].shape)
            out.append((W, B))
    assert len(out) == 3
    return out
