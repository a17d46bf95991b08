"""Extracts the reference's own trained balance/move policy into a small fixture.

Source: /root/reference/src/balance_robot/envs/RobotMovePolicy.tflite (int8-quantised PPO MlpPolicy, trained by the
reference author against real MuJoCo and loaded by RobotMoveBaseEnv.py:81-98 with tf.lite.Interpreter).  It is the only
artefact in the reference that was produced by the real `mj_step`, so running it on this repo's physics is the one
behavioural pin against MuJoCo that is available offline (tests/test_reference_policy.py, tests/test_gpu_reference_policy.py).

TensorFlow / flatbuffers are not installed, so the flatbuffer is walked by hand (TFLite schema v3: Model.subgraphs[0]
.tensors / .operators, Tensor.quantization.{scale, zero_point}, Buffer.data).  Output: robot_move_policy.npz holding, per
tensor, shape / dtype / scale / zero_point / constant data, and the operator list (builtin code, inputs, outputs).

Run here (needs /root/reference):  python tests/golden/make_policy_fixture.py
"""
import pathlib
import struct
import sys

import numpy as np

SRC = pathlib.Path("/root/reference/src/balance_robot/envs/RobotMovePolicy.tflite")
OUT = pathlib.Path(__file__).resolve().parent / "robot_move_policy.npz"


class FlatBuf:
    def __init__(self, b): self.b = b
    def u32(self, p): return struct.unpack_from("<I", self.b, p)[0]
    def i32(self, p): return struct.unpack_from("<i", self.b, p)[0]
    def u16(self, p): return struct.unpack_from("<H", self.b, p)[0]

    def field(self, t, i):                       # absolute position of field i of table t, 0 if absent
        vt = t - self.i32(t)
        o = 4 + 2 * i
        if o >= self.u16(vt): return 0
        f = self.u16(vt + o)
        return t + f if f else 0

    def vec(self, t, i):
        p = self.field(t, i)
        if not p: return 0, 0
        v = p + self.u32(p)
        return v + 4, self.u32(v)

    def table(self, t, i):
        p = self.field(t, i)
        return p + self.u32(p) if p else 0

    def tables(self, t, i):
        st, n = self.vec(t, i)
        return [st + 4 * k + self.u32(st + 4 * k) for k in range(n)]

    def string(self, t, i):
        st, n = self.vec(t, i)
        return self.b[st:st + n].decode() if st else ""

    def arr(self, t, i, dt):
        st, n = self.vec(t, i)
        return np.frombuffer(self.b, dtype=dt, count=n, offset=st).copy() if st else np.zeros(0, dt)

    def scalar(self, t, i, fmt, default=0):
        p = self.field(t, i)
        return struct.unpack_from(fmt, self.b, p)[0] if p else default


TENSOR_TYPES = {0: "float32", 2: "int32", 9: "int8"}


def main():
    f = FlatBuf(SRC.read_bytes())
    m = f.u32(0)
    assert f.scalar(m, 0, "<I") == 3, "TFLite schema version"
    codes = [max(f.scalar(t, 0, "<b"), f.scalar(t, 3, "<i")) for t in f.tables(m, 1)]
    bufs = f.tables(m, 4)
    sg = f.tables(m, 2)[0]
    out = {"inputs": f.arr(sg, 1, "<i4"), "outputs": f.arr(sg, 2, "<i4")}
    tensors = f.tables(sg, 0)
    out["n_tensors"] = np.int64(len(tensors))
    for k, t in enumerate(tensors):
        q = f.table(t, 4)
        dt = TENSOR_TYPES[f.scalar(t, 1, "<b")]
        shape = f.arr(t, 0, "<i4")
        out[f"t{k}_shape"] = shape
        out[f"t{k}_dtype"] = np.array(dt)
        out[f"t{k}_scale"] = f.arr(q, 2, "<f4") if q else np.zeros(0, "<f4")
        out[f"t{k}_zero"] = f.arr(q, 3, "<i8") if q else np.zeros(0, "<i8")
        st, n = f.vec(bufs[f.scalar(t, 2, "<I")], 0)
        if n:
            out[f"t{k}_data"] = np.frombuffer(f.b, dtype=dt, count=n // np.dtype(dt).itemsize, offset=st).reshape(shape).copy()
    ops = f.tables(sg, 3)
    out["n_ops"] = np.int64(len(ops))
    for k, o in enumerate(ops):
        out[f"op{k}_code"] = np.int64(codes[f.scalar(o, 0, "<I")])
        out[f"op{k}_in"] = f.arr(o, 1, "<i4")
        out[f"op{k}_out"] = f.arr(o, 2, "<i4")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes;", len(tensors), "tensors,", len(ops), "operators")


if __name__ == "__main__":
    sys.exit(main())
