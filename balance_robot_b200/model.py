"""ModelSpec -> per-model constant block for the CUDA env-step kernels.

The kernels integrate the robot in the CHASSIS frame, where the joint-space inertia is constant
(both wheels are axisymmetric about their hinge axes, so M(q) = T' M_b T with
T = blockdiag(R', I3, I2)).  This module derives, in fp64 numpy, everything that closed form needs:

  * composite mass / COM / inertia about the chassis origin, wheel axial inertia -> M_b, M_b^-1
  * body_invweight0 (mj_setConst [third party]: trace of J M^-1 J' / 3 at qpos0) -> contact R, D
  * contact reference parameters K, B, impedance for the wheel-floor pairs
  * actuator / damping constants (reference envs/robot-02.xml:11,16,23-24)

and checks that the parsed model really belongs to the class the kernel hard-codes (one free chassis,
two hinge wheels on the +-x axis, z-axis COM, diagonal composite inertia, plane floor with normal +z,
flat impedance).  Anything else raises `UnsupportedModel` — there is no generic fallback.

The layout of `BrbModelConsts` must match include/brb.h.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import Dict

import numpy as np

from . import mjcf
from .mjcf import GEOM_BOX, GEOM_CYLINDER, GEOM_PLANE, JNT_FREE, JNT_HINGE, ModelSpec, quat_to_mat

MINVAL = 1e-15


class UnsupportedModel(ValueError):
    pass


class BrbModelConsts(C.Structure):
    """Mirror of `struct BrbModelConsts` in include/brb.h (all fp32 except the step-time table scale)."""
    _fields_ = [
        ("h", C.c_float), ("grav", C.c_float), ("mass", C.c_float), ("mcz", C.c_float),
        ("Ixx", C.c_float), ("Iyy", C.c_float), ("Izz", C.c_float), ("Ia", C.c_float),
        # M_b^-1: {ux,wy} block (3), uz, wz, {uy,wx,sL,sR} block (10, row-major upper triangle)
        ("minv_xy", C.c_float * 3), ("minv_uz", C.c_float), ("minv_wz", C.c_float), ("minv_blk", C.c_float * 10),
        ("ox", C.c_float), ("oz", C.c_float), ("rad", C.c_float), ("hl", C.c_float), ("zfloor", C.c_float), ("zfloor_lo", C.c_float),
        ("damping", C.c_float), ("kv", C.c_float),
        ("ctrl_lo", C.c_float), ("ctrl_hi", C.c_float), ("frc_lo", C.c_float), ("frc_hi", C.c_float),
        ("mu", C.c_float), ("D", C.c_float), ("Kimp", C.c_float), ("Bdamp", C.c_float),
        # implicitfast: a+ = a - Wm (Cinv + G)^-1 [aL; aR]
        ("impl_W", C.c_float * 8), ("impl_G", C.c_float * 3), ("impl_cinv_full", C.c_float), ("impl_cinv_damp", C.c_float),
        # (Cinv + G)^-1 as (i00, i01, i11) for the four servo clamp states: index = clampL + 2 clampR
        ("impl_Kinv", C.c_float * 3 * 4),
        ("chassis_half", C.c_float * 3), ("chassis_pos", C.c_float * 3),
        ("frame_skip", C.c_int), ("max_episode_steps", C.c_int), ("env_kind", C.c_int), ("flags", C.c_int),
        ("pp", C.c_float * 8 * 3),
        ("blk_half", C.c_float), ("blk_mass", C.c_float), ("blk_inertia", C.c_float), ("blk_radius", C.c_float),
        ("chassis_radius", C.c_float),
        ("geo_lo", C.c_float * 4),
        ("nq", C.c_int), ("nv", C.c_int), ("wb_D1", C.c_float),
    ]


@dataclasses.dataclass
class RobotModel:
    """fp64 view of the compiled model (tests compare these against the oracle)."""
    spec: ModelSpec
    M_b: np.ndarray            # 8x8 body-frame inertia (dof order ux,uy,uz,wx,wy,wz,sL,sR)
    M_b_inv: np.ndarray
    mass: float
    com_z: float
    inertia_origin: np.ndarray  # 3 (diagonal)
    wheel_axial_inertia: float
    invweight0: Dict[str, tuple]
    meaninertia: float
    contact: Dict[str, float]   # mu, imp, K, B, diag_approx, R, D
    consts: BrbModelConsts
    time_table: np.ndarray      # time_table[k] = MuJoCo's data.time after k env steps (sequential fp64 adds)


def _skew(v):
    x, y, z = v
    return np.array([[0, -z, y], [z, 0, -x], [-y, x, 0.0]])


def compile_model(spec: ModelSpec, env_kind: int, max_episode_steps: int, frame_skip: int = 250,
                  actderiv_skip_clamped: bool = True, truncate_unsupported: bool = False, wheel_block: bool = False) -> RobotModel:
    # ---- class checks ------------------------------------------------------------------------
    free_roots = [b for b in range(1, len(spec.bodies)) if spec.bodies[b].joint >= 0
                  and spec.joints[spec.bodies[b].joint].type == JNT_FREE]
    if not free_roots:
        raise UnsupportedModel("no free-joint chassis")
    chassis = free_roots[0]
    block = None
    if len(free_roots) > 1:
        if env_kind != 3 or len(free_roots) != 2:
            raise UnsupportedModel("a second free body is only supported as the Env03-v2 block")
        block = free_roots[1]
    if spec.joints[spec.bodies[chassis].joint].dofadr != 0:
        raise UnsupportedModel("chassis free joint must own dofs 0..5")
    wheels = [b for b in range(1, len(spec.bodies)) if spec.bodies[b].parent == chassis]
    if len(wheels) != 2 or (spec.nv, spec.nq) != ((8, 9) if block is None else (14, 16)):
        raise UnsupportedModel("expected exactly two hinge wheels on the chassis")
    if env_kind == 3 and block is None:
        raise UnsupportedModel("Env03-v2 needs the block scene")
    if tuple(spec.gravity[:2]) != (0.0, 0.0) or spec.gravity[2] >= 0:
        raise UnsupportedModel("gravity must be (0, 0, -g)")
    cb = spec.bodies[chassis]
    if tuple(cb.pos) != (0, 0, 0) or tuple(cb.quat) != (1, 0, 0, 0):
        raise UnsupportedModel("chassis must start at the origin with identity orientation")

    # ---- composite rigid body about the chassis origin (wheels locked) -------------------------
    mass = cb.mass
    mc = cb.mass * np.asarray(cb.ipos)
    I_o = np.asarray(cb.inertia) - cb.mass * _skew(cb.ipos) @ _skew(cb.ipos)
    axes, offs, axial = [], [], []
    for w in wheels:
        wb = spec.bodies[w]
        jn = spec.joints[wb.joint]
        if jn.type != JNT_HINGE or tuple(jn.pos) != (0, 0, 0) or tuple(wb.quat) != (1, 0, 0, 0):
            raise UnsupportedModel("wheels must be hinges through their body origin, unrotated")
        ax = np.asarray(jn.axis)
        if abs(abs(ax[0]) - 1) > 1e-12:
            raise UnsupportedModel("wheel hinge axes must be +-x")
        if np.linalg.norm(wb.ipos) > 1e-12:
            raise UnsupportedModel("wheel COM must sit on the hinge")
        Iw = np.asarray(wb.inertia)
        # axisymmetric about the hinge axis: I = It*1 + (Ia-It) a a'
        Ia = float(ax @ Iw @ ax)
        It = (np.trace(Iw) - Ia) / 2
        if np.abs(Iw - (It * np.eye(3) + (Ia - It) * np.outer(ax, ax))).max() > 1e-12 * max(1.0, Ia):
            raise UnsupportedModel("wheel inertia must be axisymmetric about the hinge axis")
        off = np.asarray(wb.pos, float)
        mass += wb.mass
        mc = mc + wb.mass * off
        I_o = I_o + Iw - wb.mass * _skew(off) @ _skew(off)
        axes.append(ax); offs.append(off); axial.append(Ia)
    com = mc / mass
    if abs(com[0]) > 1e-12 or abs(com[1]) > 1e-12:
        raise UnsupportedModel("composite COM must lie on the chassis z axis")
    if np.abs(I_o - np.diag(np.diag(I_o))).max() > 1e-12:
        raise UnsupportedModel("composite inertia about the chassis origin must be diagonal")
    if abs(axial[0] - axial[1]) > 1e-15 or spec.joints[spec.bodies[wheels[0]].joint].damping != \
            spec.joints[spec.bodies[wheels[1]].joint].damping:
        raise UnsupportedModel("wheels must be identical")
    # wheel order: dof 6 = left (axis -x, offset -x), dof 7 = right
    sig = [float(a[0]) for a in axes]
    if sig != [-1.0, 1.0] or offs[0][0] >= 0 or offs[1][0] <= 0 or abs(offs[0][0] + offs[1][0]) > 1e-15 \
            or offs[0][1] != 0 or offs[1][1] != 0 or offs[0][2] != offs[1][2]:
        raise UnsupportedModel("expected left wheel (axis -x, at -x) then right wheel (axis +x, at +x)")
    Ia = axial[0]

    # ---- body-frame joint-space inertia --------------------------------------------------------
    M = np.zeros((8, 8))
    M[0:3, 0:3] = mass * np.eye(3)
    M[0:3, 3:6] = -mass * _skew(com)
    M[3:6, 0:3] = M[0:3, 3:6].T
    M[3:6, 3:6] = I_o
    for k in range(2):
        M[3:6, 6 + k] = Ia * axes[k]
        M[6 + k, 3:6] = Ia * axes[k]
        M[6 + k, 6 + k] = Ia
    Minv = np.linalg.inv(M)
    meaninertia = float(np.trace(M) / 8)

    # ---- body_invweight0 (translational, rotational) at qpos0 ------------------------------------
    def invweight(point, wheel=None):
        Jp = np.zeros((3, 8)); Jr = np.zeros((3, 8))
        Jp[:, 0:3] = np.eye(3); Jp[:, 3:6] = -_skew(point); Jr[:, 3:6] = np.eye(3)
        if wheel is not None:
            Jp[:, 6 + wheel] = np.cross(axes[wheel], point - offs[wheel]); Jr[:, 6 + wheel] = axes[wheel]
        return (max(MINVAL, np.trace(Jp @ Minv @ Jp.T) / 3), max(MINVAL, np.trace(Jr @ Minv @ Jr.T) / 3))
    invw = {cb.name: invweight(np.asarray(cb.ipos, float))}
    for k, w in enumerate(wheels):
        invw[spec.bodies[w].name] = invweight(offs[k], k)

    # ---- wheel-floor contact pairs ---------------------------------------------------------------
    wheel_pairs = []
    for p in spec.pairs:
        g1, g2 = spec.geoms[p.geom1], spec.geoms[p.geom2]
        if g1.type == GEOM_PLANE and g2.type == GEOM_CYLINDER and g2.body in wheels:
            wheel_pairs.append((p, g1, g2))
    if len(wheel_pairs) != 2:
        raise UnsupportedModel("expected one floor-wheel pair per wheel")
    p0, floor, cyl = wheel_pairs[0]
    for p, g1, g2 in wheel_pairs:
        if (p.friction, p.solref, p.solimp, p.margin, p.gap, p.condim) != \
                (p0.friction, p0.solref, p0.solimp, p0.margin, p0.gap, p0.condim) or g1 is not floor \
                or g2.size != cyl.size or tuple(g2.pos) != (0, 0, 0):
            raise UnsupportedModel("wheel-floor pairs must be identical")
        # cylinder axis (geom z) must be the body x axis
        if np.abs(np.abs(quat_to_mat(g2.quat)[:, 2]) - np.array([1, 0, 0])).max() > 1e-6:
            raise UnsupportedModel("wheel cylinders must be aligned with the hinge axis")
    if p0.condim != 3 or p0.margin != 0 or p0.gap != 0 or p0.friction[0] != p0.friction[1]:
        raise UnsupportedModel("wheel-floor pairs must be condim 3, zero margin/gap, isotropic friction")
    if np.abs(quat_to_mat(floor.quat) - np.eye(3)).max() > 1e-12 or floor.pos[0] != 0 or floor.pos[1] != 0:
        raise UnsupportedModel("floor must be a z-up plane")
    d0, d1 = (min(0.9999, max(0.0001, x)) for x in p0.solimp[:2])
    if d0 != d1 and block is None:
        raise UnsupportedModel("position-dependent impedance (solimp d0 != dmax) is only built into the Env03-v2 kernel")
    imp = 0.5 * (d0 + d1) if d0 == d1 else d1     # Env03-v2 evaluates imp(dist) per contact (pp[] below); this is its saturated value
    tc, dr = p0.solref
    if tc <= 0:
        raise UnsupportedModel("direct solref (negative) not supported")
    tc = max(tc, 2 * spec.timestep)      # refsafe
    K = 1.0 / max(MINVAL, d1 * d1 * tc * tc * dr * dr)
    Bd = 2.0 / max(MINVAL, d1 * tc)
    mu = p0.friction[0]
    tran = invw[spec.bodies[wheels[0]].name][0]      # world body contributes 0
    diag_approx = tran * (1 + mu * mu)
    R_first = max(MINVAL, (1 - imp) / imp * diag_approx)
    R_py = 2 * mu * mu * R_first
    contact = dict(mu=mu, imp=imp, K=K, B=Bd, diag_approx=diag_approx, R=R_py, D=1.0 / R_py)

    # ---- actuators ---------------------------------------------------------------------------------
    if spec.nu != 2:
        raise UnsupportedModel("expected two velocity actuators")
    a0, a1 = spec.actuators
    if (a0.kv, a0.gear, a0.ctrlrange, a0.forcerange) != (a1.kv, a1.gear, a1.ctrlrange, a1.forcerange) or a0.gear != 1 \
            or not (a0.ctrllimited and a0.forcelimited and a1.ctrllimited and a1.forcelimited):
        raise UnsupportedModel("expected identical ctrl- and force-limited velocity servos, gear 1")
    if [spec.joints[a.joint].dofadr for a in (a0, a1)] != [6, 7]:
        raise UnsupportedModel("actuator 0 must drive the left wheel, actuator 1 the right wheel")
    damping = spec.joints[spec.bodies[wheels[0]].joint].damping

    # ---- pack --------------------------------------------------------------------------------------
    c = BrbModelConsts()
    h = spec.timestep
    c.h, c.grav, c.mass, c.mcz = h, -spec.gravity[2], mass, mass * com[2]
    c.Ixx, c.Iyy, c.Izz, c.Ia = I_o[0, 0], I_o[1, 1], I_o[2, 2], Ia
    c.minv_xy[:] = [Minv[0, 0], Minv[0, 4], Minv[4, 4]]
    c.minv_uz, c.minv_wz = Minv[2, 2], Minv[5, 5]
    blk = [1, 3, 6, 7]
    c.minv_blk[:] = [Minv[blk[i], blk[j]] for i in range(4) for j in range(i, 4)]
    # sparsity the kernel relies on
    dense = np.zeros((8, 8), bool)
    for grp in ([0, 4], [2], [5], blk):
        dense[np.ix_(grp, grp)] = True
    if np.abs(Minv[~dense]).max() > 1e-9 * np.abs(Minv).max():
        raise UnsupportedModel("M_b^-1 does not have the expected block structure")
    c.ox, c.oz, c.rad, c.hl, c.zfloor = offs[1][0], offs[1][2], cyl.size[0], cyl.size[1], floor.pos[2]
    c.zfloor_lo = floor.pos[2] - float(np.float32(floor.pos[2]))
    c.geo_lo[:] = [v - float(np.float32(v)) for v in (offs[1][0], offs[1][2], cyl.size[0], cyl.size[1])]
    c.damping, c.kv = damping, a0.kv
    c.ctrl_lo, c.ctrl_hi, c.frc_lo, c.frc_hi = a0.ctrlrange[0], a0.ctrlrange[1], a0.forcerange[0], a0.forcerange[1]
    c.mu, c.D, c.Kimp, c.Bdamp = mu, contact["D"], K * imp, Bd
    c.impl_W[:] = [Minv[r, 6 + k] for r in blk for k in range(2)]
    c.impl_G[:] = [Minv[6, 6], Minv[6, 7], Minv[7, 7]]
    c.impl_cinv_damp = 1.0 / (h * damping) if damping > 0 else 3.0e38
    c.impl_cinv_full = 1.0 / (h * (damping + a0.kv))
    cinv = (1.0 / (h * (damping + a0.kv)), 1.0 / (h * damping) if damping > 0 else 3.0e38)     # servo active / on its forcerange (A.9)
    for ci in range(4):
        k00, k01, k11 = cinv[ci & 1] + Minv[6, 6], Minv[6, 7], cinv[ci >> 1] + Minv[7, 7]
        det = k00 * k11 - k01 * k01
        c.impl_Kinv[ci][:] = [k11 / det, -k01 / det, k00 / det]
    chassis_geoms = [g for g in spec.geoms if g.body == chassis and g.type == GEOM_BOX]
    if len(chassis_geoms) == 1:
        c.chassis_half[:] = chassis_geoms[0].size
        c.chassis_pos[:] = chassis_geoms[0].pos
    c.frame_skip, c.max_episode_steps, c.env_kind = frame_skip, max_episode_steps, env_kind
    c.nq, c.nv = spec.nq, spec.nv

    def pair_params(p, tran_sum):
        """{mu, K, B, D1, d0, d1, width, margin}: R_row = 2 mu^2 (1-imp)/imp tran (1+mu^2) (A.7) -> D = D1 imp/(1-imp)."""
        a0, a1, width, mid, power = p.solimp
        a0, a1 = (min(0.9999, max(0.0001, x)) for x in (a0, a1))
        if (mid, power) != (0.5, 2.0) and a0 != a1:
            raise UnsupportedModel("impedance curve other than midpoint 0.5 / power 2")
        if p.condim != 3 or p.gap != 0 or p.friction[0] != p.friction[1] or p.solref[0] <= 0:
            raise UnsupportedModel("pair outside the supported contact class")
        tc_ = max(p.solref[0], 2 * spec.timestep)
        Kp = 1.0 / max(MINVAL, a1 * a1 * tc_ * tc_ * p.solref[1] * p.solref[1])
        Bp = 2.0 / max(MINVAL, a1 * tc_)
        mu_ = p.friction[0]
        return [mu_, Kp, Bp, 1.0 / (2 * mu_ * mu_ * tran_sum * (1 + mu_ * mu_)), a0, a1, max(width, 1e-15), p.margin]
    c.pp[0][:] = pair_params(p0, tran)
    if block is not None:
        bb = spec.bodies[block]
        bgeoms = [g for g in spec.geoms if g.body == block]
        if len(bgeoms) != 1 or bgeoms[0].type != GEOM_BOX or len(set(bgeoms[0].size)) != 1 or tuple(bgeoms[0].pos) != (0, 0, 0) \
                or np.linalg.norm(bb.ipos) > 1e-12:
            raise UnsupportedModel("the block must be a single cube geom centred on its body")
        Ib = np.asarray(bb.inertia)
        if np.abs(Ib - Ib[0, 0] * np.eye(3)).max() > 1e-15:
            raise UnsupportedModel("block inertia must be isotropic")
        invw[bb.name] = (1.0 / bb.mass, 1.0 / Ib[0, 0])
        gid = {id(g): k for k, g in enumerate(spec.geoms)}
        fid, bid_, cid = gid[id(floor)], gid[id(bgeoms[0])], gid[id(chassis_geoms[0])]
        by_geoms = {frozenset((p.geom1, p.geom2)): p for p in spec.pairs}
        c.pp[1][:] = pair_params(by_geoms[frozenset((fid, bid_))], invw[bb.name][0])
        cbpp = pair_params(by_geoms[frozenset((cid, bid_))], invw[cb.name][0] + invw[bb.name][0])
        c.pp[2][:] = cbpp
        # wheel-block pairs: same mixed solref / solimp / margin / friction as chassis-block (checked), own diagonal approximation
        wgeoms = [g for g in spec.geoms if g.body in wheels]
        wpp = [pair_params(by_geoms[frozenset((gid[id(g)], bid_))], invw[spec.bodies[g.body].name][0] + invw[bb.name][0]) for g in wgeoms]
        if len(wpp) != 2 or any(abs(a - b) > 1e-12 * max(1.0, abs(b)) for w_ in wpp for a, b in zip(w_[:3] + w_[4:], cbpp[:3] + cbpp[4:])) \
                or abs(wpp[0][3] - wpp[1][3]) > 1e-9 * abs(wpp[0][3]):
            raise UnsupportedModel("wheel-block pairs must share the chassis-block contact parameters")
        c.wb_D1 = wpp[0][3]
        hb = bgeoms[0].size[0]
        c.blk_half, c.blk_mass, c.blk_inertia, c.blk_radius = hb, bb.mass, Ib[0, 0], hb * math.sqrt(3.0)
        c.chassis_radius = float(np.linalg.norm(chassis_geoms[0].size))
    c.flags = (1 if actderiv_skip_clamped else 0) | (2 if truncate_unsupported else 0) | (4 if wheel_block else 0)     # include/brb.h BRB_FLAG_*

    # MuJoCo accumulates data.time += h once per substep in fp64; reproduce the exact sequence
    # (np.cumsum adds sequentially, left to right, exactly like the C loop `time += h`)
    nt = max_episode_steps + 2
    tt = np.concatenate([[0.0], np.cumsum(np.full((nt - 1) * frame_skip, h))[frame_skip - 1::frame_skip]])
    return RobotModel(spec, M, Minv, mass, float(com[2]), np.diag(I_o).copy(), Ia, invw, meaninertia, contact, c, tt)
