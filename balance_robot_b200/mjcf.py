"""Host-side MJCF-subset compiler for the balance-robot scenes.

Parses exactly the MJCF features the reference scenes use
(reference: envs/robot-02.xml:1-27, envs/env01_v1.xml:1-35, envs/env03_v1.xml:1-40) and produces a
`ModelSpec`: bodies in a kinematic tree (free root joints + hinge children), geoms
(plane / box / cylinder), velocity actuators and the list of geom pairs that can collide.

What the MuJoCo compiler [third party, 3.2.0, conda-environment.yaml:7] does for these files and
what is restated here:
  * `<include>`: the included file's top-level sections are spliced in place.
  * `inertiafromgeom="true"`: every `<inertial>` element is ignored; mass / inertia come from the
    body's geoms at the default density 1000.
  * explicit `<contact><pair>` entries are always tested; all other geom pairs are "dynamic" pairs,
    filtered by contype/conaffinity, same-body and parent-child (world exempt), with parameters mixed
    from the two geoms.

Nothing here touches the GPU; the output feeds `model.py` (device constant block) and, in tests, the
fp64 oracle.
"""
from __future__ import annotations

import dataclasses
import math
import pathlib
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional, Tuple

import numpy as np

ASSET_DIR = pathlib.Path(__file__).parent / "assets"

GEOM_PLANE, GEOM_BOX, GEOM_CYLINDER = 0, 6, 5  # MuJoCo mjtGeom numbering (plane=0, cylinder=5, box=6)
JNT_FREE, JNT_HINGE = 0, 3                     # mjtJoint numbering (free=0, hinge=3)

# MuJoCo defaults for geoms that take part in dynamic pairs
DEFAULT_FRICTION = (1.0, 0.005, 0.0001)
DEFAULT_SOLREF = (0.02, 1.0)
DEFAULT_SOLIMP = (0.9, 0.95, 0.001, 0.5, 2.0)
DEFAULT_DENSITY = 1000.0


class MjcfError(ValueError):
    """Raised for MJCF content outside the supported subset (fail loudly, never guess)."""


def _floats(text: Optional[str], n: Optional[int] = None, default=None) -> Tuple[float, ...]:
    if text is None:
        if default is None:
            raise MjcfError("missing required numeric attribute")
        return tuple(default)
    vals = tuple(float(t) for t in text.split())
    if n is not None and len(vals) != n:
        raise MjcfError(f"expected {n} numbers, got {text!r}")
    return vals


def quat_to_mat(q) -> np.ndarray:
    w, x, y, z = (float(v) for v in q)
    n = math.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


@dataclasses.dataclass
class Geom:
    name: str
    type: int
    body: int
    size: Tuple[float, float, float]
    pos: Tuple[float, float, float]
    quat: Tuple[float, float, float, float]
    friction: Tuple[float, float, float] = DEFAULT_FRICTION
    solref: Tuple[float, float] = DEFAULT_SOLREF
    solimp: Tuple[float, float, float, float, float] = DEFAULT_SOLIMP
    margin: float = 0.0
    gap: float = 0.0
    condim: int = 3
    contype: int = 1
    conaffinity: int = 1
    density: float = DEFAULT_DENSITY


@dataclasses.dataclass
class Joint:
    name: str
    type: int
    body: int
    axis: Tuple[float, float, float]
    pos: Tuple[float, float, float]
    damping: float
    qposadr: int = 0
    dofadr: int = 0


@dataclasses.dataclass
class Body:
    name: str
    parent: int
    pos: Tuple[float, float, float]
    quat: Tuple[float, float, float, float]
    joint: int = -1            # at most one joint per body in this subset
    mass: float = 0.0
    ipos: Tuple[float, float, float] = (0.0, 0.0, 0.0)   # COM in body frame
    inertia: Optional[np.ndarray] = None                  # 3x3 about the COM, body-frame axes


@dataclasses.dataclass
class Actuator:
    name: str
    joint: int
    kv: float
    gear: float
    ctrllimited: bool
    ctrlrange: Tuple[float, float]
    forcelimited: bool
    forcerange: Tuple[float, float]


@dataclasses.dataclass
class Pair:
    geom1: int
    geom2: int
    condim: int
    friction: Tuple[float, float, float, float, float]
    solref: Tuple[float, float]
    solimp: Tuple[float, float, float, float, float]
    margin: float
    gap: float
    explicit: bool


@dataclasses.dataclass
class ModelSpec:
    timestep: float
    gravity: Tuple[float, float, float]
    integrator: str
    bodies: List[Body]
    joints: List[Joint]
    geoms: List[Geom]
    actuators: List[Actuator]
    pairs: List[Pair]
    dropped_pairs: List[Tuple[str, str, str]]
    nq: int = 0
    nv: int = 0

    @property
    def nu(self) -> int:
        return len(self.actuators)

    def body_id(self, name: str) -> int:
        for i, b in enumerate(self.bodies):
            if b.name == name:
                return i
        raise KeyError(name)

    def geom_id(self, name: str) -> int:
        for i, g in enumerate(self.geoms):
            if g.name == name:
                return i
        raise KeyError(name)

    def joint_id(self, name: str) -> int:
        for i, j in enumerate(self.joints):
            if j.name == name:
                return i
        raise KeyError(name)

    def qpos0(self) -> np.ndarray:
        q = np.zeros(self.nq)
        for j in self.joints:
            if j.type == JNT_FREE:
                b = self.bodies[j.body]
                q[j.qposadr:j.qposadr + 3] = b.pos
                q[j.qposadr + 3:j.qposadr + 7] = b.quat
        return q


# ------------------------------------------------------------------------------------------------
def _load_tree(path: pathlib.Path) -> ET.Element:
    root = ET.parse(str(path)).getroot()
    if root.tag != "mujoco":
        raise MjcfError(f"{path}: root element must be <mujoco>")
    _expand_includes(root, path.parent)
    return root


def _expand_includes(elem: ET.Element, base: pathlib.Path) -> None:
    i = 0
    while i < len(elem):
        child = elem[i]
        if child.tag == "include":
            sub = ET.parse(str(base / child.attrib["file"])).getroot()
            _expand_includes(sub, base)
            elem.remove(child)
            for k, sc in enumerate(list(sub)):
                elem.insert(i + k, sc)
            i += len(sub)
        else:
            _expand_includes(child, base)
            i += 1


def _geom_inertia(g: Geom) -> Tuple[float, np.ndarray]:
    """mass and inertia (about the geom centre, geom-frame axes) at the geom's density."""
    if g.type == GEOM_BOX:
        a, b, c = g.size
        m = 8.0 * a * b * c * g.density
        return m, np.diag([m / 3.0 * (b * b + c * c), m / 3.0 * (a * a + c * c), m / 3.0 * (a * a + b * b)])
    if g.type == GEOM_CYLINDER:
        r, hl = g.size[0], g.size[1]
        m = math.pi * r * r * 2.0 * hl * g.density
        it = m * (3.0 * r * r + 4.0 * hl * hl) / 12.0
        return m, np.diag([it, it, 0.5 * m * r * r])
    raise MjcfError(f"geom {g.name}: no inertia rule for type {g.type}")


def parse(path) -> ModelSpec:
    path = pathlib.Path(path)
    if not path.is_absolute() and not path.exists():
        path = ASSET_DIR / path
    root = _load_tree(path)

    comp = root.find("compiler")
    angle = comp.attrib.get("angle", "degree") if comp is not None else "degree"
    inertiafromgeom = (comp.attrib.get("inertiafromgeom", "auto") if comp is not None else "auto")
    if angle != "radian":
        raise MjcfError("only compiler angle='radian' is supported")
    if inertiafromgeom != "true":
        raise MjcfError("only inertiafromgeom='true' is supported (reference scenes all set it)")
    opt = root.find("option")
    timestep = float(opt.attrib.get("timestep", 0.002)) if opt is not None else 0.002
    gravity = _floats(opt.attrib.get("gravity") if opt is not None else None, 3, (0, 0, -9.81))
    integrator = opt.attrib.get("integrator", "Euler") if opt is not None else "Euler"
    if integrator != "implicitfast":
        raise MjcfError(f"integrator {integrator!r} not supported (reference uses implicitfast)")
    for tag in ("default", "equality", "tendon", "sensor", "keyframe"):
        if root.find(tag) is not None:
            raise MjcfError(f"<{tag}> is outside the supported MJCF subset")

    bodies: List[Body] = [Body("world", -1, (0, 0, 0), (1, 0, 0, 0))]
    joints: List[Joint] = []
    geoms: List[Geom] = []
    unnamed = [0]

    def add_geom(e: ET.Element, body: int) -> None:
        tname = e.attrib.get("type", "sphere")
        types = {"plane": GEOM_PLANE, "box": GEOM_BOX, "cylinder": GEOM_CYLINDER}
        if tname not in types:
            raise MjcfError(f"geom type {tname!r} not supported")
        size = _floats(e.attrib.get("size"), None)
        size = tuple(size) + (0.0,) * (3 - len(size))
        name = e.attrib.get("name")
        if name is None:
            name = f"_geom{unnamed[0]}"
            unnamed[0] += 1
        fr = _floats(e.attrib.get("friction"), None, DEFAULT_FRICTION)
        fr = tuple(fr) + DEFAULT_FRICTION[len(fr):]
        si = _floats(e.attrib.get("solimp"), None, DEFAULT_SOLIMP)
        si = tuple(si) + DEFAULT_SOLIMP[len(si):]
        geoms.append(Geom(
            name=name, type=types[tname], body=body, size=size[:3],
            pos=_floats(e.attrib.get("pos"), 3, (0, 0, 0)),
            quat=_floats(e.attrib.get("quat"), 4, (1, 0, 0, 0)),
            friction=fr, solref=_floats(e.attrib.get("solref"), 2, DEFAULT_SOLREF), solimp=si,
            margin=float(e.attrib.get("margin", 0.0)), gap=float(e.attrib.get("gap", 0.0)),
            condim=int(e.attrib.get("condim", 3)), contype=int(e.attrib.get("contype", 1)),
            conaffinity=int(e.attrib.get("conaffinity", 1)),
            density=float(e.attrib.get("density", DEFAULT_DENSITY))))

    def add_body(e: ET.Element, parent: int) -> None:
        bid = len(bodies)
        bodies.append(Body(e.attrib.get("name", f"_body{bid}"), parent,
                           _floats(e.attrib.get("pos"), 3, (0, 0, 0)),
                           _floats(e.attrib.get("quat"), 4, (1, 0, 0, 0))))
        jl = e.findall("joint") + e.findall("freejoint")
        if len(jl) > 1:
            raise MjcfError("at most one joint per body is supported")
        for je in jl:
            jt = "free" if je.tag == "freejoint" else je.attrib.get("type", "hinge")
            if jt not in ("free", "hinge"):
                raise MjcfError(f"joint type {jt!r} not supported")
            if jt == "free" and parent != 0:
                raise MjcfError("free joints must be attached to children of the world body")
            if any(k in je.attrib for k in ("armature", "frictionloss", "stiffness", "range", "limited")):
                raise MjcfError("joint armature/frictionloss/stiffness/limits are not supported")
            bodies[bid].joint = len(joints)
            joints.append(Joint(je.attrib.get("name", f"_joint{len(joints)}"),
                                JNT_FREE if jt == "free" else JNT_HINGE, bid,
                                _floats(je.attrib.get("axis"), 3, (0, 0, 1)),
                                _floats(je.attrib.get("pos"), 3, (0, 0, 0)),
                                float(je.attrib.get("damping", 0.0))))
        for ge in e.findall("geom"):
            add_geom(ge, bid)
        for be in e.findall("body"):
            add_body(be, bid)

    for wb in root.findall("worldbody"):
        for ge in wb.findall("geom"):
            add_geom(ge, 0)
        for be in wb.findall("body"):
            add_body(be, 0)

    # address assignment (MuJoCo orders joints/dofs depth-first in body order == our append order)
    nq = nv = 0
    for j in joints:
        j.qposadr, j.dofadr = nq, nv
        if j.type == JNT_FREE:
            nq, nv = nq + 7, nv + 6
        else:
            ax = np.asarray(j.axis, float)
            j.axis = tuple(ax / np.linalg.norm(ax))
            nq, nv = nq + 1, nv + 1

    # inertiafromgeom: body mass / COM / inertia from its geoms (Q8)
    for bid, b in enumerate(bodies):
        if bid == 0:
            continue
        gl = [g for g in geoms if g.body == bid]
        if not gl:
            raise MjcfError(f"body {b.name}: inertiafromgeom needs at least one geom")
        ms, coms, inerts = [], [], []
        for g in gl:
            m, ig = _geom_inertia(g)
            rg = quat_to_mat(g.quat)
            ms.append(m)
            coms.append(np.asarray(g.pos, float))
            inerts.append(rg @ ig @ rg.T)
        mt = float(sum(ms))
        com = sum(m * c for m, c in zip(ms, coms)) / mt
        it = np.zeros((3, 3))
        for m, c, ig in zip(ms, coms, inerts):
            d = c - com
            it += ig + m * (float(d @ d) * np.eye(3) - np.outer(d, d))
        b.mass, b.ipos, b.inertia = mt, tuple(com), it

    # actuators
    acts: List[Actuator] = []
    ae = root.find("actuator")
    if ae is not None:
        for e in ae:
            if e.tag != "velocity":
                raise MjcfError(f"actuator <{e.tag}> not supported (only <velocity>)")
            jname = e.attrib["joint"]
            jid = next(i for i, j in enumerate(joints) if j.name == jname)
            if joints[jid].type != JNT_HINGE:
                raise MjcfError("actuators must drive hinge joints")
            truthy = lambda s: str(s).lower() == "true"
            acts.append(Actuator(e.attrib.get("name", f"_act{len(acts)}"), jid, float(e.attrib.get("kv", 1.0)),
                                 _floats(e.attrib.get("gear"), None, (1.0,))[0],
                                 truthy(e.attrib.get("ctrllimited", "false")),
                                 _floats(e.attrib.get("ctrlrange"), 2, (0, 0)),
                                 truthy(e.attrib.get("forcelimited", "false")),
                                 _floats(e.attrib.get("forcerange"), 2, (0, 0))))

    spec = ModelSpec(timestep, gravity, integrator, bodies, joints, geoms, acts, [], [], nq, nv)

    # explicit pairs
    explicit = set()
    ce = root.find("contact")
    if ce is not None:
        for e in ce:
            if e.tag != "pair":
                raise MjcfError(f"<contact><{e.tag}> not supported")
            g1, g2 = spec.geom_id(e.attrib["geom1"]), spec.geom_id(e.attrib["geom2"])
            fr = _floats(e.attrib.get("friction"), None, (1, 1, 0.005, 0.0001, 0.0001))
            fr = tuple(fr) + (1.0, 1.0, 0.005, 0.0001, 0.0001)[len(fr):]
            si = _floats(e.attrib.get("solimp"), None, DEFAULT_SOLIMP)
            si = tuple(si) + DEFAULT_SOLIMP[len(si):]
            spec.pairs.append(Pair(g1, g2, int(e.attrib.get("condim", 3)), fr,
                                   _floats(e.attrib.get("solref"), 2, DEFAULT_SOLREF), si,
                                   float(e.attrib.get("margin", 0.0)), float(e.attrib.get("gap", 0.0)), True))
            explicit.add((min(g1, g2), max(g1, g2)))

    # dynamic pairs
    supported = {(GEOM_PLANE, GEOM_BOX), (GEOM_PLANE, GEOM_CYLINDER), (GEOM_BOX, GEOM_BOX),
                 (GEOM_CYLINDER, GEOM_BOX)}

    def weld_parent(b: int) -> int:
        return bodies[b].parent

    for i in range(len(geoms)):
        for k in range(i + 1, len(geoms)):
            if (i, k) in explicit:
                continue
            ga, gb = geoms[i], geoms[k]
            if ga.body == gb.body:
                continue
            if not ((ga.contype & gb.conaffinity) or (gb.contype & ga.conaffinity)):
                continue
            # parent-child filter; the world body is exempt
            if ga.body != 0 and gb.body != 0 and (weld_parent(ga.body) == gb.body or weld_parent(gb.body) == ga.body):
                spec.dropped_pairs.append((ga.name, gb.name, "parent-child filter"))
                continue
            # MuJoCo orders each pair so that the lower geom type comes first
            g1, g2 = (i, k) if ga.type <= gb.type else (k, i)
            t = (geoms[g1].type, geoms[g2].type)
            if t not in supported:
                spec.dropped_pairs.append((ga.name, gb.name, f"type combination {t} has no collider"))
                continue
            a, b = geoms[g1], geoms[g2]
            mu = tuple(max(x, y) for x, y in zip(a.friction, b.friction))
            if a.solref[0] > 0 and b.solref[0] > 0:
                solref = tuple(0.5 * (x + y) for x, y in zip(a.solref, b.solref))
            else:
                solref = tuple(min(x, y) for x, y in zip(a.solref, b.solref))
            solimp = tuple(0.5 * (x + y) for x, y in zip(a.solimp, b.solimp))
            spec.pairs.append(Pair(g1, g2, max(a.condim, b.condim), (mu[0], mu[0], mu[1], mu[2], mu[2]),
                                   solref, solimp, max(a.margin, b.margin), max(a.gap, b.gap), False))
    return spec
