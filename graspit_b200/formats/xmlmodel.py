"""GraspIt! world / robot / body / virtual-contact / eigengrasp XML readers (host side).

These are the data formats on the input side of the hot path.  The
reference parsers are Qt/Coin-bound (``World::loadFromXml`` src/world.cpp:549-797,
``Robot::loadFromXml`` src/robot.cpp:98-420, ``KinematicChain::initChainFromXml``
src/kinematicChain.cpp:148-350, ``RevoluteJoint/PrismaticJoint::initJointFromXml``
src/joint.cpp:206-370, ``Body::loadFromXml`` src/body.cpp:225-367, ``getTransform``
src/mytools.cpp:275-368, ``VirtualContact::loadFromXml`` src/contact/virtualContact.cpp:399-494,
``EigenGraspInterface::readFromFile`` src/eigenGrasp.cpp:372-478); this module reads the same
files into plain dataclasses (``Scene``) that both the CUDA engine and the test oracle consume.
"""
from __future__ import annotations

import math
import os
import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import transf as tf
from .scenegraph import load_triangles, F32

Transf = Tuple[np.ndarray, np.ndarray]


# ------------------------------------------------------------------------------------------
# data model
# ------------------------------------------------------------------------------------------
@dataclass
class BodyDesc:
    name: str
    tris: np.ndarray                      # (N,3,3) float32, body frame, mm
    kind: str = "object"                  # object | obstacle | base | link | mount
    robot: int = -1
    chain: int = -1
    link: int = -1
    youngs: float = -1.0                  # > 0  => elastic (soft contacts)
    material: str = ""
    filename: str = ""


@dataclass
class JointDesc:
    dof: int
    ratio: float
    offset: float                         # c: rad (revolute, folded into theta) / mm (prismatic)
    revolute: bool
    theta: float                          # DH theta at joint value 0 (includes c for revolute)
    d: float                              # DH d at joint value 0 (includes c for prismatic)
    a: float
    alpha: float
    min: float
    max: float
    stiffness: float = 0.0                # <springStiffness> (joint.cpp:227,319); only CompliantDOF::getStaticRatio reads it
    coupling: Optional[float] = None      # the DOF expression's ratio where `ratio` had to be replaced (CompliantDOF)


@dataclass
class ChainDesc:
    tran: Transf
    joints: List[JointDesc]
    link_bodies: List[int]                # scene body index per link
    last_joint: List[int]                 # link l takes the running product after this joint
    children: List[Tuple[int, Transf]] = field(default_factory=list)   # (robot index, offset)


@dataclass
class VirtualContactDesc:
    body: int                             # scene body index
    finger: int
    link: int
    loc: np.ndarray
    normal: np.ndarray
    frame: Transf


@dataclass
class RobotDesc:
    name: str
    base_body: int
    chains: List[ChainDesc]
    n_dof: int
    dof_type: List[str]
    dof_min: np.ndarray
    dof_max: np.ndarray
    dof_default: np.ndarray
    dof_values: np.ndarray                # world-file <dofValues> (or defaults)
    tran: Transf                          # base pose in the world
    approach: Transf
    eg: Optional[np.ndarray] = None       # (nEG, nDOF) normalised eigengrasps
    eg_origin: Optional[np.ndarray] = None
    eg_norm: Optional[np.ndarray] = None
    eg_limits: Optional[np.ndarray] = None   # (nEG,2), NaN when not predefined
    vcontacts: List[VirtualContactDesc] = field(default_factory=list)
    mount_body: int = -1
    parent: int = -1
    parent_chain: int = -1
    dof_velocity: Optional[np.ndarray] = None   # <defaultVelocity> per DOF (Hand::autoGrasp step = velocity x 0.01)

    @property
    def n_joints(self):
        return sum(len(c.joints) for c in self.chains)

    def body_list(self) -> List[int]:
        """Bodies in World::findVirtualContacts order: finger links chain-major, then the palm
        (src/world.cpp:1623-1639); mount piece last if present."""
        out = [b for c in self.chains for b in c.link_bodies] + [self.base_body]
        if self.mount_body >= 0:
            out.append(self.mount_body)
        return out


@dataclass
class Scene:
    bodies: List[BodyDesc] = field(default_factory=list)
    body_tran: List[Transf] = field(default_factory=list)     # free bodies: pose; robot bodies: pose at load
    robots: List[RobotDesc] = field(default_factory=list)
    disabled_pairs: List[Tuple[int, int]] = field(default_factory=list)   # structural (file-defined)
    touching_pairs: List[Tuple[int, int]] = field(default_factory=list)   # in contact at load (world.cpp:1080-1086)
    name: str = ""

    def add_body(self, b: BodyDesc, tran: Transf) -> int:
        self.bodies.append(b)
        self.body_tran.append(tran)
        return len(self.bodies) - 1

    def disable(self, a: int, b: int):
        if a == b:
            return
        p = (min(a, b), max(a, b))
        if p not in self.disabled_pairs:
            self.disabled_pairs.append(p)

    def free_bodies(self) -> List[int]:
        return [i for i, b in enumerate(self.bodies) if b.robot < 0]


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------
def _text(e) -> str:
    return " ".join((e.text or "").split()) if e is not None else ""


def _get_double(root, tag, default=None):
    e = root.find(tag)
    if e is None or not _text(e):
        return default
    try:
        return float(_text(e))
    except ValueError:
        return default


_FULL_RE = re.compile(r"\(\s*([^)]*)\)\s*\[\s*([^\]]*)\]")


def parse_transform(elem) -> Transf:
    """``<transform>``: children composed left to right, total = total % child (mytools.cpp:364)."""
    total = tf.IDENTITY
    if elem is None:
        return total
    for child in list(elem):
        vals = _text(child).split()
        tag = child.tag.strip()
        if tag == "translation":
            new = tf.translation([float(v) for v in vals[:3]])
        elif tag == "rotation":
            ang = float(vals[0]) * math.pi / 180.0
            axis = {"x": (1, 0, 0), "y": (0, 1, 0), "z": (0, 0, 1)}[vals[1]]
            new = tf.angle_axis(ang, axis)
        elif tag == "rotationMatrix":
            r = np.array([float(v) for v in vals[:9]])
            # Eigen's mat3(const double*) is column-major
            new = tf.from_matrix(r.reshape(3, 3).T, [0, 0, 0])
        elif tag == "fullTransform":
            m = _FULL_RE.search(child.text or "")
            q = [float(v) for v in m.group(1).split()]
            t = [float(v) for v in m.group(2).split()]
            new = tf.make(q, t)
        else:
            continue
        total = tf.compose(total, new)
    return total


def _pre_model(root) -> Optional[np.ndarray]:
    """<geometryScaling*> and <geometryOffset> as an fp32 model matrix above the file geometry.

    Coin graph order is [offset, scale, geometry] (src/body.cpp:339,362: both inserted at index 0,
    offset last => outermost), i.e. p' = offset(scale(p)); row-vector matrix = S * O."""
    sx = sy = sz = 1.0
    s = _get_double(root, "geometryScaling")
    if s is not None:
        sx = sy = sz = s
    else:
        sx = _get_double(root, "geometryScalingX", 1.0)
        sy = _get_double(root, "geometryScalingY", 1.0)
        sz = _get_double(root, "geometryScalingZ", 1.0)
    off = root.find("geometryOffset")
    if (sx, sy, sz) == (1.0, 1.0, 1.0) and off is None:
        return None
    model = np.eye(4, dtype=F32)
    if off is not None:
        q, t = parse_transform(off.find("transform"))
        # SoTransform stores fp32 rotation + translation
        R = tf.qmat(q.astype(F32).astype(np.float64)).astype(F32)
        O = np.eye(4, dtype=F32)
        O[:3, :3] = R.T            # row-vector convention
        O[3, :3] = np.asarray(t, dtype=F32)
        model = (O @ model).astype(F32)
    S = np.eye(4, dtype=F32)
    S[0, 0], S[1, 1], S[2, 2] = F32(sx), F32(sy), F32(sz)
    model = (S @ model).astype(F32)
    return model


def load_body(xml_path: str, name: str = "", kind: str = "object") -> BodyDesc:
    """Body XML -> triangles (src/body.cpp:225-367, 456-481)."""
    root = ET.parse(xml_path).getroot()
    g = root.find("geometryFile")
    if g is None:
        raise ValueError(f"{xml_path}: no <geometryFile>")
    gtype = (g.get("type") or "Inventor").strip()
    gfile = os.path.join(os.path.dirname(xml_path), _text(g))
    if gtype not in ("Inventor", "inventor"):
        raise NotImplementedError(f"{xml_path}: geometry type {gtype!r} (only Inventor/VRML handled)")
    tris = load_triangles(gfile, _pre_model(root))
    y = _get_double(root, "youngs", -1.0)
    return BodyDesc(name=name or os.path.splitext(os.path.basename(xml_path))[0], tris=tris, kind=kind,
                    youngs=y if y is not None else -1.0, material=_text(root.find("material")),
                    filename=xml_path)


_DOF_EXPR = re.compile(r"^d(\d+)(?:\*([0-9.\-]+))?(?:\+(.*))?$")


def _parse_joint(e) -> JointDesc:
    """src/joint.cpp:206-370."""
    jtype = (e.get("type") or "").strip()
    rev = jtype == "Revolute"
    expr = _text(e.find("theta" if rev else "d")).replace(" ", "")
    m = _DOF_EXPR.match(expr)
    if not m:
        raise ValueError(f"cannot parse joint DOF expression {expr!r}")
    dof = int(m.group(1))
    ratio = float(m.group(2)) if m.group(2) else 1.0
    c = float(m.group(3)) if m.group(3) else 0.0
    a = _get_double(e, "a", 0.0)
    alpha = _get_double(e, "alpha", 0.0) * math.pi / 180.0
    mn = _get_double(e, "minValue", 0.0)
    mx = _get_double(e, "maxValue", 0.0)
    if rev:
        c = c * math.pi / 180.0
        d = _get_double(e, "d", 0.0)
        return JointDesc(dof, ratio, c, True, 0.0 + c, d, a, alpha, mn * math.pi / 180.0, mx * math.pi / 180.0, _get_double(e, "springStiffness", 0.0))
    theta = _get_double(e, "theta", 0.0) * math.pi / 180.0
    return JointDesc(dof, ratio, c, False, theta, 0.0 + c, a, alpha, mn, mx, _get_double(e, "springStiffness", 0.0))


def apply_static_ratios(chains, dof_type):
    """JointDesc.ratio is what the reference calls DOF::getStaticRatio(joint): joint value = ratio x DOF value.  For Rigid
    and BreakAway DOFs that is the coupling ratio of the DOF expression (dof.cpp:484-487).  CompliantDOF::getStaticRatio
    (dof.cpp:779-795) keeps the coupling ratio for joints number 0, 2, 4, 6 of the robot and gives every other joint
    (ref / k) * tau1 * tau2 / (tau1 + tau2): tau1, ref = coupling ratio and spring stiffness of the DOF's first joint,
    tau2, k = its own (k == 0 -> ref).  Joint numbers are chain-major (Joint::getNum)."""
    joints = [j for c in chains for j in c.joints]
    first = {}
    for jn, j in enumerate(joints):
        first.setdefault(j.dof, jn)
    coupling = [j.ratio for j in joints]
    for jn, j in enumerate(joints):
        if dof_type[j.dof] != "c" or jn in (0, 2, 4, 6):
            continue
        f = first[j.dof]
        ref = joints[f].stiffness if joints[f].stiffness != 0.0 else 1.0        # the reference asserts ref != 0
        k = j.stiffness if j.stiffness != 0.0 else ref
        j.coupling = coupling[jn]
        j.ratio = (ref / k) * (coupling[f] * coupling[jn]) / (coupling[f] + coupling[jn])


_DYN_STEP = {"Revolute": 1, "Ball": 3, "Universal": 2, "Prismatic": 1, "Fixed": 0}


def load_robot(scene: Scene, xml_path: str, tran: Transf, dof_values=None) -> int:
    """Robot XML -> RobotDesc + its bodies appended to ``scene`` in World::addRobot order
    (base first, then links chain-major: src/world.cpp:1060-1068)."""
    root = ET.parse(xml_path).getroot()
    rdir = os.path.dirname(xml_path)
    ivdir = os.path.join(rdir, "iv")
    ridx = len(scene.robots)
    rname = (root.get("type") or os.path.splitext(os.path.basename(xml_path))[0]).strip()

    base = load_body(os.path.join(ivdir, _text(root.find("palm"))), name=f"{rname}_base", kind="base")
    base.robot = ridx
    base_idx = scene.add_body(base, tran)

    dofs = root.findall("dof")
    n_dof = len(dofs)
    dof_type = [(d.get("type") or "r").strip() for d in dofs]
    dof_default = np.array([(_get_double(d, "defaultValue", 0.0) or 0.0) * math.pi / 180.0 for d in dofs])
    dof_velocity = np.array([(_get_double(d, "defaultVelocity", 0.0) or 0.0) for d in dofs])

    chains: List[ChainDesc] = []
    for ci, ce in enumerate(root.findall("chain")):
        ctran = parse_transform(ce.find("transform"))
        joints = [_parse_joint(j) for j in ce.findall("joint")]
        link_bodies, last_joint = [], []
        last = -1
        chain_first = len(scene.bodies)
        for li, le in enumerate(ce.findall("link")):
            jt = (le.get("dynamicJointType") or "").strip()
            last += _DYN_STEP[jt]
            lb = load_body(os.path.join(ivdir, _text(le)), name=f"{rname}_chain{ci}_link{li}", kind="link")
            lb.robot, lb.chain, lb.link = ridx, ci, li
            bidx = scene.add_body(lb, tf.IDENTITY)
            link_bodies.append(bidx)
            last_joint.append(last)
            rule = (le.get("collisionRule") or "").strip()
            if rule == "Off":
                scene.disable(bidx, bidx)       # whole-body off is recorded via kind below
                lb.kind = "link_nocollide"
            elif rule == "OverlappingPair":
                tchain = (le.get("targetChain") or "").strip()
                tlink = le.get("targetLink")
                if tchain == "base":
                    scene.disable(bidx, base_idx)
                elif tlink is not None:
                    # only links already created can be referenced
                    tci = ci if tchain == "" else int(tchain)
                    tgt = (chain_first + int(tlink)) if tci == ci else chains[tci].link_bodies[int(tlink)]
                    scene.disable(bidx, tgt)
        chains.append(ChainDesc(ctran, joints, link_bodies, last_joint))

    if "c" in dof_type:
        apply_static_ratios(chains, dof_type)
    # DOF limits: intersection of joint ranges / ratio (DOF::updateMinMax, src/dof.cpp:93-121)
    dof_min = np.full(n_dof, -np.inf)
    dof_max = np.full(n_dof, np.inf)
    for d in range(n_dof):
        first = True
        for c in chains:
            for j in c.joints:
                if j.dof != d:
                    continue
                mx, mn = j.max / j.ratio, j.min / j.ratio
                if mx < mn:
                    mx, mn = mn, mx
                if first:
                    dof_max[d], dof_min[d] = mx, mn
                    first = False
                else:
                    dof_max[d] = min(dof_max[d], mx)
                    dof_min[d] = max(dof_min[d], mn)
        dof_default[d] = min(max(dof_default[d], dof_min[d]), dof_max[d])

    # structural pair disabling (src/world.cpp:1069-1078)
    for c in chains:
        scene.disable(c.link_bodies[0], base_idx)
        for a in c.link_bodies:
            for b in c.link_bodies:
                scene.disable(a, b)
    # <disableCollision> (src/robot.cpp:246-280)
    for e in root.findall("disableCollision"):
        try:
            b1c, b1l = int(e.get("body1_chain")), int(e.get("body1_link"))
            b2c, b2l = int(e.get("body2_chain")), int(e.get("body2_link"))
        except (TypeError, ValueError):
            continue
        try:
            x = base_idx if b1c < 0 else chains[b1c].link_bodies[b1l]
            y = base_idx if b2c < 0 else chains[b2c].link_bodies[b2l]
        except IndexError:
            continue
        scene.disable(x, y)

    # approach direction (src/robot.cpp:283-340)
    approach = tf.IDENTITY
    ae = root.find("approachDirection")
    if ae is not None:
        loc = np.array([float(np.float32(v)) for v in _text(ae.find("referenceLocation")).split()])
        z = np.array([float(np.float32(v)) for v in _text(ae.find("direction")).split()])
        z = z / np.linalg.norm(z)
        x = np.array([0.0, 1.0, 0.0]) if abs(z[0]) > 0.9 else np.array([1.0, 0.0, 0.0])
        y = np.cross(z, x)
        x = np.cross(y, z)
        approach = tf.from_matrix(np.stack([x, y, z], axis=1), loc)

    robot = RobotDesc(name=rname, base_body=base_idx, chains=chains, n_dof=n_dof, dof_type=dof_type,
                      dof_min=dof_min, dof_max=dof_max, dof_default=dof_default,
                      dof_values=dof_default.copy(), tran=tran, approach=approach, dof_velocity=dof_velocity)

    eg = root.find("eigenGrasps")
    if eg is not None and _text(eg):
        _load_eigengrasps(robot, os.path.join(rdir, _text(eg)))
    vc = root.find("virtualContacts")
    if vc is not None and _text(vc):
        _load_virtual_contacts(robot, os.path.join(rdir, _text(vc)))
    if dof_values is not None:
        dv = np.asarray(dof_values, dtype=np.float64)[:n_dof]
        robot.dof_values = np.clip(dv, dof_min, dof_max)
    scene.robots.append(robot)
    return ridx


def _load_eigengrasps(robot: RobotDesc, path: str):
    """src/eigenGrasp.cpp:372-478 (+ normalize :92-104, checkOrigin :480-494)."""
    root = ET.parse(path).getroot()
    n = int(float(root.get("dimensions")))

    def dimvals(e):
        dv = e.find("DimVals")
        return np.array([float(dv.get(f"d{i}")) if dv is not None and dv.get(f"d{i}") not in (None, "") else 0.0
                         for i in range(n)])

    egs, limits = [], []
    for e in root.findall("EG"):
        v = dimvals(e)
        v = v / math.sqrt(float(v @ v))
        egs.append(v)
        lim = e.find("Limits")
        limits.append([float(lim.get("min")), float(lim.get("max"))] if lim is not None else [np.nan, np.nan])
    org = root.findall("ORIGIN")
    if org:
        origin = np.clip(dimvals(org[0]), robot.dof_min[:n], robot.dof_max[:n])
    else:
        # setSimpleOrigin: halfway between the limits
        origin = 0.5 * (robot.dof_min[:n] + robot.dof_max[:n])
    nrm = root.findall("NORM")
    norm = dimvals(nrm[0]) if nrm else np.ones(n)
    robot.eg = np.array(egs)
    robot.eg_origin = origin
    robot.eg_norm = norm
    robot.eg_limits = np.array(limits)


def _load_virtual_contacts(robot: RobotDesc, path: str):
    """src/robot.cpp:435-475, src/contact/virtualContact.cpp:399-494."""
    root = ET.parse(path).getroot()
    for e in root:
        if e.tag != "virtual_contact":
            continue
        f = int(float(_text(e.find("finger_number"))))
        l = int(float(_text(e.find("link_number"))))
        loc = np.array([float(v) for v in _text(e.find("location")).split()])
        q = np.array([float(v) for v in _text(e.find("rotation")).split()])
        t = np.array([float(v) for v in _text(e.find("translation")).split()])
        nrm = np.array([float(v) for v in _text(e.find("normal")).split()])
        body = robot.base_body if f < 0 else robot.chains[f].link_bodies[l]
        robot.vcontacts.append(VirtualContactDesc(body, f, l, loc, nrm, tf.make(q, t)))


def load_world(xml_path: str, graspit_root: str) -> Scene:
    """World XML (src/world.cpp:549-797): elements in file order."""
    root = ET.parse(xml_path).getroot()
    scene = Scene(name=os.path.splitext(os.path.basename(xml_path))[0])
    for child in list(root):
        tag = child.tag.strip()
        if tag in ("obstacle", "graspableBody"):
            fn = os.path.join(graspit_root, _text(child.find("filename")))
            b = load_body(fn, kind="obstacle" if tag == "obstacle" else "object")
            scene.add_body(b, parse_transform(child.find("transform")))
        elif tag == "robot":
            fn = os.path.join(graspit_root, _text(child.find("filename")))
            dv = child.find("dofValues")
            vals = None
            if dv is not None and _text(dv):
                # per-DOF state tokens; rigid DOFs store one number, breakaway DOFs store more
                vals = _parse_dof_values(_text(dv))
            ridx = load_robot(scene, fn, parse_transform(child.find("transform")), None)
            if vals is not None:
                _assign_dof_values(scene.robots[ridx], vals)
        elif tag == "connection":
            p = int(_text(child.find("parentRobot")))
            pc = int(_text(child.find("parentChain")))
            c = int(_text(child.find("childRobot")))
            off = parse_transform(child.find("transform"))
            parent, kid = scene.robots[p], scene.robots[c]
            last_link = parent.chains[pc].link_bodies[-1]
            mf = child.find("mountFilename")
            if mf is not None and _text(mf):
                mb = load_body(os.path.join(graspit_root, _text(mf)), name=f"{kid.name}_mount", kind="mount")
                mb.robot = c
                midx = scene.add_body(mb, tf.IDENTITY)
                kid.mount_body = midx
                scene.disable(last_link, midx)
                scene.disable(midx, kid.base_body)
            # Robot::attachRobot (src/robot.cpp:851-865)
            kid.parent, kid.parent_chain = p, pc
            parent.chains[pc].children.append((c, off))
            scene.disable(last_link, kid.base_body)
            if kid.mount_body >= 0:
                scene.disable(last_link, kid.mount_body)
    return scene


def _parse_dof_values(text: str) -> List[float]:
    return [float(v) for v in text.replace("+", " +").split()]


def _assign_dof_values(robot: RobotDesc, vals: List[float]):
    """World files store DOF state via DOF::writeToStream: a rigid DOF writes ``q``; a breakaway
    DOF writes ``q n_joints`` + per-joint flags (src/dof.cpp:304-320,700-731).  The batched
    engine only needs q per DOF."""
    out = robot.dof_default.copy()
    k = 0
    for d in range(robot.n_dof):
        if k >= len(vals):
            break
        out[d] = vals[k]
        k += 1
        if robot.dof_type[d] == "b" and k < len(vals):
            # BreakAwayDOF::writeToStream: q, then for each joint "inBreakAway [value]"
            nj = sum(1 for c in robot.chains for j in c.joints if j.dof == d)
            for _ in range(nj):
                if k >= len(vals):
                    break
                flag = vals[k]
                k += 1
                if flag != 0 and k < len(vals):
                    k += 1
    robot.dof_values = np.clip(out, robot.dof_min, robot.dof_max)
