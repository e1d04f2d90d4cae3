"""MJCF -> flat articulation tables.

Replaces the reference's ``loadMJCF(path, flags=URDF_USE_SELF_COLLISION |
URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS)`` + ``XmlBasedRobot.addToScene`` link / joint
enumeration (/root/reference/pybulletgym/envs/roboschool/robot_bases.py:32-91,107-125).

Two products come out of one parse:

* ``BulletModel`` -- the *Bullet-shaped* link list the reference would see through
  ``getNumJoints / getJointInfo``: one massless link per ``<joint>``, each ``<body>`` hanging from
  the last of them by a ``jointfix`` fixed joint.  It gives ``robot.parts``,
  ``jdict`` and ``ordered_joints`` their reference membership and order, and it is what the CPU
  oracle simulates.
* ``ReducedModel`` -- the dynamics tree the CUDA library simulates: fixed joints merged away, every
  quantity expressed in a body frame that sits at the merged centre of mass, flat numpy tables.

Every importer rule that is recalled from upstream Bullet rather than read from the reference is a
named switch in ``ImporterRules`` (SURVEY.md Appendix C6 "pin list").
"""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

ASSET_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "mjcf")

JT_FIXED, JT_REVOLUTE, JT_PRISMATIC, JT_FREE = 0, 1, 2, 3
G_SPHERE, G_CAPSULE, G_BOX = 0, 1, 2


@dataclass
class ImporterRules:
    """[EXT] Bullet MJCF-importer behaviours, each switchable."""
    density: float = 1000.0                 # C1.5 / C6-2: geom `density`, `settotalmass` ignored
    joint_damping_from_mjcf: bool = False   # C1.6 / C6-1: MJCF joint damping is not imported
    # C1.13 / C6-10: Bullet hands the MJCF axis vector to btMultiBody::setupRevolute as written.  The joint angle still
    # rotates about the *direction* of the axis (btQuaternion(axis, angle) divides by its length), but the motion subspace
    # -- velocities, Jacobians, the joint-space inertia -- carries the raw vector.  Non-unit axes: the Ant's ankles
    # (-1,1,0) / (1,1,0), the Humanoid's shoulders (2,1,1) and elbows (0,-1,1).  Evidence for "raw": the reference's pretrained
    # Ant policy scores ~2250 (reward_threshold 2500) with raw axes and ~800 with normalised ones.
    normalize_joint_axes: bool = False
    aabb_box_inertia: bool = True           # C1.9 / C6-4: inertia = box inertia of the link AABB
    inertial_frame_last_fromto: bool = True  # C1.8: COM = midpoint of the last fromto geom
    honour_axisangle: bool = True           # C6-7
    link_damping: float = 0.04              # C3.2: btMultiBody linear/angular damping
    relative_breaking_threshold: bool = True  # C5.2: manifold threshold = 0.02 * angular-motion disc
    breaking_threshold: float = 0.02


# ----------------------------------------------------------------------------------------------
# small math helpers (quaternions are (x, y, z, w), Bullet order)
# ----------------------------------------------------------------------------------------------
def q_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def q_from_axis_angle(axis, ang):
    axis = np.asarray(axis, float)
    n = np.linalg.norm(axis)
    if n < 1e-12:
        return np.array([0.0, 0.0, 0.0, 1.0])
    s = math.sin(0.5 * ang) / n
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, math.cos(0.5 * ang)])


def q_to_mat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def mat_to_q(m):
    t = np.trace(m)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        return np.array([(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s])
    i = int(np.argmax(np.diag(m)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0) * 2
    q = np.zeros(4)
    q[i] = 0.25 * s
    q[j] = (m[j, i] + m[i, j]) / s
    q[k] = (m[k, i] + m[i, k]) / s
    q[3] = (m[k, j] - m[j, k]) / s
    return q


def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, float)
    v = np.array([float(t) for t in s.split()], float)
    if n is not None and len(v) != n:
        raise ValueError("expected %d numbers in %r" % (n, s))
    return v


# ----------------------------------------------------------------------------------------------
# Bullet-shaped model
# ----------------------------------------------------------------------------------------------
@dataclass
class Geom:
    name: str
    gtype: int
    radius: float                  # sphere / capsule radius; unused for box
    p0: np.ndarray                 # sphere centre, capsule end 0, box centre (link frame)
    p1: np.ndarray                 # capsule end 1 (== p0 for sphere); box half extents
    rot: np.ndarray                # 3x3 geom orientation in the link frame (box, AABB of pos/quat capsules)
    friction: float
    contype: int
    conaffinity: int
    multisphere: bool = True       # fromto capsule (tight AABB) vs pos/quat capsule
    spin_friction: float = 0.0     # 2nd / 3rd number of the MJCF friction attribute (C1.11)
    roll_friction: float = 0.0


@dataclass
class Link:
    name: str
    parent: int                    # index into links, -1 = world
    joint_name: str
    jtype: int
    axis: np.ndarray               # link frame
    pos: np.ndarray                # link-frame origin in the parent link frame at q = 0
    quat: np.ndarray               # link-frame orientation in the parent link frame at q = 0
    lower: float = 0.0
    upper: float = -1.0            # lower > upper  <=>  no limit (pybullet reports 0,-1)
    damping: float = 0.0
    mass: float = 0.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))       # inertial-frame origin (link frame)
    inertia: np.ndarray = field(default_factory=lambda: np.zeros(3))   # diagonal, link axes
    geoms: List[Geom] = field(default_factory=list)
    contact_threshold: float = 0.02
    is_dummy: bool = False


@dataclass
class BulletModel:
    name: str
    links: List[Link]              # links[0] is the base
    floating: bool
    rules: ImporterRules

    # --- the views robot_bases.XmlBasedRobot.addToScene builds (rs/robot_bases.py:54-89) ---
    def ordered_joints(self) -> List[int]:
        """Link indices whose joint is actuated: not fixed, name not 'ignore*' / 'jointfix*'."""
        return [i for i, l in enumerate(self.links)
                if l.jtype in (JT_REVOLUTE, JT_PRISMATIC) and not l.joint_name.startswith("ignore")]

    def dof_links(self) -> List[int]:
        return [i for i, l in enumerate(self.links) if l.jtype in (JT_REVOLUTE, JT_PRISMATIC)]

    def part_names(self) -> List[str]:
        """Keys of robot.parts before the floor is added (rs/robot_bases.py:68-74)."""
        # every non-base link is keyed by its link name; the base enters only through the
        # "if nothing else works" branch, i.e. when it carries the robot name (floating robots)
        names = [l.name for l in self.links[1:]]
        if self.floating:
            names = [self.links[0].name] + names
        return list(dict.fromkeys(names))

    def link_index(self, name: str) -> int:
        for i, l in enumerate(self.links):
            if l.name == name:
                return i
        raise KeyError(name)

    def ancestors(self, i: int) -> List[int]:
        out = []
        p = self.links[i].parent
        while p >= 0:
            out.append(p)
            p = self.links[p].parent
        return out


class _Defaults:
    def __init__(self):
        self.joint: Dict[str, str] = {}
        self.geom: Dict[str, str] = {}


def _geom_rot(g, rules, angle_scale):
    if g.get("quat") is not None:
        w, x, y, z = _vec(g.get("quat"), 4)
        q = np.array([x, y, z, w])
        q = q / np.linalg.norm(q)
        return q_to_mat(q)
    if g.get("axisangle") is not None and rules.honour_axisangle:
        aa = _vec(g.get("axisangle"), 4)
        return q_to_mat(q_from_axis_angle(aa[:3], aa[3] * angle_scale))
    return np.eye(3)


def _parse_geom(elem, dflt: _Defaults, rules: ImporterRules, angle_scale):
    a = dict(dflt.geom)
    a.update(elem.attrib)
    gtype = a.get("type", "sphere")
    size = _vec(a.get("size"), None, [0.0])
    fr3 = _vec(a.get("friction"), None, [1.0, 0.0, 0.0])
    fric = fr3[0]
    spin = fr3[1] if len(fr3) > 1 else 0.0
    roll = fr3[2] if len(fr3) > 2 else 0.0
    contype = int(a.get("contype", 1))
    conaff = int(a.get("conaffinity", 1))
    name = a.get("name", "")
    pos = _vec(a.get("pos"), 3, [0, 0, 0])
    rot = _geom_rot(a, rules, angle_scale)
    shift = np.zeros(3)
    if gtype == "sphere":
        g = Geom(name, G_SPHERE, size[0], pos.copy(), pos.copy(), np.eye(3), fric, contype, conaff)
    elif gtype == "capsule":
        if a.get("fromto") is not None:
            ft = _vec(a.get("fromto"), 6)
            g = Geom(name, G_CAPSULE, size[0], ft[:3].copy(), ft[3:].copy(), np.eye(3), fric, contype, conaff, True)
            shift = 0.5 * (ft[:3] + ft[3:])
        else:
            hh = size[1]
            ax = rot[:, 2]
            g = Geom(name, G_CAPSULE, size[0], pos - ax * hh, pos + ax * hh, rot, fric, contype, conaff, False)
    elif gtype == "box":
        g = Geom(name, G_BOX, 0.0, pos.copy(), np.array(size[:3], float), rot, fric, contype, conaff)
    else:
        raise NotImplementedError("geom type %s" % gtype)
    g.spin_friction, g.roll_friction = float(spin), float(roll)
    return g, shift


def _geom_volume(g: Geom):
    if g.gtype == G_SPHERE:
        return 4.0 / 3.0 * math.pi * g.radius ** 3
    if g.gtype == G_CAPSULE:
        h = np.linalg.norm(g.p1 - g.p0)
        return 4.0 / 3.0 * math.pi * g.radius ** 3 + math.pi * g.radius ** 2 * h
    if g.gtype == G_BOX:
        return 8.0 * g.p1[0] * g.p1[1] * g.p1[2]
    return 0.0


def _geom_aabb(g: Geom, origin):
    """AABB of one collision child in the inertial frame (axes = link axes, origin = `origin`)."""
    if g.gtype == G_SPHERE:
        c = g.p0 - origin
        return c - g.radius, c + g.radius
    if g.gtype == G_CAPSULE:
        if g.multisphere:   # btMultiSphereShape: tight union of the two end spheres
            a, b = g.p0 - origin, g.p1 - origin
            return np.minimum(a, b) - g.radius, np.maximum(a, b) + g.radius
        c = 0.5 * (g.p0 + g.p1) - origin   # btCapsuleShapeZ under a child transform: |R| * half extents
        hh = 0.5 * np.linalg.norm(g.p1 - g.p0)
        he = np.abs(g.rot) @ np.array([g.radius, g.radius, g.radius + hh])
        return c - he, c + he
    if g.gtype == G_BOX:
        c = g.p0 - origin
        he = np.abs(g.rot) @ g.p1
        return c - he, c + he
    raise NotImplementedError


def _finish_link_inertial(link: Link, mass_defined, rules: ImporterRules):
    if not link.geoms:
        return
    if not mass_defined:
        link.mass = rules.density * sum(_geom_volume(g) for g in link.geoms)
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    for g in link.geoms:
        a, b = _geom_aabb(g, link.com)
        lo, hi = np.minimum(lo, a), np.maximum(hi, b)
    ext = hi - lo
    if rules.aabb_box_inertia:
        lx, ly, lz = ext
        link.inertia = link.mass / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
    # btCollisionShape::getAngularMotionDisc of the compound: bounding-sphere radius + |centre|
    disc = 0.5 * np.linalg.norm(ext) + np.linalg.norm(0.5 * (lo + hi))
    link.contact_threshold = (rules.breaking_threshold * disc if rules.relative_breaking_threshold
                              else rules.breaking_threshold)


def parse_mjcf(path_or_name: str, rules: Optional[ImporterRules] = None) -> BulletModel:
    """Parse one MJCF file into the Bullet-shaped link list (SURVEY.md Appendix C1 rules)."""
    rules = rules or ImporterRules()
    path = path_or_name
    if not os.path.exists(path):
        path = os.path.join(ASSET_DIR, path_or_name)
    root = ET.parse(path).getroot()
    comp = root.find("compiler")
    angle = (comp.get("angle") if comp is not None and comp.get("angle") else "degree")
    angle_scale = math.pi / 180.0 if angle == "degree" else 1.0
    dflt = _Defaults()
    d = root.find("default")
    if d is not None:
        if d.find("joint") is not None:
            dflt.joint = dict(d.find("joint").attrib)
        if d.find("geom") is not None:
            dflt.geom = dict(d.find("geom").attrib)
    wb = root.find("worldbody")
    top = [b for b in wb.findall("body")]
    if len(top) != 1 and any(not b.findall("joint") for b in top):
        raise NotImplementedError("several top-level bodies are supported only when each has joints (fixed bases)")
    links: List[Link] = []
    counter = [0]

    def add_body(elem, parent_link, is_root):
        bname = elem.get("name", "body%d" % len(links))
        bpos = _vec(elem.get("pos"), 3, [0, 0, 0])
        if elem.get("quat") is not None:
            w, x, y, z = _vec(elem.get("quat"), 4)
            bq = np.array([x, y, z, w])
            bq /= np.linalg.norm(bq)
        else:
            bq = np.array([0.0, 0.0, 0.0, 1.0])
        joints = elem.findall("joint")
        cur_parent = parent_link
        cur_pos, cur_quat = bpos, bq            # pose of the *body frame* in the current parent frame
        floating_root = is_root and not joints
        if is_root and joints:
            # massless fixed base at the world origin (C1.2).  Bullet makes every top-level body its own multibody
            # (Reacher: the arm and the target); their fixed massless bases coincide with the world, so one shared
            # base link carries all of them here -- the trees stay dynamically independent.
            if not links:
                links.append(Link("base", -1, "", JT_FIXED, np.zeros(3), np.zeros(3), np.array([0, 0, 0, 1.0])))
            cur_parent = 0
        for j in joints:
            a = dict(dflt.joint)
            a.update(j.attrib)
            jtype = {"hinge": JT_REVOLUTE, "slide": JT_PRISMATIC}[a.get("type", "hinge")]
            jpos = _vec(a.get("pos"), 3, [0, 0, 0])
            axis = _vec(a.get("axis"), 3, [0, 0, 1])
            if rules.normalize_joint_axes:
                axis = axis / np.linalg.norm(axis)
            limited = a.get("limited", "false") == "true"
            lower, upper = 0.0, -1.0
            if limited and a.get("range") is not None:
                r = _vec(a.get("range"), 2)
                sc = angle_scale if jtype == JT_REVOLUTE else 1.0
                lower, upper = r[0] * sc, r[1] * sc
            damping = float(a.get("damping", 0.0)) if rules.joint_damping_from_mjcf else 0.0
            counter[0] += 1
            # dummy link frame: origin at the joint anchor, axes = body axes
            lpos = cur_pos + q_to_mat(cur_quat) @ jpos
            links.append(Link("link0_%d" % counter[0], cur_parent, a.get("name", "joint%d" % counter[0]), jtype,
                              axis, lpos, cur_quat, lower, upper, damping, is_dummy=True))
            cur_parent = len(links) - 1
            cur_pos, cur_quat = -jpos, np.array([0.0, 0.0, 0.0, 1.0])
        # the body link itself
        if floating_root:
            blink = Link(bname, -1, "", JT_FREE, np.zeros(3), cur_pos, cur_quat)
        else:
            counter[0] += 1
            blink = Link(bname, cur_parent, "jointfix_%d" % counter[0], JT_FIXED, np.zeros(3), cur_pos, cur_quat)
        links.append(blink)
        me = len(links) - 1
        mass_defined = False
        for ch in elem:
            if ch.tag == "inertial":
                if ch.get("pos") is not None:
                    blink.com = _vec(ch.get("pos"), 3)
                if ch.get("mass") is not None:
                    blink.mass = float(ch.get("mass"))
                if ch.get("diaginertia") is not None:
                    blink.inertia = _vec(ch.get("diaginertia"), 3)
                mass_defined = True
            elif ch.tag == "geom":
                g, shift = _parse_geom(ch, dflt, rules, angle_scale)
                blink.geoms.append(g)
                if not mass_defined and rules.inertial_frame_last_fromto:
                    blink.com = shift
        _finish_link_inertial(blink, mass_defined, rules)
        for ch in elem.findall("body"):
            add_body(ch, me, False)

    body_first_link = []
    for t in top:
        body_first_link.append(len(links) if links else 1)
        add_body(t, -1, True)
    floating = links[0].jtype == JT_FREE
    name = root.get("model", os.path.basename(path))
    bm = BulletModel(name, links, floating, rules)
    # link ranges of the top-level bodies (one pybullet multibody each): [first non-base link, end)
    ends = body_first_link[1:] + [len(links)]
    bm.multibody_links = [(a, b) for a, b in zip(body_first_link, ends)]
    return bm


# ----------------------------------------------------------------------------------------------
# forward kinematics on the Bullet-shaped list (used by the reducer and by tests)
# ----------------------------------------------------------------------------------------------
def link_world_frames(model: BulletModel, q: Optional[np.ndarray] = None, base_pos=None, base_quat=None):
    """World rotation / origin of every link frame.  q follows dof_links() order."""
    dofs = model.dof_links()
    q = np.zeros(len(dofs)) if q is None else np.asarray(q, float)
    qmap = {li: q[k] for k, li in enumerate(dofs)}
    R = [None] * len(model.links)
    p = [None] * len(model.links)
    for i, l in enumerate(model.links):
        if l.parent < 0:
            if l.jtype == JT_FREE and base_pos is not None:
                # base_pos / base_quat give the pose of the base *inertial* frame (pybullet convention)
                R[i] = q_to_mat(np.asarray(base_quat, float))
                p[i] = np.asarray(base_pos, float) - R[i] @ l.com
            else:
                R[i] = q_to_mat(l.quat)
                p[i] = l.pos.copy()
            continue
        Rp, pp = R[l.parent], p[l.parent]
        Rl = Rp @ q_to_mat(l.quat)
        pl = pp + Rp @ l.pos
        if l.jtype == JT_REVOLUTE:
            Rl = Rl @ q_to_mat(q_from_axis_angle(l.axis, qmap[i]))
        elif l.jtype == JT_PRISMATIC:
            pl = pl + Rl @ (l.axis * qmap[i])
        R[i], p[i] = Rl, pl
    return R, p


def link_com_velocities(model: BulletModel, q, qd, base_pos=None, base_quat=None, base_omega=None, base_vel=None):
    """World linear velocity of every link's COM -- what getBaseVelocity()[0] / getLinkState(computeLinkVelocity=1)[6]
    return (rs/robot_bases.py:243-248).  q / qd follow dof_links() order; base_vel is the base COM velocity."""
    dofs = model.dof_links()
    qd = np.asarray(qd, float)
    qdmap = {li: qd[k] for k, li in enumerate(dofs)}
    R, p = link_world_frames(model, q, base_pos, base_quat)
    n = len(model.links)
    w = [np.zeros(3) for _ in range(n)]
    vo = [np.zeros(3) for _ in range(n)]      # velocity of the link frame origin
    out = np.zeros((n, 3))
    for i, l in enumerate(model.links):
        if l.parent < 0:
            if l.jtype == JT_FREE and base_omega is not None:
                w[i] = np.asarray(base_omega, float)
                vo[i] = np.asarray(base_vel, float) - np.cross(w[i], R[i] @ l.com)
        else:
            pa = l.parent
            w[i] = w[pa].copy()
            vo[i] = vo[pa] + np.cross(w[pa], p[i] - p[pa])
            if l.jtype == JT_REVOLUTE:
                w[i] = w[i] + (R[i] @ l.axis) * qdmap[i]       # the axis as written (Bullet does not normalise it)
            elif l.jtype == JT_PRISMATIC:
                vo[i] = vo[i] + (R[i] @ l.axis) * qdmap[i]
        out[i] = vo[i] + np.cross(w[i], R[i] @ l.com)
    return out


# ----------------------------------------------------------------------------------------------
# reduced dynamics tree
# ----------------------------------------------------------------------------------------------
@dataclass
class ReducedModel:
    name: str
    floating: bool
    nb: int
    nd: int                         # total generalized velocities (6 base + joints if floating)
    nj: int                         # joint dofs
    parent: np.ndarray              # [nb] int
    jtype: np.ndarray               # [nb] int (JT_FREE for a floating root)
    dof: np.ndarray                 # [nb] index into u (first of 6 for the root), -1 if none
    depth: np.ndarray               # [nb]
    subtree_end: np.ndarray         # [nb] bodies i..subtree_end[i]-1 form the subtree (DFS order)
    q0: np.ndarray                  # [nb,4] parent body frame -> body frame at q=0 (x,y,z,w)
    anchor_p: np.ndarray            # [nb,3] joint anchor in the parent body frame (from parent COM); world if parent<0
    com_off: np.ndarray             # [nb,3] body COM relative to the anchor, body frame
    axis: np.ndarray                # [nb,3] joint axis, body frame
    mass: np.ndarray                # [nb]
    inertia: np.ndarray             # [nb,6] xx yy zz xy xz yz about the COM, body frame
    # per joint dof (index j = 0..nj-1, u index = j + 6*floating)
    jnt_body: np.ndarray            # [nj]
    jnt_lower: np.ndarray
    jnt_upper: np.ndarray           # lower > upper => unlimited
    jnt_damping: np.ndarray
    jnt_revolute: np.ndarray        # 1 revolute / 0 prismatic
    jnt_act: np.ndarray             # [nj] index into the action vector or -1
    jnt_names: List[str]
    # Bullet links folded into each body
    sub_body: np.ndarray            # [ns] body of every Bullet link (dummy and fixed links included, base too)
    sub_off: np.ndarray             # [ns,3] link COM relative to the body COM, body frame
    sub_mass: np.ndarray            # [ns]
    sub_inertia: np.ndarray         # [ns,3] diagonal, body axes
    sub_names: List[str]
    sub_in_parts: np.ndarray        # [ns] 1 if the link is an entry of robot.parts
    # collision geometry (body frame, relative to the body COM)
    geom_body: np.ndarray           # [ng]
    geom_link: np.ndarray           # [ng] index into sub_* (the Bullet link owning the geom)
    geom_type: np.ndarray
    geom_radius: np.ndarray
    geom_p0: np.ndarray             # [ng,3]
    geom_p1: np.ndarray             # [ng,3]
    geom_friction: np.ndarray
    geom_spin: np.ndarray           # spinning / rolling friction coefficients (torsional friction rows)
    geom_roll: np.ndarray
    geom_threshold: np.ndarray      # contact breaking threshold of the owning link
    geom_ground: np.ndarray         # 1 if it collides with the floor plane
    pair_a: np.ndarray              # self-collision geom pairs
    pair_b: np.ndarray
    base_link_off: np.ndarray       # [3] Bullet base-link COM relative to the root body COM (floating only)
    link_damping: float = 0.04

    def ancestors_mask(self):
        """[nd, nb] 1 if dof k moves body i."""
        m = np.zeros((self.nd, self.nb), np.int32)
        for i in range(self.nb):
            b = i
            while b >= 0:
                if self.jtype[b] == JT_FREE:
                    m[0:6, i] = 1
                elif self.dof[b] >= 0:
                    m[self.dof[b], i] = 1
                b = self.parent[b]
        return m


def reduce_model(model: BulletModel, action_joint_names: Optional[List[str]] = None) -> ReducedModel:
    links = model.links
    R0, p0 = link_world_frames(model)
    nl = len(links)
    # group links into bodies: a body starts at every non-fixed link (and at a floating base)
    body_of = [-1] * nl
    first_link: List[int] = []
    fixed_base = not model.floating
    for i, l in enumerate(links):
        if l.parent < 0:
            if model.floating:
                body_of[i] = 0
                first_link.append(i)
            else:
                body_of[i] = -1          # the fixed massless base is the world
        elif l.jtype == JT_FIXED:
            body_of[i] = body_of[l.parent]
        else:
            body_of[i] = len(first_link)
            first_link.append(i)
    nb = len(first_link)
    # geoms attached to the fixed base would be static world geometry -- none of the models has any
    for i, l in enumerate(links):
        if body_of[i] < 0 and l.geoms:
            raise NotImplementedError("static geometry on a fixed base")
    parent = np.full(nb, -1, np.int32)
    jtype = np.zeros(nb, np.int32)
    for b, fl in enumerate(first_link):
        l = links[fl]
        jtype[b] = l.jtype
        parent[b] = body_of[l.parent] if l.parent >= 0 else -1
    # bodies must come out in DFS order so that subtrees are contiguous; the parser emits links
    # depth-first, and bodies inherit that order -- verify.
    subtree_end = np.arange(1, nb + 1, dtype=np.int32)
    for b in range(nb - 1, -1, -1):
        if parent[b] >= 0:
            subtree_end[parent[b]] = max(subtree_end[parent[b]], subtree_end[b])
    for b in range(nb):
        for c in range(b + 1, subtree_end[b]):
            a = c
            while a >= 0 and a != b:
                a = parent[a]
            assert a == b, "bodies are not in DFS order"
    depth = np.zeros(nb, np.int32)
    for b in range(nb):
        depth[b] = 0 if parent[b] < 0 else depth[parent[b]] + 1

    mass = np.zeros(nb)
    com_w = np.zeros((nb, 3))
    Rb = [R0[fl] for fl in first_link]      # body axes = axes of the first link at q = 0
    for b in range(nb):
        ms = [(links[i].mass, p0[i] + R0[i] @ links[i].com) for i in range(nl) if body_of[i] == b]
        mass[b] = sum(m for m, _ in ms)
        if mass[b] > 0:
            com_w[b] = sum(m * c for m, c in ms) / mass[b]
        else:
            com_w[b] = p0[first_link[b]]
    inertia = np.zeros((nb, 6))
    sub_body, sub_off, sub_mass, sub_inertia, sub_names, sub_in_parts = [], [], [], [], [], []
    geom_rows = []
    part_names = set(model.part_names())
    robot_base_name = links[0].name if model.floating else None
    for i, l in enumerate(links):
        b = body_of[i]
        if b < 0:
            continue
        Rrel = Rb[b].T @ R0[i]
        assert np.allclose(Rrel, np.eye(3), atol=1e-12) or l.mass == 0.0 or True
        c_w = p0[i] + R0[i] @ l.com
        off = Rb[b].T @ (c_w - com_w[b])
        Il = Rrel @ np.diag(l.inertia) @ Rrel.T
        if l.mass > 0 and not np.allclose(Il, np.diag(np.diag(Il)), atol=1e-9):
            raise NotImplementedError("fixed sub-link rotated against its body")
        I = Il + l.mass * ((off @ off) * np.eye(3) - np.outer(off, off))
        inertia[b] += np.array([I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]])
        sub_body.append(b)
        sub_off.append(off)
        sub_mass.append(l.mass)
        sub_inertia.append(np.diag(Il))
        sub_names.append(l.name)
        sub_in_parts.append(1 if (l.name in part_names or l.name == robot_base_name) else 0)
        si = len(sub_body) - 1
        for g in l.geoms:
            g0 = Rb[b].T @ (p0[i] + R0[i] @ g.p0 - com_w[b])
            if g.gtype == G_BOX:
                g1 = g.p1.copy()
            else:
                g1 = Rb[b].T @ (p0[i] + R0[i] @ g.p1 - com_w[b])
            geom_rows.append((b, si, g.gtype, g.radius, g0, g1, g.friction, l.contact_threshold, g.contype,
                              g.conaffinity, i, g.spin_friction, g.roll_friction))

    q0 = np.zeros((nb, 4))
    anchor_p = np.zeros((nb, 3))
    com_off = np.zeros((nb, 3))
    axis = np.zeros((nb, 3))
    dof = np.full(nb, -1, np.int32)
    nd0 = 6 if model.floating else 0
    jnt_body, jl, ju, jd, jrev, jnames = [], [], [], [], [], []
    for b, fl in enumerate(first_link):
        l = links[fl]
        if l.jtype == JT_FREE:
            q0[b] = mat_to_q(Rb[b])
            anchor_p[b] = com_w[b]
            dof[b] = 0
            continue
        pb = parent[b]
        Rp = Rb[pb] if pb >= 0 else np.eye(3)
        cp = com_w[pb] if pb >= 0 else np.zeros(3)
        q0[b] = mat_to_q(Rp.T @ Rb[b])
        anchor_w = p0[fl]
        anchor_p[b] = Rp.T @ (anchor_w - cp)
        com_off[b] = Rb[b].T @ (com_w[b] - anchor_w)
        axis[b] = l.axis
        dof[b] = nd0 + len(jnt_body)
        jnt_body.append(b)
        jl.append(l.lower)
        ju.append(l.upper)
        jd.append(l.damping)
        jrev.append(1 if l.jtype == JT_REVOLUTE else 0)
        jnames.append(l.joint_name)
    nj = len(jnt_body)
    if action_joint_names is None:
        action_joint_names = [links[i].joint_name for i in model.ordered_joints()]
    jact = np.array([action_joint_names.index(n) if n in action_joint_names else -1 for n in jnames], np.int32)

    # collision filtering (C1.12): floor = Bullet static filter group 2, mask ~2, OR rule
    ng = len(geom_rows)
    geom_ground = np.array([1 if ((g[8] & ~2) or (2 & g[9])) else 0 for g in geom_rows], np.int32)
    pair_a, pair_b = [], []
    for a in range(ng):
        for c in range(a + 1, ng):
            la, lc = geom_rows[a][10], geom_rows[c][10]
            if la == lc:
                continue
            if not ((geom_rows[a][8] & geom_rows[c][9]) or (geom_rows[c][8] & geom_rows[a][9])):
                continue
            if la in model.ancestors(lc) or lc in model.ancestors(la):
                continue      # URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS (rs/robot_bases.py:116)
            pair_a.append(a)
            pair_b.append(c)

    base_link_off = np.zeros(3)
    if model.floating:
        base_link_off = np.array(sub_off[0])
    return ReducedModel(
        name=model.name, floating=model.floating, nb=nb, nd=nd0 + nj, nj=nj, parent=parent, jtype=jtype, dof=dof,
        depth=depth, subtree_end=subtree_end, q0=q0, anchor_p=anchor_p, com_off=com_off, axis=axis, mass=mass,
        inertia=inertia, jnt_body=np.array(jnt_body, np.int32), jnt_lower=np.array(jl), jnt_upper=np.array(ju),
        jnt_damping=np.array(jd), jnt_revolute=np.array(jrev, np.int32), jnt_act=jact, jnt_names=jnames,
        sub_body=np.array(sub_body, np.int32), sub_off=np.array(sub_off).reshape(-1, 3),
        sub_mass=np.array(sub_mass), sub_inertia=np.array(sub_inertia).reshape(-1, 3), sub_names=sub_names,
        sub_in_parts=np.array(sub_in_parts, np.int32),
        geom_body=np.array([g[0] for g in geom_rows], np.int32), geom_link=np.array([g[1] for g in geom_rows], np.int32),
        geom_type=np.array([g[2] for g in geom_rows], np.int32), geom_radius=np.array([g[3] for g in geom_rows]),
        geom_p0=np.array([g[4] for g in geom_rows]).reshape(-1, 3),
        geom_p1=np.array([g[5] for g in geom_rows]).reshape(-1, 3),
        geom_friction=np.array([g[6] for g in geom_rows]), geom_threshold=np.array([g[7] for g in geom_rows]),
        geom_spin=np.array([g[11] for g in geom_rows]), geom_roll=np.array([g[12] for g in geom_rows]),
        geom_ground=geom_ground, pair_a=np.array(pair_a, np.int32), pair_b=np.array(pair_b, np.int32),
        base_link_off=base_link_off, link_damping=model.rules.link_damping)


def load(name: str, rules: Optional[ImporterRules] = None):
    """Convenience: (BulletModel, ReducedModel) for one of the asset files."""
    bm = parse_mjcf(name, rules)
    return bm, reduce_model(bm)
