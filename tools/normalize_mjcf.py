#!/usr/bin/env python
"""Re-emit the six hot-path MJCF models as canonical, physics-only XML.

The GPU box has no /root/reference, so the model files the loader needs at run time must live in
this repo.  Instead of shipping the upstream files we ship a canonical re-emission that keeps only
what Bullet's MJCF importer reads: the <compiler> angle/coordinate flags,
the <default> joint/geom attributes, and the worldbody tree (body / joint / geom / inertial with
their kinematic, collision-filter and friction attributes).  Rendering attributes, comments,
<actuator>/<tendon>/<option>/<size>/<visual>/<custom> blocks are dropped: the reference never reads
them (robot_locomotors.py hard-codes the motor constants, rs/robot_locomotors.py:86-164).

Usage:  python tools/normalize_mjcf.py /root/reference/pybulletgym/envs/assets/mjcf
Source files: mjcf/{inverted_pendulum,hopper,walker2d,half_cheetah,ant,humanoid_symmetric}.xml
"""
import os
import sys
import xml.etree.ElementTree as ET

MODELS = ["inverted_pendulum", "inverted_double_pendulum", "reacher", "hopper", "walker2d", "half_cheetah", "ant", "humanoid_symmetric"]

KEEP = {
    "mujoco": ["model"],
    "compiler": ["angle", "coordinate"],
    "default": ["class"],
    "joint": ["name", "type", "axis", "pos", "range", "limited", "damping", "armature", "stiffness", "ref"],
    "geom": ["name", "type", "size", "pos", "quat", "axisangle", "fromto", "contype", "conaffinity",
             "friction", "density", "margin"],
    "body": ["name", "pos", "quat"],
    "inertial": ["mass", "pos", "quat", "diaginertia"],
    "worldbody": [],
}
KEEP_CHILDREN = {
    "mujoco": ["compiler", "default", "worldbody"],
    "default": ["joint", "geom", "default"],
    "worldbody": ["geom", "body"],
    "body": ["joint", "geom", "inertial", "body"],
}


def canon(elem):
    out = ET.Element(elem.tag)
    for k in KEEP.get(elem.tag, []):
        if k in elem.attrib:
            out.set(k, " ".join(elem.attrib[k].split()))
    for ch in elem:
        if not isinstance(ch.tag, str):
            continue
        if ch.tag in KEEP_CHILDREN.get(elem.tag, []):
            out.append(canon(ch))
    return out


def indent(e, lvl=0):
    pad = "\n" + " " * lvl
    if len(e):
        e.text = pad + " "
        for c in e:
            indent(c, lvl + 1)
            c.tail = pad + " "
        e[-1].tail = pad
    return e


def main(src_dir):
    dst = os.path.join(os.path.dirname(__file__), "..", "pybullet_gym_b200", "assets", "mjcf")
    os.makedirs(dst, exist_ok=True)
    for m in MODELS:
        root = ET.parse(os.path.join(src_dir, m + ".xml")).getroot()
        out = indent(canon(root))
        txt = ET.tostring(out, encoding="unicode")
        with open(os.path.join(dst, m + ".xml"), "w") as f:
            f.write("<!-- canonical physics-only re-emission; see tools/normalize_mjcf.py -->\n")
            f.write(txt + "\n")
        print("wrote", m)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/pybulletgym/envs/assets/mjcf")
