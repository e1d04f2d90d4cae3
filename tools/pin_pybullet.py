#!/usr/bin/env python
"""Pin the restated physics against a real pybullet -- SURVEY.md Appendix C6 as a tool.

pybullet cannot be installed in the build image, so the oracle's stepSimulation restatement is "parity unpinned"
(oracle/oracle.h).  This script is what closes that gap the day a wheel is available:

  python tools/pin_pybullet.py dump pins.json      # needs `pybullet`, `gym` and the reference (pybulletgym) importable
  python tools/pin_pybullet.py compare pins.json   # needs only this repo; prints a report, exit code 1 on table mismatches

`dump` runs the reference's own env classes and records, per env id: what loadMJCF built (getJointInfo / getDynamicsInfo of
every link: names, types, limits, axes, parent frames, masses, inertia diagonals, inertial frames, friction), the engine
parameters, and a rollout under a fixed action tape with the full state (base pose / velocity, joint states), observation,
reward, done flag and contact links after every step.
`compare` checks our MJCF compiler's tables against the dumped ones (the C1 rules) and replays the tape through the CPU oracle
one step at a time from each dumped state (the C2-C5 rules): the per-step state error is the number the north_star's
"single-step state must match within a stated tolerance" tier asks for.

tests/test_pin_tool.py runs dump on the stub client of tools/fake_pybullet.py (physics = our oracle) and requires compare to
report zero differences -- a self-consistency check of this tool, not a pin.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))

ENVS = {
    "InvertedPendulumPyBulletEnv-v0": ("gym_pendulum_envs", "InvertedPendulumBulletEnv"),
    "InvertedDoublePendulumPyBulletEnv-v0": ("gym_pendulum_envs", "InvertedDoublePendulumBulletEnv"),
    "HopperPyBulletEnv-v0": ("gym_locomotion_envs", "HopperBulletEnv"),
    "Walker2DPyBulletEnv-v0": ("gym_locomotion_envs", "Walker2DBulletEnv"),
    "HalfCheetahPyBulletEnv-v0": ("gym_locomotion_envs", "HalfCheetahBulletEnv"),
    "AntPyBulletEnv-v0": ("gym_locomotion_envs", "AntBulletEnv"),
    "HumanoidPyBulletEnv-v0": ("gym_locomotion_envs", "HumanoidBulletEnv"),
}
STEPS = 120


def _robot_body(env):
    """pybullet body id of the articulated robot (the last multibody loadMJCF returned that has joints)."""
    p = env._p
    ids = [b for b in env.robot.objects if p.getNumJoints(b) > 0]
    return ids[-1] if len(ids) == 1 else ids[0]


def _state(p, body, nj, floating):
    st = {}
    if floating:
        pos, orn = p.getBasePositionAndOrientation(body)
        lin, ang = p.getBaseVelocity(body)
        st["base"] = list(pos) + list(orn) + list(ang) + list(lin)
    js = [p.getJointState(body, j) for j in range(nj)]
    st["q"] = [s[0] for s in js]
    st["qd"] = [s[1] for s in js]
    return st


def dump(path, envs=None, steps=STEPS, importer=None):
    """importer(module, cls) -> env class; default imports the reference."""
    import importlib
    out = {}
    for env_id, (mod, cls) in ENVS.items():
        if envs and env_id not in envs:
            continue
        if importer is None:
            klass = getattr(importlib.import_module("pybulletgym.envs.roboschool." + mod), cls)
        else:
            klass = importer(env_id, mod, cls)
        env = klass()
        (env._seed if hasattr(env, "_seed") else env.seed)(0)
        obs0 = (env._reset if hasattr(env, "_reset") else env.reset)()
        p = env._p
        body = _robot_body(env)
        nj = p.getNumJoints(body)
        links = []
        for j in range(-1, nj):
            d = p.getDynamicsInfo(body, j)
            rec = {"mass": d[0], "lateral_friction": d[1], "inertia_diag": list(d[2]), "inertial_pos": list(d[3]),
                   "inertial_orn": list(d[4]), "restitution": d[5], "rolling_friction": d[6], "spinning_friction": d[7]}
            if j >= 0:
                ji = p.getJointInfo(body, j)
                rec.update(joint_name=ji[1].decode(), joint_type=ji[2], damping=ji[6], friction=ji[7], lower=ji[8], upper=ji[9],
                           max_force=ji[10], max_velocity=ji[11], link_name=ji[12].decode(), axis=list(ji[13]),
                           parent_frame_pos=list(ji[14]), parent_frame_orn=list(ji[15]), parent_index=ji[16])
            links.append(rec)
        floating = links[0]["mass"] > 0
        rec = {"links": links, "floating": floating, "obs0": np.asarray(obs0, float).tolist(),
               "state0": _state(p, body, nj, floating), "steps": []}
        try:
            rec["engine"] = {k: (v if isinstance(v, (int, float)) else str(v)) for k, v in p.getPhysicsEngineParameters().items()}
        except Exception:
            rec["engine"] = {}
        rng = np.random.RandomState(1)
        nA = env.action_space.shape[0]
        step = env._step if hasattr(env, "_step") else env.step
        for t in range(steps):
            a = rng.uniform(-1, 1, nA)
            obs, rew, done, _ = step(a)
            cps = p.getContactPoints(body)
            rec["steps"].append({"a": a.tolist(), "obs": np.asarray(obs, float).tolist(), "reward": float(rew), "done": bool(done),
                                 "state": _state(p, body, nj, floating),
                                 "contacts": sorted({(int(c[3]), int(c[2]), int(c[4])) for c in cps})})
            if done:
                break
        out[env_id] = rec
        print("dumped %-40s links=%d steps=%d" % (env_id, len(links), len(rec["steps"])))
    with open(path, "w") as f:
        json.dump(out, f, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o))
    return out


# ------------------------------------------------------------------------------------------------------------------
def _canon(st, floating, dof_joint_idx):
    """dumped state -> the oracle's canonical layout [base pos3 quat4 omega3 vel3] + q + qd (movable joints only)."""
    q = [st["q"][j] for j in dof_joint_idx]
    qd = [st["qd"][j] for j in dof_joint_idx]
    return np.array((st["base"] if floating else []) + q + qd, float)


def compare(path, verbose=True):
    from oracle.oracle import OracleEnv
    from pybullet_gym_b200.mjcf import compiler as mj
    from pybullet_gym_b200.spec import SPECS
    data = json.load(open(path))
    report, table_mismatch = {}, 0
    for env_id, rec in data.items():
        spec = SPECS[env_id]
        bm = mj.parse_mjcf(spec.xml)
        L = bm.links
        rows = []
        if len(L) != len(rec["links"]):
            rows.append("link count: ours %d, pybullet %d" % (len(L), len(rec["links"])))
        for i, (l, d) in enumerate(zip(L, rec["links"])):
            def chk(what, ours, theirs, tol):
                ours, theirs = np.atleast_1d(np.asarray(ours, float)), np.atleast_1d(np.asarray(theirs, float))
                if ours.shape != theirs.shape or np.abs(ours - theirs).max() > tol * (1 + np.abs(theirs).max()):
                    rows.append("link %d %-18s %s: ours %s, pybullet %s" % (i, l.name, what, np.round(ours, 6), np.round(theirs, 6)))
            chk("mass", l.mass, d["mass"], 1e-6)
            chk("inertia diagonal", l.inertia, d["inertia_diag"], 1e-5)
            chk("inertial frame position", l.com, d["inertial_pos"], 1e-6)
            if i > 0:
                if l.name != d["link_name"] or l.joint_name != d["joint_name"]:
                    rows.append("link %d names: ours (%s, %s), pybullet (%s, %s)" % (i, l.name, l.joint_name, d["link_name"], d["joint_name"]))
                jt = {mj.JT_REVOLUTE: 0, mj.JT_PRISMATIC: 1, mj.JT_FIXED: 4}[l.jtype]
                if jt != d["joint_type"]:
                    rows.append("link %d joint type: ours %d, pybullet %d" % (i, jt, d["joint_type"]))
                if l.jtype != mj.JT_FIXED:
                    chk("axis", l.axis, d["axis"], 1e-6)
                    chk("limits", [l.lower, l.upper], [d["lower"], d["upper"]], 1e-6)
                    chk("joint damping", l.damping, d["damping"], 1e-9)
                if l.parent - 1 != d["parent_index"]:
                    rows.append("link %d parent: ours %d, pybullet %d" % (i, l.parent - 1, d["parent_index"]))
                if l.geoms:
                    chk("lateral friction", l.geoms[0].friction, d["lateral_friction"], 1e-6)
        table_mismatch += len(rows)
        # one-step replay from every dumped state
        dof_joint_idx = [i - 1 for i in bm.dof_links()]
        env = OracleEnv(env_id)
        env.reset(noise=np.zeros(spec.noise_dim))
        prev = _canon(rec["state0"], rec["floating"], dof_joint_idx)
        errs, obs_err = [], []
        for st in rec["steps"]:
            env.set_state(prev)
            env.physics_step(st["a"])
            ours = env.get_state()
            ref = _canon(st["state"], rec["floating"], dof_joint_idx)
            errs.append(np.abs(ours - ref).max() / (1.0))
            prev = ref
        errs = np.array(errs) if errs else np.zeros(1)
        report[env_id] = {"table_rows": rows, "one_step_max": float(errs.max()), "one_step_median": float(np.median(errs)),
                          "one_step_p95": float(np.quantile(errs, 0.95)), "steps": len(rec["steps"])}
        if verbose:
            print("== %s: %d table differences; one-step state error over %d steps: median %.3g, p95 %.3g, max %.3g" % (
                env_id, len(rows), len(rec["steps"]), np.median(errs), np.quantile(errs, 0.95), errs.max()))
            for r in rows[:40]:
                print("   ", r)
    return report, table_mismatch


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("dump", "compare"):
        print(__doc__)
        sys.exit(2)
    if sys.argv[1] == "dump":
        dump(sys.argv[2])
    else:
        _, bad = compare(sys.argv[2])
        sys.exit(1 if bad else 0)
