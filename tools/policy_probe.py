import numpy as np, glob, os, dataclasses, sys, itertools
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.oracle import OracleEnv
from pybullet_gym_b200.spec import SPECS
from pybullet_gym_b200.mjcf import compiler as mj
def act(w, ob):
    ob=ob+w["obs_shift"]      # the Humanoid demos' `ob[0] += -1.4 + 0.8` (tools/extract_policy_weights.py)
    x=np.maximum(ob@w["dense1_w"]+w["dense1_b"],0); x=np.maximum(x@w["dense2_w"]+w["dense2_b"],0); return x@w["final_w"]+w["final_b"]
def run(name, rules=None, scene_kw=None, eps=2, T=1000, **orc_kw):
    eid=name+"PyBulletEnv-v0"; w=np.load("tests/golden/policy_%s.npz"%name)
    spec=SPECS[eid]
    if scene_kw: spec=dataclasses.replace(spec, scene=dataclasses.replace(spec.scene, **scene_kw))
    bm=mj.parse_mjcf(spec.xml, rules) if rules else None
    out=[]
    for ep in range(eps):
        e=OracleEnv(spec, seed=5, env_index=ep, bm=bm, **orc_kw); ob=e.reset(); score=0; n=0
        for t in range(T):
            ob,r,d,_=e.step(act(w,ob)); score+=r; n+=1
            if d: break
        out.append((round(score),n))
    return out
names=["Hopper","Walker2D","HalfCheetah","Ant","Humanoid"]
configs={
 "base": (None,None),
 "jdamp": (mj.ImporterRules(joint_damping_from_mjcf=True),None),
 "nolinkdamp": (mj.ImporterRules(link_damping=0.0),None),
 "warm0.85": (None,dict(warmstarting_factor=0.85)),
 "erp_contact0.2": (None,dict(contact_erp=0.2)),
 "split": (None,dict(limit_split_impulse=True)),
 # fillMultiBodyConstraint may read m_erp2 (= setDefaultContactERP 0.9) instead of m_erp (0.2) for the joint-limit rows
 "erp_limit0.9": (None,dict(erp=0.9)),
 "erp_limit0.9_nosplit": (None,dict(erp=0.9, limit_split_impulse=False)),
}
names_all=["Hopper","Walker2D","HalfCheetah","Ant","Humanoid","HumanoidFlagrun"]
sel=[c for c in (sys.argv[1:] or list(configs)) if c in configs]
for c in sel:
    r,s=configs[c]
    print(c, {n: run(n,r,s) for n in (names_all if c.startswith("erp_limit") or c == "base" else names)}, flush=True)

# torque-scale probe
import sys
if "torque" not in sys.argv: sys.argv.append("skiptorque")
print("--- torque scale")
for sc in ((0.25, 0.5, 2.0) if "torque" in sys.argv else ()):
    res = {}
    for n in names:
        eid=n+"PyBulletEnv-v0"; w=np.load("tests/golden/policy_%s.npz"%n)
        spec=dataclasses.replace(SPECS[eid], power=SPECS[eid].power*sc)
        out=[]
        for ep in range(2):
            e=OracleEnv(spec, seed=5, env_index=ep); ob=e.reset(); score=0; k=0
            for t in range(1000):
                ob,r,d,_=e.step(act(w,ob)); score+=r; k+=1
                if d: break
            out.append((round(score),k))
        res[n]=out
    print("power x%.2f"%sc, res, flush=True)

print("--- torsional friction rows")
for kw in (dict(torsional_friction=True), dict(torsional_friction=True, warmstarting_factor=0.85)):
    print(kw, {n: run(n, None, kw, eps=3) for n in names}, flush=True)

# persistent floor manifolds (oracle-only switch `ground_manifold`): one new point per collision pass,
# up to four cached points per geom, instead of the instantaneous end-sphere candidates
if "manifold" in sys.argv:
    print("--- persistent floor manifolds")
    for kw in (dict(), dict(ground_manifold=1)):
        print(kw or "candidates (default)", {n: run(n, None, None, eps=4, **kw) for n in names_all}, flush=True)
