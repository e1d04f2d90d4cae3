"""Development: find a state where CUDA and oracle disagree after one env step, then localise the substep."""
import sys, os, dataclasses, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from pybullet_gym_b200.vector_env import VectorEnv
from pybullet_gym_b200 import _lib
from pybullet_gym_b200.spec import SPECS
from oracle.oracle import OracleEnv
np.set_printoptions(precision=5, suppress=True, linewidth=220)
eid = sys.argv[1]
E = 64
spec = SPECS[eid]
spec1 = dataclasses.replace(spec, scene=dataclasses.replace(spec.scene, frame_skip=1))
mc = _lib.lib().pbg_max_contacts(spec.kind)
rng = np.random.default_rng(0)
env = VectorEnv(eid, E, auto_reset=False); env.reset()
env1 = VectorEnv(eid, 2, auto_reset=False, spec=spec1); env1.reset()
nA = env.action_dim
noise = rng.uniform(-0.1, 0.1, (E, nA)).astype(np.float32)
orcs = [OracleEnv(eid, max_contacts=mc) for _ in range(E)]
for i, o in enumerate(orcs):
    o.reset(noise=noise[i].astype(np.float64), floor_in_parts=False)
L = _lib.lib()
L.pbg_dev_physics_step_rows.argtypes = [C.c_void_p] * 2 + [C.c_int32] + [C.c_void_p] * 2
found = 0
for t in range(40):
    a = rng.uniform(-1, 1, (E, nA)).astype(np.float32)
    ost = np.stack([o.get_state() for o in orcs]).astype(np.float32)
    env.set_state(torch.from_numpy(ost))
    for i, o in enumerate(orcs):
        o.set_state(ost[i].astype(np.float64))
    env.physics_step(torch.from_numpy(a))
    gst = env.get_state().cpu().numpy()
    for i, o in enumerate(orcs):
        o.physics_step(a[i].astype(np.float64))
    ost2 = np.stack([o.get_state() for o in orcs])
    rel = np.abs(gst - ost2) / (1 + np.abs(ost2))
    bad = np.where(rel.max(axis=1) > 0.05)[0]
    for i in bad[:1]:
        found += 1
        print("=== t", t, "env", i, "rel err", rel[i].max(), "idx", rel[i].argmax())
        # substep by substep from the same start state
        o1 = OracleEnv(spec1, max_contacts=mc); o1.reset(noise=np.zeros(nA)); o1.set_state(ost[i].astype(np.float64))
        env1.set_state(torch.from_numpy(np.stack([ost[i], ost[i]])))
        # note: torque is recomputed every call here but damping is 0, so 4 x frame_skip=1 == 1 x frame_skip=4
        for sub in range(4):
            o1.physics_step(a[i].astype(np.float64))
            dbg = torch.zeros(512, device="cuda")
            av = torch.from_numpy(np.stack([a[i], a[i]])).cuda()
            L.pbg_dev_physics_step_rows(env1._h, C.c_void_p(av.data_ptr()), 0, C.c_void_p(dbg.data_ptr()), None)
            torch.cuda.synchronize()
            d = dbg.cpu().numpy(); gnl, gnc = int(d[0]), int(d[1])
            g = d[2:2 + 4 * (gnl + 3 * gnc)].reshape(-1, 4)
            nl, nc, rows = o1.rows()
            g1 = env1.get_state().cpu().numpy()[0]; s1 = o1.get_state()
            r1 = np.abs(g1 - s1) / (1 + np.abs(s1))
            print(" substep", sub, "orc nl nc", nl, nc, "gpu", gnl, gnc, "state rel err %.3g at %d" % (r1.max(), r1.argmax()),
                  "max|u| orc %.1f" % np.abs(s1[-nA:]).max())
            if r1.max() > 1e-3:
                for k in range(max(len(rows), len(g))):
                    print("   row", k, rows[k] if k < len(rows) else None, "|", g[k] if k < len(g) else None)
                nu = o1.model.nu
                u0, uf, dv = o1.debug()
                print("   orc u0   ", u0); print("   gpu u0   ", d[380:380 + nu])
                print("   orc h*qdd", uf - u0); print("   gpu du   ", d[340:340 + nu])
                print("   orc dv   ", dv); print("   orc total", uf - u0 + dv)
                print("   gpu u", g1[7:13] if spec.kind >= 5 else "", g1[-nA:])
                print("   orc u", s1[7:13] if spec.kind >= 5 else "", s1[-nA:])
                break
    if found >= 3:
        break
