"""Development: compare the constraint rows of one substep, CUDA vs oracle, on an Ant state in contact."""
import sys, os, dataclasses, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from pybullet_gym_b200.vector_env import VectorEnv
from pybullet_gym_b200 import _lib
from pybullet_gym_b200.spec import SPECS, SceneSpec
from oracle.oracle import OracleEnv
np.set_printoptions(precision=6, suppress=True, linewidth=220)
eid = sys.argv[1] if len(sys.argv) > 1 else "AntPyBulletEnv-v0"
spec1 = dataclasses.replace(SPECS[eid], scene=dataclasses.replace(SPECS[eid].scene, frame_skip=1))
mc = _lib.lib().pbg_max_contacts(spec1.kind)
rng = np.random.default_rng(0)
full = OracleEnv(eid, max_contacts=mc)
full.reset(noise=rng.uniform(-.1, .1, full.nact))
for t in range(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    full.physics_step(rng.uniform(-1, 1, full.nact))
s0 = full.get_state().astype(np.float32)
a = rng.uniform(-1, 1, (1, full.nact)).astype(np.float32)
orc = OracleEnv(spec1, max_contacts=mc)
orc.reset(noise=np.zeros(full.nact)); orc.set_state(s0.astype(np.float64))
orc.physics_step(a[0].astype(np.float64))
nl, nc, rows = orc.rows()
env = VectorEnv(eid, 2, auto_reset=False, spec=spec1)
env.reset()
env.set_state(torch.from_numpy(np.stack([s0, s0])))
dbg = torch.zeros(2 + 4 * 64, device="cuda")
L = _lib.lib()
L.pbg_dev_physics_step_rows.argtypes = [C.c_void_p] * 2 + [C.c_int32] + [C.c_void_p] * 2
rc = L.pbg_dev_physics_step_rows(env._h, C.c_void_p(torch.from_numpy(np.repeat(a, 2, 0)).cuda().data_ptr()), 0, C.c_void_p(dbg.data_ptr()), None)
torch.cuda.synchronize()
d = dbg.cpu().numpy()
gnl, gnc = int(d[0]), int(d[1])
g = d[2:2 + 4 * (gnl + 3 * gnc)].reshape(-1, 4)
print("oracle nl nc", nl, nc, " gpu", gnl, gnc)
print("row   orc(rhs dinv lam) | gpu(rhs dinv lam resid)")
for i in range(max(len(rows), len(g))):
    print(i, rows[i] if i < len(rows) else None, "|", g[i] if i < len(g) else None)
gs = env.get_state().cpu().numpy()[0]
os_ = orc.get_state()
print("state diff", np.abs(gs - os_).max(), "\n gpu", gs[7:], "\n orc", os_[7:])
