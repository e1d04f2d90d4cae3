"""Scratch diagnostics for round 2 (GPU): FlagrunHarder T2 outliers, T3 per-component drift."""
import dataclasses, sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import oracle as oracle_lib
from pybullet_gym_b200 import _lib
from pybullet_gym_b200.spec import SPECS
from pybullet_gym_b200.vector_env import VectorEnv
sys.path.insert(0, "tests")
from test_gpu_parity import _throw_cube_at_robots, _rel

E = 48
env_id = "HumanoidFlagrunHarderPyBulletEnv-v0"
spec = SPECS[env_id]
spec1 = dataclasses.replace(spec, scene=dataclasses.replace(spec.scene, frame_skip=1))
env1 = VectorEnv(env_id, E, device="cuda:0", seed=1, auto_reset=False, spec=spec1)
mc = _lib.lib().pbg_max_contacts(spec.kind)
orcs = [oracle_lib.OracleEnv(env_id, seed=1, env_index=i, max_contacts=mc) for i in range(E)]
orcs1 = [oracle_lib.OracleEnv(spec1, seed=1, env_index=i, max_contacts=mc) for i in range(E)]
rng = np.random.default_rng(5)
noise = rng.uniform(-0.1, 0.1, (E, env1.noise_dim)).astype(np.float32)
env1.reset(joint_noise=torch.from_numpy(noise))
for i in range(E):
    orcs[i].reset(noise=noise[i].astype(np.float64)); orcs1[i].reset(noise=noise[i].astype(np.float64))
bad = []
for t in range(40):
    a = rng.uniform(-1, 1, (E, 17)).astype(np.float32)
    ost = np.stack([o.get_state() for o in orcs]).astype(np.float32)
    if t % 4 == 1:
        ost = _throw_cube_at_robots(ost, rng).astype(np.float32)
    env1.set_state(torch.from_numpy(ost))
    n1 = env1.physics_step(torch.from_numpy(a), want_contacts=True).cpu().numpy()
    g1 = env1.get_state().cpu().numpy()
    for i in range(E):
        si = ost[i].astype(np.float64)
        orcs[i].set_state(si); orcs1[i].set_state(si)
        orcs[i].physics_step(a[i].astype(np.float64)); orcs1[i].physics_step(a[i].astype(np.float64))
        o1 = orcs1[i].get_state()
        e = _rel(g1[i], o1)
        if e.max() > 1e-2:
            k = int(np.argmax(e))
            print("t %d env %d err %.3g at comp %d  gpu nc %d oracle nc %d  gpu %.4g orc %.4g" % (t, i, e.max(), k, n1[i], orcs1[i].num_contacts(), g1[i][k], o1[k]))
            bad.append((ost[i], a[i], g1[i], o1))
np.savez("gpurun_out/dbg_harder.npz", st=np.array([b[0] for b in bad]), a=np.array([b[1] for b in bad]), g=np.array([b[2] for b in bad]), o=np.array([b[3] for b in bad]))

# ---- T3 per component
env_id = "InvertedPendulumSwingupPyBulletEnv-v0"
n = 16
env = VectorEnv(env_id, n, device="cuda:0", seed=1, auto_reset=False)
rng = np.random.default_rng(11)
noise = rng.uniform(-0.1, 0.1, (n, 1)).astype(np.float32)
env.reset(joint_noise=torch.from_numpy(noise))
orcs = [oracle_lib.OracleEnv(env_id) for _ in range(n)]
twins = [oracle_lib.OracleEnv(env_id) for _ in range(n)]
for i, o in enumerate(orcs):
    o.reset(noise=noise[i].astype(np.float64)); twins[i].reset(noise=noise[i].astype(np.float64) + 1e-7)
tape = rng.uniform(-1, 1, (1000, n, 1)).astype(np.float32) * 0.3
for t in range(1000):
    obs, rew, done, _ = env.step(torch.from_numpy(tape[t]))
    res = np.stack([o.step(tape[t, i].astype(np.float64))[0] for i, o in enumerate(orcs)])
    rtw = np.stack([o.step(tape[t, i].astype(np.float64))[0] for i, o in enumerate(twins)])
    if t in (99, 299, 499, 999):
        g = obs.cpu().numpy()
        print("T3 t", t, "gpu err per comp", np.abs(g - res).max(axis=0), "twin drift", np.abs(rtw - res).max(axis=0), "max |obs|", np.abs(res).max(axis=0))
