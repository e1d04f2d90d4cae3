"""Development check: CUDA path vs CPU oracle, per env id (run on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
from pybullet_gym_b200.vector_env import VectorEnv
from pybullet_gym_b200 import _lib
from oracle.oracle import OracleEnv

np.set_printoptions(precision=5, suppress=True, linewidth=200)
THR = "--thr" in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith("--")]
ids = args or ["InvertedPendulumPyBulletEnv-v0", "AntPyBulletEnv-v0", "HopperPyBulletEnv-v0",
                       "Walker2DPyBulletEnv-v0", "HalfCheetahPyBulletEnv-v0", "HumanoidPyBulletEnv-v0"]
E = 64
for eid in ids:
    print("=====", eid, flush=True)
    env = VectorEnv(eid, E, seed=1, auto_reset=False)
    mc = _lib.lib().pbg_max_contacts(env.spec.kind)
    rng = np.random.default_rng(0)
    nA = env.action_dim
    noise = rng.uniform(-0.1, 0.1, (E, nA)).astype(np.float32)
    obs = env.reset(joint_noise=torch.from_numpy(noise)).cpu().numpy()
    orcs = [OracleEnv(eid, max_contacts=mc) for _ in range(E)]
    oobs = np.stack([o.reset(noise=noise[i].astype(np.float64), floor_in_parts=False) for i, o in enumerate(orcs)])
    print("reset obs maxdiff", np.abs(obs - oobs).max())
    st = env.get_state().cpu().numpy()
    ost = np.stack([o.get_state() for o in orcs])
    print("reset state maxdiff", np.abs(st - ost).max())
    # single-step parity from identical states along an oracle trajectory
    worst = 0
    T = 60
    alive = np.ones(E, bool)
    for t in range(T):
        a = rng.uniform(-1, 1, (E, nA)).astype(np.float32)
        # sync GPU state to oracle state (float32)
        ost = np.stack([o.get_state() for o in orcs]).astype(np.float32)
        env.set_state(torch.from_numpy(ost))
        for i, o in enumerate(orcs):
            o.set_state(ost[i].astype(np.float64))
        nct = env.physics_step(torch.from_numpy(a), want_contacts=True).cpu().numpy()
        gst = env.get_state().cpu().numpy()
        for i, o in enumerate(orcs):
            o.physics_step(a[i].astype(np.float64))
        ost2 = np.stack([o.get_state() for o in orcs])
        onct = np.array([o.num_contacts() for o in orcs])
        d = np.abs(gst - ost2)
        rel = d / (1.0 + np.abs(ost2))
        w = rel.max()
        worst = max(worst, w)
        if t % 10 == 0 or w > 1e-3:
            i, j = np.unravel_index(rel.argmax(), rel.shape)
            print("t", t, "single-step max rel diff %.3g" % w, "at env", i, "idx", j, "gpu", gst[i, j], "orc", ost2[i, j],
                  "nct gpu/orc", nct[i], onct[i], "contact count mismatch:", int((nct != onct).sum()))
    print("worst single-step rel diff", worst)
    # observe parity on oracle states
    a = rng.uniform(-1.5, 1.5, (E, nA)).astype(np.float32)
    ost = np.stack([o.get_state() for o in orcs]).astype(np.float32)
    env.set_state(torch.from_numpy(ost))
    for i, o in enumerate(orcs):
        o.set_state(ost[i].astype(np.float64))
    gobs, grew, gdone, gterms = [x.cpu().numpy() for x in env.observe(torch.from_numpy(a))]
    res = [o.observe(a[i].astype(np.float64)) for i, o in enumerate(orcs)]
    oobs = np.stack([r[0] for r in res]); orew = np.array([r[1] for r in res]); oterms = np.stack([r[3] for r in res])
    print("observe: obs maxdiff", np.abs(gobs - oobs).max(), "terms maxdiff (excl progress)", np.abs(gterms - oterms)[:, [0, 2, 3, 4]].max())
    # free-running trajectory comparison
    env2 = VectorEnv(eid, E, seed=1, auto_reset=False)
    env2.reset(joint_noise=torch.from_numpy(noise))
    orcs = [OracleEnv(eid, max_contacts=mc) for _ in range(E)]
    for i, o in enumerate(orcs):
        o.reset(noise=noise[i].astype(np.float64), floor_in_parts=False)
    for t in range(100):
        a = rng.uniform(-1, 1, (E, nA)).astype(np.float32)
        o_g, r_g, d_g, info = env2.step(torch.from_numpy(a))
        o_g = o_g.cpu().numpy(); r_g = r_g.cpu().numpy(); d_g = d_g.cpu().numpy()
        res = [o.step(a[i].astype(np.float64)) for i, o in enumerate(orcs)]
        o_o = np.stack([r[0] for r in res]); r_o = np.array([r[1] for r in res]); d_o = np.array([r[2] for r in res])
        if t in (0, 1, 2, 5, 10, 20, 50, 99):
            print("traj t", t, "obs maxdiff %.3g" % np.abs(o_g - o_o).max(), "rew maxdiff %.3g" % np.abs(r_g - r_o).max(),
                  "done mismatch", int((d_g.astype(bool) != d_o).sum()), "done frac", d_o.mean())
    if not THR:
        continue
    # throughput smoke
    big = VectorEnv(eid, 4096, seed=0)
    big.reset()
    acts = torch.rand(4096, nA, device="cuda") * 2 - 1
    for _ in range(20):
        big.step_fast(acts)
    torch.cuda.synchronize()
    t0 = time.time()
    n = 200
    for _ in range(n):
        big.step_fast(acts)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("throughput 4096 envs: %.3g env-steps/s (%.1f us/step)" % (4096 * n / dt, 1e6 * dt / n), big.stats())
