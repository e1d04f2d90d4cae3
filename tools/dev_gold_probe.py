"""Development: per-step error of the CUDA path on one golden episode against the oracle twins' drift.  usage: dev_gold_probe.py [ENV] [EPISODE]"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import oracle as oracle_lib
from pybullet_gym_b200.vector_env import VectorEnv
name = sys.argv[1] if len(sys.argv) > 1 else "HalfCheetah"
ei = int(sys.argv[2]) if len(sys.argv) > 2 else 3
env_id = name + "PyBulletEnv-v0"
g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "task_%s.json" % name)))
env = VectorEnv(env_id, 1, device="cuda:0", seed=0, auto_reset=False)
rng = np.random.default_rng(23)
ep = g["episodes"][ei]
noise = np.array(ep["noise"], np.float64)
env.reset(joint_noise=torch.tensor([ep["noise"]], dtype=torch.float32), floor_in_parts=ei > 0)
tw7 = [oracle_lib.OracleEnv(env_id) for _ in range(8)]
tw6 = [oracle_lib.OracleEnv(env_id) for _ in range(8)]
for tw in tw7: tw.reset(noise=noise + rng.normal(size=noise.shape) * 1e-7, floor_in_parts=ei > 0)
for tw in tw6: tw.reset(noise=noise + rng.normal(size=noise.shape) * 1e-6, floor_in_parts=ei > 0)
for t, st in enumerate(ep["steps"]):
    a = np.array(st["a"], np.float64); gold = np.array(st["obs"])
    obs, rew, done, info = env.step(torch.tensor([st["a"]], dtype=torch.float32))
    d7 = [np.abs(tw.step(a)[0] - gold).max() for tw in tw7]
    d6 = [np.abs(tw.step(a)[0] - gold).max() for tw in tw6]
    e = np.abs(obs.cpu().numpy()[0] - gold)
    print("t %2d err %.2e (obs[%d])  1e-7 twins: med %.1e max %.1e   1e-6 twins: med %.1e max %.1e   feet %s done %s/%s" % (
        t, e.max(), int(e.argmax()), np.median(d7), max(d7), np.median(d6), max(d6), gold[-6:].astype(int), bool(done[0]), st["done"]))
