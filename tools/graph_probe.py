#!/usr/bin/env python
"""CUDA-graph capture of a block of env steps (launch-bound small envs): pbg_step only enqueues one kernel on the caller's
stream, so torch.cuda.graph can capture K steps and replay them with one launch.  Prints eager vs graph-replay throughput and
checks that the replay is bit-identical to eager stepping."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from pybullet_gym_b200 import VectorEnv

def run(env_id, E, K=64, reps=40):
    mk = lambda: VectorEnv(env_id, E, device="cuda:0", seed=1, auto_reset=True)
    eager, graphed = mk(), mk()
    eager.reset(); graphed.reset()
    acts = torch.rand(K, E, eager.action_dim, device="cuda") * 2 - 1
    obs_log = torch.empty(K, E, eager.obs_dim, device="cuda"); rew_log = torch.empty(K, E, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        snap = graphed.snapshot()
        with torch.cuda.graph(g, stream=s):
            for k in range(K):
                o, r, d = graphed.step_fast(acts[k])
                obs_log[k].copy_(o); rew_log[k].copy_(r)
        graphed.restore(snap)                      # capture does not execute; start both from the same state
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        ref_o = torch.empty_like(obs_log); ref_r = torch.empty_like(rew_log)
        for k in range(K):
            o, r, d = eager.step_fast(acts[k]); ref_o[k].copy_(o); ref_r[k].copy_(r)
        torch.cuda.synchronize()
        same = torch.equal(ref_o, obs_log) and torch.equal(ref_r, rew_log)
        t0 = time.perf_counter()
        for _ in range(reps): g.replay()
        torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / (reps * K)
        t0 = time.perf_counter()
        for _ in range(reps):
            for k in range(K):
                o, r, d = eager.step_fast(acts[k]); ref_o[k].copy_(o); ref_r[k].copy_(r)
        torch.cuda.synchronize(); te = (time.perf_counter() - t0) / (reps * K)
    print("%-38s E=%5d  eager %6.1f us/step (%.2e env-steps/s)   graph of %d steps %6.1f us/step (%.2e)   bit-identical: %s" % (
        env_id, E, te * 1e6, E / te, K, tg * 1e6, E / tg, same))

if __name__ == "__main__":
    for env_id, E in (("InvertedPendulumPyBulletEnv-v0", 4096), ("InvertedDoublePendulumPyBulletEnv-v0", 4096),
                      ("ReacherPyBulletEnv-v0", 4096), ("HopperPyBulletEnv-v0", 1024), ("AntPyBulletEnv-v0", 4096)):
        run(env_id, E)
