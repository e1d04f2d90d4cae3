#!/usr/bin/env python
"""Run one of the reference's pretrained policies on the batched backend, thousands of episodes at a time.

Counterpart of the reference's `pybulletgym/examples/roboschool-weights/enjoy_TF_<env>_2017may.py` scripts (one env, one
episode, rendered).  The weights those scripts embed were extracted once into `tests/golden/policy_<Name>.npz`
(`tools/extract_policy_weights.py`); here the two-hidden-layer ReLU policy runs inside the step kernel (`pbg_set_policy` /
`pbg_rollout_policy`), so a whole 1000-step episode of every env is a single launch.

    python examples/enjoy.py Ant --envs 4096
    python examples/enjoy.py Hopper --envs 4096 --host-policy     # policy as torch matmuls between step() calls instead

Envs restart inside the kernel when their episode ends; the script prints the mean return and length of the episodes that
finished (device-side statistics, `pbg_stats`) and the env-steps per second.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from pybullet_gym_b200 import VectorEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name", help="Ant, Hopper, Humanoid, InvertedDoublePendulum, ... (tests/golden/policy_<name>.npz)")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--host-policy", action="store_true", help="evaluate the policy with torch between step() calls")
    a = ap.parse_args()
    w = np.load(os.path.join(ROOT, "tests", "golden", "policy_%s.npz" % a.name))
    env = VectorEnv(a.name + "PyBulletEnv-v0", a.envs, device="cuda:0", seed=a.seed, auto_reset=True)
    obs = env.reset()
    shift = w["obs_shift"] if "obs_shift" in w.files else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if a.host_policy:
        W = [torch.as_tensor(w[k], device="cuda") for k in ("dense1_w", "dense1_b", "dense2_w", "dense2_b", "final_w", "final_b")]
        sh = torch.as_tensor(shift, device="cuda", dtype=torch.float32) if shift is not None else 0.0
        for t in range(a.steps):
            x = torch.relu((obs + sh) @ W[0] + W[1])
            x = torch.relu(x @ W[2] + W[3])
            obs, r, d, _ = env.step(x @ W[4] + W[5])
    else:
        env.set_policy(w["dense1_w"], w["dense1_b"], w["dense2_w"], w["dense2_b"], w["final_w"], w["final_b"], obs_shift=shift)
        env.rollout_policy(a.steps)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    st = env.stats()
    n = max(st["episodes"], 1)
    print("%s: %d envs x %d steps, %d episodes finished: mean return %.1f, mean length %.1f   (%.2e env-steps/s, %s)" % (
        a.name, a.envs, a.steps, st["episodes"], st["return_sum"] / n, st["length_sum"] / n, a.steps * a.envs / el,
        "torch policy + step()" if a.host_policy else "policy fused into the step kernel"))


if __name__ == "__main__":
    main()
