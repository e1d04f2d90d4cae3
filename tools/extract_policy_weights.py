#!/usr/bin/env python
"""Extract the inline MLP weights of the reference's pretrained-policy demos into tests/golden/policy_*.npz.

The reference's README calls these Roboschool-trained policies its "unit tests"
(/root/reference/README.md:6,89; pybulletgym/examples/roboschool-weights/enjoy_TF_*.py).  They are the
only behavioural fixture the reference has: a policy trained on a *different* simulator that still
walks is strong evidence the dynamics are in the right regime.  Only the numeric
arrays are extracted (float32); the demo code is not copied.  Run in the build container.
"""
import os
import re
import sys

import numpy as np

SRC = "/root/reference/pybulletgym/examples/roboschool-weights"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
FILES = {
    "InvertedPendulumPyBulletEnv-v0": "enjoy_TF_InvertedPendulumPyBulletEnv_v0_2017may.py",
    "InvertedPendulumSwingupPyBulletEnv-v0": "enjoy_TF_InvertedPendulumSwingupPyBulletEnv_v0_2017may.py",
    "InvertedDoublePendulumPyBulletEnv-v0": "enjoy_TF_InvertedDoublePendulumPyBulletEnv_v0_2017may.py",
    "ReacherPyBulletEnv-v0": "enjoy_TF_ReacherPyBulletEnv_v0_2017may.py",
    "HopperPyBulletEnv-v0": "enjoy_TF_HopperPyBulletEnv_v0_2017may.py",
    "Walker2DPyBulletEnv-v0": "enjoy_TF_Walker2DPyBulletEnv_v0_2017may.py",
    "HalfCheetahPyBulletEnv-v0": "enjoy_TF_HalfCheetahPyBulletEnv_v0_2017may.py",
    "AntPyBulletEnv-v0": "enjoy_TF_AntPyBulletEnv_v0_2017may.py",
    "HumanoidPyBulletEnv-v0": "enjoy_TF_HumanoidPyBulletEnv_v0_2017may.py",
    "HumanoidFlagrunPyBulletEnv-v0": "enjoy_TF_HumanoidFlagrunPyBulletEnv_v0_2017may.py",
    "HumanoidFlagrunHarderPyBulletEnv-v0": "enjoy_TF_HumanoidFlagrunHarderPyBulletEnv_v1_2017jul.py",
}


def main():
    for env_id, fn in FILES.items():
        txt = open(os.path.join(SRC, fn)).read()
        arrs = {}
        for m in re.finditer(r"^(weights_\w+)\s*=\s*np\.array\(\[(.*?)\]\)\s*$", txt, re.S | re.M):
            arrs[m.group(1)] = np.array(eval("[" + m.group(2) + "]", {"__builtins__": {}}), dtype=np.float32)
        # the only observation pre-processing any demo applies: `ob[0] += -1.4 + 0.8` in SmallReactivePolicy.act of the Humanoid
        # and HumanoidFlagrun demos (enjoy_TF_HumanoidPyBulletEnv_v0_2017may.py:27) -- the policies were trained with
        # initial_z = 1.4, the env reports z - 0.8
        shift = np.zeros(arrs["weights_dense1_w"].shape[0], dtype=np.float32)
        m = re.search(r"ob\[0\]\s*\+=\s*(-?[\d.]+)\s*\+\s*(-?[\d.]+)", txt)
        if m:
            shift[0] = float(m.group(1)) + float(m.group(2))
        keys = ["weights_dense1_w", "weights_dense1_b", "weights_dense2_w", "weights_dense2_b", "weights_final_w", "weights_final_b"]
        assert all(k in arrs for k in keys), (fn, list(arrs))
        path = os.path.join(OUT, "policy_%s.npz" % env_id.split("PyBullet")[0])
        np.savez_compressed(path, obs_shift=shift, **{k[8:]: arrs[k] for k in keys})
        print(env_id, [arrs[k].shape for k in keys[::2]], "->", os.path.basename(path), os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
