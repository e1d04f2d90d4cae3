#!/usr/bin/env python
"""A/B timing of experiment builds on ONE box: python tools/ab.py ENV_ID ENVS lib1 lib2 ... [--rounds R] [--phases]
Each lib is timed in its own process (PBG_LIB), round-robin, steady-state pre-rolled batch, L2-resident back-to-back steps."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import os, sys, json, ctypes as C
sys.path.insert(0, %r)
import torch
from bench import Workload
env_id, E, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
wl = Workload(env_id, E, dev, 0, 1, 1000)
for _ in range(20): wl.step()
torch.cuda.synchronize()
from pybullet_gym_b200 import _lib
L = _lib.lib()
ph = (C.c_ulonglong * 32)()
has = hasattr(L, "pbg_debug_phases")
if has: L.pbg_debug_phases.argtypes = [C.c_void_p, C.c_void_p, C.c_int]; L.pbg_debug_phases(wl.env._h, ph, 1)
best = []
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): wl.step()
    b.record(); torch.cuda.synchronize()
    best.append(E * steps / (a.elapsed_time(b) * 1e-3))
if has: L.pbg_debug_phases(wl.env._h, ph, 0)
print(json.dumps({"rate": sorted(best)[1], "phases": list(ph)}))
''' % ROOT


def main():
    argv = sys.argv[1:]
    for opt in ("--rounds", "--steps"):
        if opt in argv:
            k = argv.index(opt); del argv[k:k + 2]
    args = [a for a in argv if not a.startswith("--")]
    env_id, E, libs = args[0], int(args[1]), args[2:]
    rounds = int(sys.argv[sys.argv.index("--rounds") + 1]) if "--rounds" in sys.argv else 2
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 300
    res = {l: [] for l in libs}
    phases = {}
    for r in range(rounds):
        for l in libs:
            env = dict(os.environ, PBG_LIB=os.path.join(ROOT, l) if not os.path.isabs(l) else l)
            out = subprocess.run([sys.executable, "-c", WORKER, env_id, str(E), str(steps)], env=env, capture_output=True, text=True)
            if out.returncode:
                print(l, "FAILED", out.stderr[-800:]); continue
            d = json.loads(out.stdout.strip().splitlines()[-1])
            res[l].append(d["rate"]); phases[l] = d["phases"]
    for l in libs:
        print("%-40s %s  rates: %s" % (l, env_id.split("PyBullet")[0], " ".join("%.3e" % x for x in res[l])))
        if "--phases" in sys.argv and sum(phases.get(l, [0])) > 0:
            names = ["barrier", "fk", "collide", "wrench+sums", "S/H+M", "cholesky", "freevel+limits", "rows", "delassus", "pgs", "du+integrate", "task", "load", "store"]
            tot = float(sum(phases[l]))
            print("   " + "  ".join("%s %.1f%%" % (n, 100 * p / tot) for n, p in zip(names, phases[l])))


if __name__ == "__main__":
    main()
