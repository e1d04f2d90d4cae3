"""Per-warp duration of one steady-state launch (needs a -DPBG_PHASE_CLOCKS build): who is the slowest warp / CTA / SM?"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import Workload
env_id, E = sys.argv[1], int(sys.argv[2])
wl = Workload(env_id, E, torch.device("cuda", 0), 0, 1, 1000)
from pybullet_gym_b200 import _lib
L = _lib.lib(); L.pbg_debug_phases.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
buf = (C.c_uint * (4096 * 3))()
W = int(sys.argv[3]) if len(sys.argv) > 3 else 14
allw = []
for rep in range(6):
    for _ in range(7): wl.step()
    L.pbg_debug_phases(wl.env._h, buf, 2)
    a = np.array(buf[:], dtype=np.int64).reshape(-1, 3)
    nw = (E + 2 * W - 1) // (2 * W) * W if W else 0
    a = a[:nw]
    cyc, smid, ovf = a[:, 0], a[:, 1], a[:, 2]
    cta = cyc.reshape(-1, W)
    print("launch %d: warp cycles mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f | CTA max: mean %.0f min %.0f max %.0f | CTA mean-of-warps: min %.0f max %.0f"
          % (rep, cyc.mean(), np.median(cyc), np.quantile(cyc, .9), np.quantile(cyc, .99), cyc.max(), cta.max(1).mean(), cta.max(1).min(), cta.max(1).max(),
             cta.mean(1).min(), cta.mean(1).max()))
    allw.append((cta.max(1), smid.reshape(-1, W)[:, 0]))
    nov = ovf - 100
    for v in range(-3, 11):
        sel = nov == v
        if sel.sum() > 3:
            print("     warps with max rows-beyond-LPE %2d: %4d  mean cycles %.0f  p90 %.0f" % (v, sel.sum(), cyc[sel].mean(), np.quantile(cyc[sel], .9)))
m = np.stack([x[0] for x in allw]); sm = allw[0][1]
print("per-CTA max-warp cycles, correlation between launches (same CTA slow again?):", np.corrcoef(m)[0, 1:].round(2))
order = np.argsort(m.mean(0))
print("slowest CTAs (idx, smid, mean cycles):", [(int(i), int(sm[i]), int(m.mean(0)[i])) for i in order[-8:]])
print("fastest CTAs:", [(int(i), int(sm[i]), int(m.mean(0)[i])) for i in order[:8]])
