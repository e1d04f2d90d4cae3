"""Host-side time per VectorEnv.step_host call (python wrapper + C ABI + kernel + PCIe), Ant 4096."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from pybullet_gym_b200.vector_env import VectorEnv, _ptr
E = 4096
env = VectorEnv("AntPyBulletEnv-v0", E, device="cuda:0", seed=0, auto_reset=True); env.reset()
a = [(torch.rand(E, 8) * 2 - 1).pin_memory() for _ in range(8)]
obs = torch.empty(E, 28).pin_memory(); rew = torch.empty(E).pin_memory(); done = torch.empty(E, dtype=torch.uint8).pin_memory()
for i in range(300): env.step_fast(a[i % 8].cuda())
torch.cuda.synchronize()
def timeit(f, n=600):
    for i in range(50): f(i)
    t = time.perf_counter()
    for i in range(n): f(i)
    return (time.perf_counter() - t) / n * 1e6
L, h = env._L, env._h
raw = [(_ptr(x), _ptr(obs), _ptr(rew), _ptr(done)) for x in a]
print("VectorEnv.step_host          %.1f us" % timeit(lambda i: env.step_host(a[i % 8], obs, rew, done)))
print("raw pbg_step_host (ctypes)   %.1f us" % timeit(lambda i: L.pbg_step_host(h, *raw[i % 8])))
print("VectorEnv.step_host          %.1f us" % timeit(lambda i: env.step_host(a[i % 8], obs, rew, done)))
