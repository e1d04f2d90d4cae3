#!/usr/bin/env python
"""Where the end-to-end step time goes: times one Ant step of 4096 envs on cuda:0 through the device API
back to back, with a stream synchronisation per step, and through pbg_step_host with all / some / none of the host outputs."""
import ctypes as C, time, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from pybullet_gym_b200.vector_env import VectorEnv, _ptr
E=4096; K=600
env=VectorEnv("AntPyBulletEnv-v0",E,device="cuda:0",seed=0,auto_reset=True); env.reset()
L,h=env._L,env._h
acts_d=torch.rand(64,E,8,device="cuda")*2-1
acts_h=[(torch.rand(E,8)*2-1).pin_memory() for _ in range(8)]
obs=torch.empty(E,28).pin_memory(); rew=torch.empty(E).pin_memory(); done=torch.empty(E,dtype=torch.uint8).pin_memory()
def timeit(f,n=K):
    for i in range(50): f(i)
    torch.cuda.synchronize(); t=time.perf_counter()
    for i in range(n): f(i)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e6
r={}
r["device back-to-back (no sync)"]=timeit(lambda i: env.step_fast(acts_d[i%64]))
def f(i): env.step_fast(acts_d[i%64]); torch.cuda.current_stream().synchronize()
r["device + stream sync each step"]=timeit(f)
r["step_host zero-copy full"]=timeit(lambda i: env.step_host(acts_h[i%8],obs,rew,done))
null=C.c_void_p()
r["step_host zero-copy, no obs"]=timeit(lambda i: L.pbg_step_host(h,_ptr(acts_h[i%8]),null,_ptr(rew),_ptr(done)))
r["step_host zero-copy, actions only"]=timeit(lambda i: L.pbg_step_host(h,_ptr(acts_h[i%8]),null,null,null))
# device actions, host obs via zero copy: use pbg_step with mapped host pointers directly
def g(i):
    L.pbg_step(h,_ptr(acts_d[i%64]),_ptr(obs),_ptr(rew),_ptr(done),None,None,None,env._stream()); torch.cuda.current_stream().synchronize()
try:
    r["device actions, host obs/rew/done (UVA ptrs)"]=timeit(g)
except Exception as e: print("g failed",e)
env.set_zero_copy(False)
r["step_host staged"]=timeit(lambda i: env.step_host(acts_h[i%8],obs,rew,done))
for k,v in r.items(): print("%-50s %7.1f us/step  %.3e env-steps/s"%(k,v,E/v*1e6))
