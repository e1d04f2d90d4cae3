import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from pybullet_gym_b200.vector_env import VectorEnv
w = dict(np.load("tests/golden/policy_Ant.npz"))
E=4096
env = VectorEnv("AntPyBulletEnv-v0", E, device="cuda:0", seed=0)
env.set_policy(*[w[k] for k in ("dense1_w", "dense1_b", "dense2_w", "dense2_b", "final_w", "final_b")])
env.reset()
env.rollout_policy(50); torch.cuda.synchronize()
for K in (1, 10, 100):
    reps = max(1, 300 // K)
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): env.rollout_policy(K)
    b.record(); torch.cuda.synchronize()
    ms=a.elapsed_time(b)
    print("fused policy K=%d: %.3e env-steps/s (%.4f ms/step)" % (K, E*K*reps/(ms*1e-3), ms/(K*reps)))
# torch MLP + step
wt = {k: torch.tensor(v, device="cuda") for k, v in w.items()}
ob = env.obs
for rep in range(2):
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for t in range(300):
        x = torch.relu(ob @ wt["dense1_w"] + wt["dense1_b"]); x = torch.relu(x @ wt["dense2_w"] + wt["dense2_b"])
        ob, r, d = env.step_fast((x @ wt["final_w"] + wt["final_b"]).contiguous())
    b.record(); torch.cuda.synchronize()
print("torch MLP + pbg_step: %.3e env-steps/s (%.4f ms/step)" % (E*300/(a.elapsed_time(b)*1e-3), a.elapsed_time(b)/300))
