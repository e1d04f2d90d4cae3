"""Cost of the fused policy MLP (pbg_rollout_policy) relative to the physics step it is fused into.
usage: python tools/bench_policy.py [Ant|Humanoid|Hopper ...] [envs]
Prints: fused rollout rate at K steps per launch; torch MLP + pbg_step per step; the step kernel alone on the same
trajectory (CUDA events around the pbg_step launches only) -> share of the MLP inside the fused kernel."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybullet_gym_b200.vector_env import VectorEnv
name = sys.argv[1] if len(sys.argv) > 1 else "Ant"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
w = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "policy_%s.npz" % name)))
env = VectorEnv(name + "PyBulletEnv-v0", E, device="cuda:0", seed=0)
env.set_policy(*[w[k] for k in ("dense1_w", "dense1_b", "dense2_w", "dense2_b", "final_w", "final_b")], obs_shift=w.get("obs_shift"))
env.reset()
env.rollout_policy(200); torch.cuda.synchronize()          # into the policy's stationary regime (auto-reset on)
print("%s x %d, policy %d-%d-%d-%d" % (name, E, w["dense1_w"].shape[0], w["dense1_w"].shape[1], w["dense2_w"].shape[1], w["final_w"].shape[1]))
fused = {}
for tc in (True, False):
  env.set_policy_tensor_cores(tc)
  print("policy on %s" % ("tensor cores (mma.sync TF32)" if tc else "scalar FP32"))
  for K in (1, 10, 50):
      reps = max(1, 300 // K)
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(reps): env.rollout_policy(K)
      b.record(); torch.cuda.synchronize()
      fused[K] = a.elapsed_time(b) / (K * reps)
      print("fused policy K=%d: %.3e env-steps/s (%.4f ms/step)" % (K, E / (fused[K] * 1e-3), fused[K]))
wt = {k: torch.tensor(v, device="cuda") for k, v in w.items()}
shift = wt["obs_shift"] if "obs_shift" in wt else 0.0
ob = env.obs
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(300)]
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for t in range(300):
    x = torch.relu((ob + shift) @ wt["dense1_w"] + wt["dense1_b"]); x = torch.relu(x @ wt["dense2_w"] + wt["dense2_b"])
    act = (x @ wt["final_w"] + wt["final_b"]).contiguous()
    ev[t][0].record()
    ob, r, d = env.step_fast(act)
    ev[t][1].record()
b.record(); torch.cuda.synchronize()
t_total = a.elapsed_time(b) / 300
t_step = sum(x.elapsed_time(y) for x, y in ev) / 300
print("torch MLP + pbg_step: %.3e env-steps/s (%.4f ms/step), of which the step kernel %.4f ms" % (E / (t_total * 1e-3), t_total, t_step))
print("MLP inside the fused kernel: %.4f ms/step = %.1f %% of a fused step (K=50, scalar FP32); outside (3 cuBLAS launches + elementwise): %.4f ms = %.1f %%"
      % (fused[50] - t_step, 100 * (fused[50] - t_step) / fused[50], t_total - t_step, 100 * (t_total - t_step) / t_total))
