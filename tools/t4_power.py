"""One-off, higher-powered T4: 16384 CUDA episodes vs 8192 oracle episodes (two-sample KS on length and return)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.stats import ks_2samp
from oracle import oracle as O
from pybullet_gym_b200 import _lib
from pybullet_gym_b200.spec import SPECS
from pybullet_gym_b200.vector_env import VectorEnv
for env_id, cap in (("HumanoidFlagrunHarderPyBulletEnv-v0", 300), ("HumanoidFlagrunPyBulletEnv-v0", 300), ("AntPyBulletEnv-v0", 300)):
    n, m = 16384, 8192
    env = VectorEnv(env_id, n, device="cuda:0", seed=11, auto_reset=False)
    env.reset(floor_in_parts=True)
    gen = torch.Generator(device="cuda").manual_seed(5)
    ret = torch.zeros(n, device="cuda"); length = torch.zeros(n, device="cuda"); alive = torch.ones(n, device="cuda")
    for t in range(cap):
        a = torch.rand(n, env.action_dim, device="cuda", generator=gen) * 2 - 1
        obs, rew, done = env.step_fast(a)
        ret += alive * rew; length += alive
        alive = alive * (1 - done.float())
    g_ret, g_len = ret.cpu().numpy(), length.cpu().numpy()
    o_ret, o_len = O.random_policy_episodes(env_id, m, cap, seed=77, **_lib.solver_budget(SPECS[env_id].kind))
    print(env_id, "len %.2f vs %.2f  ret %.2f vs %.2f (std %.1f)  p_len %.4f p_ret %.4f" % (g_len.mean(), o_len.mean(), g_ret.mean(), o_ret.mean(), o_ret.std(),
          ks_2samp(g_len, o_len).pvalue, ks_2samp(g_ret, o_ret).pvalue))
    qs = [.01, .05, .25, .5, .75, .95, .99]
    print("   gpu quantiles", np.quantile(g_ret, qs).round(1)); print("   orc quantiles", np.quantile(o_ret, qs).round(1))
