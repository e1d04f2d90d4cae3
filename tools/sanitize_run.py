"""Small run of every kernel configuration for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from pybullet_gym_b200.vector_env import VectorEnv
from pybullet_gym_b200.spec import SPECS
for env_id in SPECS:
    env = VectorEnv(env_id, 61, seed=1)          # 61: partially filled last CTA / warp
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 25):
        a = torch.rand(61, env.action_dim, device="cuda", generator=g) * 2 - 1
        obs, r, d, info = env.step(a)
    s = env.get_state(); env.set_state(s); env.physics_step(a); env.observe(a); env.feet_contact()
    torch.cuda.synchronize()
    print(env_id, "ok", float(obs.abs().max()), env.stats()["episodes"])
