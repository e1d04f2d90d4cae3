"""B200-native batched physics backend for the Roboschool-style environments of pybullet-gym.

    from pybullet_gym_b200 import VectorEnv, make
    venv = VectorEnv("AntPyBulletEnv-v0", num_envs=4096)      # torch CUDA tensors in / out
    env = make("AntPyBulletEnv-v0")                            # single-env Gym shell (numpy)

The hot path (apply_action -> stepSimulation -> calc_state -> reward) lives in libpbg_b200.so
(include/pbg.h); there is no CPU fallback.
"""
from .spec import SPECS, UNBACKED_IDS  # noqa: F401

__all__ = ["VectorEnv", "make", "SPECS", "UNBACKED_IDS"]


def __getattr__(name):
    if name == "VectorEnv":
        from .vector_env import VectorEnv
        return VectorEnv
    if name == "make":
        from .envs import make
        return make
    raise AttributeError(name)
