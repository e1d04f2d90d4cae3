"""Env registry: the ids of /root/reference/pybulletgym/envs/__init__.py:4-151.

``import pybullet_gym_b200.envs`` fills ``registry``; ``make(id)`` builds the single-env shell with
``TimeLimit(max_episode_steps)`` semantics as ``gym.make`` would.  If gym or gymnasium happens to be
importable the backed ids are also registered there under the same names.
"""
from __future__ import annotations

from ..spec import SPECS, UNBACKED_IDS

registry = {}


def register(id, entry_point, max_episode_steps=None, reward_threshold=None):
    registry[id] = dict(id=id, entry_point=entry_point, max_episode_steps=max_episode_steps,
                        reward_threshold=reward_threshold)


def get_list():
    return ["- " + i for i in registry if "Bullet" in i or "MuJoCo" in i]


def make(id, **kwargs):
    if id in UNBACKED_IDS:
        raise NotImplementedError(
            "%s is registered by the reference but not implemented on the B200 backend (see DESIGN.md, out of scope)" % id)
    if id not in registry:
        raise KeyError("No registered env with id: %s" % id)
    from .roboschool import envs as _envs
    from .mujoco import envs as _mj_envs
    from .time_limit import TimeLimit
    cls = _envs.ENTRY_POINTS[id] if id in _envs.ENTRY_POINTS else _mj_envs.ENTRY_POINTS[id]
    env = cls(**kwargs)
    return TimeLimit(env, max_episode_steps=registry[id]["max_episode_steps"])


for _s in SPECS.values():
    register(_s.id, "pybullet_gym_b200.envs.%s.envs:" % ("mujoco" if ".mujoco." in _s.entry_point else "roboschool")
             + _s.entry_point.split(":")[1], _s.max_episode_steps,
             _s.reward_threshold)


def _register_with_gym():
    for modname in ("gymnasium", "gym"):
        try:
            mod = __import__(modname + ".envs.registration", fromlist=["register"])
        except Exception:
            continue
        for r in registry.values():
            try:
                mod.register(id=r["id"], entry_point=r["entry_point"], max_episode_steps=r["max_episode_steps"],
                             reward_threshold=r["reward_threshold"])
            except Exception:
                pass


_register_with_gym()
