"""MuJoCo-style env shells (pybulletgym/envs/mujoco/): same robots and physics, gym-mujoco observation layouts.

Backed so far: InvertedDoublePendulumMuJoCoEnv-v0 (mujoco/gym_pendulum_envs.py:40-75, mujoco/robot_pendula.py:55-91).
The reference's InvertedPendulumMuJoCoEnv raises on its first reset (mujoco/robot_pendula.py:16 reads an undefined
``self.swingup``); the five MuJoCo-style walkers are SURVEY.md section 8f N2 work.
"""
from __future__ import annotations

import numpy as np
import torch

from ..roboschool import robots as R
from ..roboschool.envs import BaseBulletEnv
from ..roboschool.scenes import SingleRobotEmptyScene


class InvertedDoublePendulum(R.MJCFBasedRobot):
    def __init__(self):
        R.MJCFBasedRobot.__init__(self, "InvertedDoublePendulumMuJoCoEnv-v0")
        self.pole2 = self.parts["pole2"]
        self.slider = self.jdict["slider"]
        self.j1 = self.jdict["hinge"]
        self.j2 = self.jdict["hinge2"]


class InvertedDoublePendulumMuJoCoEnv(BaseBulletEnv):
    def __init__(self, **kw):
        self.robot = InvertedDoublePendulum()
        BaseBulletEnv.__init__(self, self.robot, **kw)
        self.stateId = -1

    def create_single_player_scene(self, bullet_client):
        return SingleRobotEmptyScene(self.robot.spec.scene)

    def _draw_reset_noise(self):
        return list(self.np_random.uniform(low=-.1, high=.1, size=[2]))     # mujoco/robot_pendula.py:66

    def _finish_reset(self, obs):
        self.robot._invalidate()
        self.stateId = 0
        return obs.astype(np.float64)

    def _step(self, a):
        a = np.asarray(a, dtype=np.float32)
        assert np.isfinite(a).all()
        obs, rew, done, info = self._backend.step(torch.from_numpy(a.reshape(1, -1)))
        state = obs[0].cpu().numpy().astype(np.float64)
        self.robot._invalidate()
        terms = info["reward_terms"][0].cpu().numpy()
        self.rewards = [float(terms[0]), float(terms[1]), float(terms[2])]     # alive_bonus, -dist_penalty, -vel_penalty
        d = bool(done[0])
        self.HUD(state, a, d)
        return state, sum(self.rewards), d, {}


ENTRY_POINTS = {"InvertedDoublePendulumMuJoCoEnv-v0": InvertedDoublePendulumMuJoCoEnv}
