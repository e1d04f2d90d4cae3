"""MuJoCo-style env shells (pybulletgym/envs/mujoco/): same robots and physics, gym-mujoco observation layouts.

Backed so far: InvertedDoublePendulumMuJoCoEnv-v0 (mujoco/gym_pendulum_envs.py:40-75, mujoco/robot_pendula.py:55-91),
HopperMuJoCoEnv-v0 and Walker2DMuJoCoEnv-v0 (mujoco/gym_locomotion_envs.py:121-206, mujoco/robot_locomotors.py:86-165).
The reference's InvertedPendulumMuJoCoEnv raises on its first reset (mujoco/robot_pendula.py:16 reads an undefined
``self.swingup``).  HalfCheetahMuJoCoEnv-v0 (mujoco/gym_locomotion_envs.py:211-244, mujoco/robot_locomotors.py:169-211): its reset's
changeDynamics call lands on the stadium floor (spec.py), which gives every ground contact torsional friction rows.  AntMuJoCoEnv-v0 and
HumanoidMuJoCoEnv-v0 (mujoco/gym_locomotion_envs.py:246-260, mujoco/robot_locomotors.py:222-319) re-pack the WalkerBase state
as qpos[2:] ++ qvel ++ zero padding (111 / 376 entries).
"""
from __future__ import annotations

import numpy as np
import torch

from ..roboschool import robots as R
from ..roboschool.envs import BaseBulletEnv
from ..roboschool.envs import WalkerBaseBulletEnv as _RsWalkerEnv
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


class _MJWalker(R.MJCFBasedRobot):
    """Hopper / Walker2D of pybulletgym/envs/mujoco/robot_locomotors.py:86-165 (add_ignored_joints=True: the three root
    joints are part of ordered_joints with power_coef 0)."""

    def __init__(self, env_id):
        R.MJCFBasedRobot.__init__(self, env_id)
        b = self.bullet
        dof_of = {li: k for k, li in enumerate(b.dof_links())}
        self.ordered_joints = [R.Joint(self, b.links[i].joint_name, i, dof_of[i]) for i in b.dof_links()]
        self.jdict = {j.joint_name: j for j in self.ordered_joints}
        for j in self.ordered_joints:
            j.power_coef = 0.0 if j.joint_name.startswith("ignore") else float(self.spec.power_coef.get(j.joint_name, 100.0))
        self.power = self.spec.power
        self.foot_list = list(self.spec.foot_list)
        self.pos_after = 0


class WalkerBaseMuJoCoEnv(BaseBulletEnv):
    def __init__(self, robot, **kw):
        BaseBulletEnv.__init__(self, robot, **kw)
        self.walk_target_x, self.walk_target_y = 1e3, 0
        self.stateId = -1

    def create_single_player_scene(self, bullet_client):
        from ..roboschool.scenes import StadiumScene
        self.stadium_scene = StadiumScene(self.robot.spec.scene)
        return self.stadium_scene

    def _draw_reset_noise(self):
        # mujoco/robot_locomotors.py:16-19: one draw per ordered joint, root joints included
        return [self.np_random.uniform(low=-0.1, high=0.1) for _ in self.robot.ordered_joints]

    def _finish_reset(self, obs):
        self.robot._invalidate()
        self.robot.pos_after = float(self.robot.robot_body.get_pose()[0])
        self.stateId = 0
        return obs

    def _step(self, a):
        a = np.asarray(a, dtype=np.float32)
        assert np.isfinite(a).all()
        obs, rew, done, info = self._backend.step(torch.from_numpy(a.reshape(1, -1)))
        state = obs[0].cpu().numpy()
        self.robot._invalidate()
        self.robot.pos_after = float(self.robot.robot_body.get_pose()[0])
        terms = info["reward_terms"][0].cpu().numpy()
        self.rewards = [float(terms[0]), float(terms[1]), float(terms[2])]     # potential, alive_bonus, power_cost
        d = bool(done[0])
        self.HUD(state, a, d)
        self.reward += sum(self.rewards)
        return state, sum(self.rewards), d, {}


class HopperMuJoCoEnv(WalkerBaseMuJoCoEnv):
    def __init__(self, **kw):
        self.robot = _MJWalker("HopperMuJoCoEnv-v0")
        WalkerBaseMuJoCoEnv.__init__(self, self.robot, **kw)


class Walker2DMuJoCoEnv(WalkerBaseMuJoCoEnv):
    def __init__(self, **kw):
        self.robot = _MJWalker("Walker2DMuJoCoEnv-v0")
        WalkerBaseMuJoCoEnv.__init__(self, self.robot, **kw)


class HalfCheetahMuJoCoEnv(WalkerBaseMuJoCoEnv):
    """rewards = [potential, power_cost], never done (mujoco/gym_locomotion_envs.py:216-244)."""

    def __init__(self, **kw):
        self.robot = _MJWalker("HalfCheetahMuJoCoEnv-v0")
        WalkerBaseMuJoCoEnv.__init__(self, self.robot, **kw)

    def _step(self, a):
        state, rew, d, info = WalkerBaseMuJoCoEnv._step(self, a)
        self.reward -= sum(self.rewards)
        self.rewards = self.rewards[:2]
        self.reward += sum(self.rewards)
        return state, sum(self.rewards), d, info


class Ant(R.WalkerBase):
    def __init__(self):
        R.WalkerBase.__init__(self, "AntMuJoCoEnv-v0")

    def alive_bonus(self, z, pitch):
        return +1 if z > 0.26 else -1


class _FloatingMuJoCoEnv(_RsWalkerEnv):
    """AntMuJoCoEnv / HumanoidMuJoCoEnv (mujoco/gym_locomotion_envs.py:246-260): WalkerBaseMuJoCoEnv._step with the
    robot's MuJoCo-style observation; rewards = [alive, progress, joints_at_limit_cost, feet_collision_cost]."""

    def _finish_reset(self, obs):
        return _RsWalkerEnv._finish_reset(self, obs).astype(np.float64)

    def _step(self, a):
        state, rew, done, info = _RsWalkerEnv._step(self, a)
        self.reward -= sum(self.rewards)
        self.rewards = self.rewards[:4]
        self.reward += sum(self.rewards)
        return state.astype(np.float64), rew, done, info


class AntMuJoCoEnv(_FloatingMuJoCoEnv):
    def __init__(self, **kw):
        self.robot = Ant()
        _RsWalkerEnv.__init__(self, self.robot, **kw)


class HumanoidMuJoCoEnv(_FloatingMuJoCoEnv):
    def __init__(self, **kw):
        self.robot = R.Humanoid("HumanoidMuJoCoEnv-v0")
        _RsWalkerEnv.__init__(self, self.robot, **kw)
        self.electricity_cost = 4.25 * _RsWalkerEnv.electricity_cost        # set but unused (mujoco/gym_locomotion_envs.py:257-260)
        self.stall_torque_cost = 4.25 * _RsWalkerEnv.stall_torque_cost


ENTRY_POINTS = {"InvertedDoublePendulumMuJoCoEnv-v0": InvertedDoublePendulumMuJoCoEnv,
                "HopperMuJoCoEnv-v0": HopperMuJoCoEnv, "Walker2DMuJoCoEnv-v0": Walker2DMuJoCoEnv,
                "HalfCheetahMuJoCoEnv-v0": HalfCheetahMuJoCoEnv,
                "AntMuJoCoEnv-v0": AntMuJoCoEnv, "HumanoidMuJoCoEnv-v0": HumanoidMuJoCoEnv}
