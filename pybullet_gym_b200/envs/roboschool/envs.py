"""Single-environment Gym shells with the reference's class names, attributes and old 4-tuple API
(roboschool/env_bases.py:9-121, gym_locomotion_envs.py:8-173, gym_pendulum_envs.py:7-49).

Each shell owns a one-env ``VectorEnv``; ``reset()/step()`` return numpy like the reference does.
For throughput use ``pybullet_gym_b200.VectorEnv`` directly -- this shell exists so that code written
against ``gym.make("AntPyBulletEnv-v0")`` keeps working.
"""
from __future__ import annotations

import numpy as np
import torch

from ...spec import SPECS
from ...vector_env import VectorEnv
from . import robots as R
from .scenes import SingleRobotEmptyScene, StadiumScene


class BaseBulletEnv:
    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 60}

    def __init__(self, robot, render=False, device=None):
        self.scene = None
        self.physicsClientId = -1
        self.ownsPhysicsClient = 0
        self.isRender = render
        self.robot = robot
        self._device = device
        self._backend = None
        self._seed()
        self.action_space = robot.action_space
        self.observation_space = robot.observation_space
        self.reward = 0
        self.rewards = []
        self.potential = 0.0

    def configure(self, args):
        self.robot.args = args

    def _seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        self.robot.np_random = self.np_random   # same generator for env and robot (env_bases.py:41-44)
        # The Flagrun flag positions and cube attacks are drawn on the device from a counter RNG; the reference draws them
        # from this np_random.  Key the device RNG by the same seed (OS entropy when unseeded, like RandomState(None)), so
        # that differently seeded envs / worker processes see different targets and the same seed repeats them.
        import os
        self._backend_seed = int(seed) & 0xFFFFFFFFFFFFFFFF if seed is not None else int.from_bytes(os.urandom(8), "little")
        if getattr(self, "_backend", None) is not None:
            self._backend.seed(self._backend_seed)
        return [seed]

    def _ensure_backend(self):
        if self._backend is None:
            # the "physics client" is created lazily on the first reset (env_bases.py:46-56)
            self._backend = VectorEnv(self.robot.spec.id, 1, device=self._device, seed=self._backend_seed, auto_reset=False)
            self._backend.enable_contact_export(True)       # BodyPart.contact_list()
            self.physicsClientId = 0
            self.ownsPhysicsClient = True
            self.robot._env = self
        if self.scene is None:
            self.scene = self.create_single_player_scene(None)
        self.scene.episode_restart(None)
        self.robot.scene = self.scene

    def _draw_reset_noise(self):
        raise NotImplementedError

    def _reset(self):
        self._ensure_backend()
        self.frame, self.done, self.reward = 0, 0, 0
        noise = torch.tensor(np.asarray(self._draw_reset_noise(), dtype=np.float32)).reshape(1, -1)
        obs = self._backend.reset(joint_noise=noise)
        return self._finish_reset(obs[0].cpu().numpy())

    def _finish_reset(self, obs):
        return obs

    def _render(self, mode="human", close=False):
        if mode == "human":
            self.isRender = True
        return np.array([])     # rendering is out of scope of the batched backend

    def _close(self):
        if self._backend is not None:
            self._backend.close()
            self._backend = None
        self.physicsClientId = -1

    def HUD(self, state, a, done):
        pass

    def step(self, *args, **kwargs):
        return self._step(*args, **kwargs)

    def close(self):
        return self._close()

    def render(self, mode="human", close=False):
        return self._render(mode, close)

    def reset(self):
        return self._reset()

    def seed(self, seed=None):
        return self._seed(seed)


class WalkerBaseBulletEnv(BaseBulletEnv):
    electricity_cost = -2.0
    stall_torque_cost = -0.1
    foot_collision_cost = -1.0
    foot_ground_object_names = set(["floor"])
    joints_at_limit_cost = -0.1

    def __init__(self, robot, render=False, device=None):
        BaseBulletEnv.__init__(self, robot, render, device)
        self.camera_x = 0
        self.walk_target_x, self.walk_target_y = robot.spec.walk_target
        self.stateId = -1

    def create_single_player_scene(self, bullet_client):
        self.stadium_scene = StadiumScene(self.robot.spec.scene)
        return self.stadium_scene

    def _draw_reset_noise(self):
        # robot_locomotors.py:18-19: one U(-0.1,0.1) draw per ordered joint, in order
        return [self.np_random.uniform(low=-0.1, high=0.1) for _ in self.robot.ordered_joints]

    def _finish_reset(self, obs):
        r = self.robot
        r.feet_contact = np.zeros(len(r.foot_list), dtype=np.float32)
        r.initial_z = None
        first = "floor" not in r.parts
        self._mirror(obs)
        if first:
            # quirk Q1: the floor joins robot.parts after the first reset (gym_locomotion_envs.py:30-31)
            r.parts["floor"] = R.BodyPart(r, "floor", None)
            self.parts, self.jdict, self.ordered_joints, self.robot_body = r.parts, r.jdict, r.ordered_joints, r.robot_body
        # (body id, link index) of the ground objects, as the reference collects them for the feet test
        # (gym_locomotion_envs.py:33-36): the stadium floor is body 0, base link
        self.ground_ids = set([(R.XmlBasedRobot.FLOOR_BODY, -1)])
        self.stateId = 0
        return obs

    def _mirror(self, obs):
        """Host views after a reset / step.  Targets, flag timeout, frame counters and the potential (fp64, with
        FlagrunHarder's crawl bookkeeping, robot_locomotors.py:280-302) are the device's own values."""
        r = self.robot
        view = {k: v.cpu().numpy() for k, v in self._backend.task_view().items()}
        if hasattr(r, "_mirror_task"):
            r._mirror_task(view)
        self.walk_target_x, self.walk_target_y = r.walk_target_x, r.walk_target_y
        r._update_views(obs)
        self.potential = float(view["potential"][0])

    def _step(self, a):
        a = np.asarray(a, dtype=np.float32)
        assert np.isfinite(a).all()
        obs, rew, done, info = self._backend.step(torch.from_numpy(a.reshape(1, -1)))
        state = obs[0].cpu().numpy()
        terms = info["reward_terms"][0].cpu().numpy()
        r = self.robot
        self._mirror(state)
        r.feet_contact = self._backend.feet_contact()[0].cpu().numpy().astype(np.float32)
        self.rewards = [float(t) for t in terms]
        self.HUD(state, a, bool(done[0]))
        self.reward += sum(self.rewards)
        return state, float(rew[0]), bool(done[0]), {}


class HopperBulletEnv(WalkerBaseBulletEnv):
    def __init__(self, **kw):
        self.robot = R.Hopper()
        WalkerBaseBulletEnv.__init__(self, self.robot, **kw)


class Walker2DBulletEnv(WalkerBaseBulletEnv):
    def __init__(self, **kw):
        self.robot = R.Walker2D()
        WalkerBaseBulletEnv.__init__(self, self.robot, **kw)


class HalfCheetahBulletEnv(WalkerBaseBulletEnv):
    def __init__(self, **kw):
        self.robot = R.HalfCheetah()
        WalkerBaseBulletEnv.__init__(self, self.robot, **kw)


class AntBulletEnv(WalkerBaseBulletEnv):
    def __init__(self, **kw):
        self.robot = R.Ant()
        WalkerBaseBulletEnv.__init__(self, self.robot, **kw)


class HumanoidBulletEnv(WalkerBaseBulletEnv):
    def __init__(self, robot=None, **kw):
        self.robot = robot if robot is not None else R.Humanoid()
        WalkerBaseBulletEnv.__init__(self, self.robot, **kw)
        self.electricity_cost = 4.25 * WalkerBaseBulletEnv.electricity_cost
        self.stall_torque_cost = 4.25 * WalkerBaseBulletEnv.stall_torque_cost


class HumanoidFlagrunBulletEnv(HumanoidBulletEnv):
    random_yaw = True          # dead in the reference: the robot is built with random_yaw=False (quirk Q7)

    def __init__(self, **kw):
        self.robot = R.HumanoidFlagrun()
        HumanoidBulletEnv.__init__(self, self.robot, **kw)

    def create_single_player_scene(self, bullet_client):
        s = HumanoidBulletEnv.create_single_player_scene(self, bullet_client)
        s.zero_at_running_strip_start_line = False
        return s


class HumanoidFlagrunHarderBulletEnv(HumanoidBulletEnv):
    random_lean = True         # dead in the reference as well (quirk Q7)

    def __init__(self, **kw):
        self.robot = R.HumanoidFlagrunHarder()
        HumanoidBulletEnv.__init__(self, self.robot, **kw)     # the reference's `electricity_cost /= 4` is dead (quirk Q6)

    def create_single_player_scene(self, bullet_client):
        s = HumanoidBulletEnv.create_single_player_scene(self, bullet_client)
        s.zero_at_running_strip_start_line = False
        return s


class InvertedPendulumBulletEnv(BaseBulletEnv):
    def __init__(self, **kw):
        self.robot = R.InvertedPendulum()
        BaseBulletEnv.__init__(self, self.robot, **kw)
        self.stateId = -1

    def create_single_player_scene(self, bullet_client):
        return SingleRobotEmptyScene(self.robot.spec.scene)

    def _draw_reset_noise(self):
        return [self.np_random.uniform(low=-.1, high=.1)]       # robot_pendula.py:16

    def _finish_reset(self, obs):
        self.robot._invalidate()
        self.robot.theta = self.robot.j1.get_position()
        self.stateId = 0
        return obs.astype(np.float64)

    def _step(self, a):
        a = np.asarray(a, dtype=np.float32)
        assert np.isfinite(a).all()
        obs, rew, done, info = self._backend.step(torch.from_numpy(a.reshape(1, -1)))
        state = obs[0].cpu().numpy().astype(np.float64)
        self.robot._invalidate()
        self.robot.theta = self.robot.j1.get_position()
        self.rewards = [float(rew[0])]
        d = bool(done[0])
        self.HUD(state, a, d)
        return state, sum(self.rewards), d, {}


class InvertedPendulumSwingupBulletEnv(InvertedPendulumBulletEnv):
    def __init__(self, **kw):
        self.robot = R.InvertedPendulumSwingup()
        BaseBulletEnv.__init__(self, self.robot, **kw)
        self.stateId = -1


class InvertedDoublePendulumBulletEnv(BaseBulletEnv):
    """gym_pendulum_envs.py:50-86."""

    def __init__(self, **kw):
        self.robot = R.InvertedDoublePendulum()
        BaseBulletEnv.__init__(self, self.robot, **kw)
        self.stateId = -1

    def create_single_player_scene(self, bullet_client):
        return SingleRobotEmptyScene(self.robot.spec.scene)

    def _draw_reset_noise(self):
        return list(self.np_random.uniform(low=-.1, high=.1, size=[2]))     # robot_pendula.py:66

    def _finish_reset(self, obs):
        self.robot._invalidate()
        self.robot.pos_x, _, self.robot.pos_y = self.robot.pole2.pose().xyz()
        self.stateId = 0
        return obs.astype(np.float64)

    def _step(self, a):
        a = np.asarray(a, dtype=np.float32)
        assert np.isfinite(a).all()
        obs, rew, done, info = self._backend.step(torch.from_numpy(a.reshape(1, -1)))
        state = obs[0].cpu().numpy().astype(np.float64)
        self.robot._invalidate()
        self.robot.pos_x, _, self.robot.pos_y = self.robot.pole2.pose().xyz()
        terms = info["reward_terms"][0].cpu().numpy()
        self.rewards = [float(terms[0]), float(terms[1]), float(terms[2])]     # alive_bonus, -dist_penalty, -vel_penalty
        d = bool(done[0])
        self.HUD(state, a, d)
        return state, sum(self.rewards), d, {}


class ReacherBulletEnv(BaseBulletEnv):
    """gym_manipulator_envs.py:7-38."""

    def __init__(self, **kw):
        self.robot = R.Reacher()
        BaseBulletEnv.__init__(self, self.robot, **kw)

    def create_single_player_scene(self, bullet_client):
        return SingleRobotEmptyScene(self.robot.spec.scene)

    def _draw_reset_noise(self):
        # robot_manipulators.py:12-21: target_x, target_y, joint0, joint1 -- in this order
        r, lim = self.np_random, self.robot.TARG_LIMIT
        return [r.uniform(low=-lim, high=lim), r.uniform(low=-lim, high=lim),
                r.uniform(low=-3.14, high=3.14), r.uniform(low=-3.14, high=3.14)]

    def _finish_reset(self, obs):
        self.robot._update_views()
        self.potential = self.robot.calc_potential()
        return obs.astype(np.float64)

    def _step(self, a):
        a = np.asarray(a, dtype=np.float32)
        assert np.isfinite(a).all()
        obs, rew, done, info = self._backend.step(torch.from_numpy(a.reshape(1, -1)))
        state = obs[0].cpu().numpy().astype(np.float64)
        self.robot._update_views()
        self.potential = self.robot.calc_potential()
        terms = info["reward_terms"][0].cpu().numpy()
        self.rewards = [float(terms[0]), float(terms[1]), float(terms[2])]     # progress, electricity_cost, stuck_joint_cost
        self.HUD(state, a, False)
        return state, sum(self.rewards), False, {}


ENTRY_POINTS = {
    "InvertedPendulumPyBulletEnv-v0": InvertedPendulumBulletEnv,
    "ReacherPyBulletEnv-v0": ReacherBulletEnv,
    "InvertedDoublePendulumPyBulletEnv-v0": InvertedDoublePendulumBulletEnv,
    "InvertedPendulumSwingupPyBulletEnv-v0": InvertedPendulumSwingupBulletEnv,
    "HopperPyBulletEnv-v0": HopperBulletEnv,
    "Walker2DPyBulletEnv-v0": Walker2DBulletEnv,
    "HalfCheetahPyBulletEnv-v0": HalfCheetahBulletEnv,
    "AntPyBulletEnv-v0": AntBulletEnv,
    "HumanoidPyBulletEnv-v0": HumanoidBulletEnv,
    "HumanoidFlagrunPyBulletEnv-v0": HumanoidFlagrunBulletEnv,
    "HumanoidFlagrunHarderPyBulletEnv-v0": HumanoidFlagrunHarderBulletEnv,
}
