"""Robot shells: the attribute surface of the reference's robot_bases.py / robot_locomotors.py /
robot_pendula.py (parts, jdict, ordered_joints, feet, feet_contact, body_xyz, body_rpy, joint_speeds,
joints_at_limit, walk_target_*, initial_z) on top of the batched backend.

Observations, rewards and physics come from the CUDA library; the inspection helpers below
(BodyPart.pose(), Joint.current_position(), ...) are computed on the host from the canonical state
with the MJCF compiler's forward kinematics, for debugging and parity checks only.
"""
from __future__ import annotations

import math

import numpy as np

from ...mjcf import compiler as mj
from ...spec import SPECS, EnvSpec
from ..spaces import Box


def _euler(q):
    x, y, z, w = q
    sarg = -2.0 * (x * z - w * y)
    return (math.atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z),
            -0.5 * math.pi if sarg <= -1 else (0.5 * math.pi if sarg >= 1 else math.asin(sarg)),
            math.atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z))


class PoseHelper:
    def __init__(self, body_part):
        self.body_part = body_part

    def xyz(self):
        return self.body_part.current_position()

    def rpy(self):
        return _euler(self.body_part.current_orientation())

    def orientation(self):
        return self.body_part.current_orientation()


class BodyPart:
    """One Bullet link of the robot (or the floor), as robot_bases.BodyPart."""

    def __init__(self, robot, name, link_index):
        self.robot, self.name, self.link_index = robot, name, link_index
        self.bp_pose = PoseHelper(self)
        self.bodyPartIndex = -1 if link_index is None else link_index - 1

    def get_pose(self):
        if self.link_index is None:      # the floor
            return np.array([0, 0, 0, 0, 0, 0, 1.0])
        R, p = self.robot._frames()
        l = self.robot.bullet.links[self.link_index]
        return np.concatenate([p[self.link_index] + R[self.link_index] @ l.com, mj.mat_to_q(R[self.link_index])])

    def current_position(self):
        return self.get_pose()[:3]

    def current_orientation(self):
        return self.get_pose()[3:]

    get_position = current_position
    get_orientation = current_orientation

    def pose(self):
        return self.bp_pose

    def speed(self):
        """COM linear velocity of the link (robot_bases.py:243-248)."""
        if self.link_index is None:
            return np.zeros(3)
        return self.robot._com_velocities()[self.link_index].copy()

    def contact_list(self):
        """Contact points of this link after the last step, shaped like pybullet's getContactPoints tuples as far as the
        reference reads them (robot_bases.py:280-281; gym_locomotion_envs.py:73 uses fields [2] = bodyB and [4] = linkB):
        (contactFlag, bodyA, bodyB, linkA, linkB, posA, posB, normal, distance, normalForce)."""
        return self.robot._contacts_of(self.name)


class Joint:
    def __init__(self, robot, joint_name, link_index, dof):
        self.robot, self.joint_name, self.link_index, self.dof = robot, joint_name, link_index, dof
        l = robot.bullet.links[link_index]
        self.jointType = 0 if l.jtype == mj.JT_REVOLUTE else 1
        self.lowerLimit, self.upperLimit = l.lower, l.upper
        self.jointHasLimits = self.lowerLimit < self.upperLimit
        self.jointMaxVelocity = 0.0
        self.power_coef = 100.0
        self.power_coeff = 0          # sic: dead attribute of the reference (quirk Q4)

    def get_state(self):
        s = self.robot._state()
        o = 13 if self.robot.bullet.floating else 0
        nd = self.robot.nd
        return float(s[o + self.dof]), float(s[o + nd + self.dof])

    current_position = get_state

    def current_relative_position(self):
        pos, vel = self.get_state()
        if self.jointHasLimits:
            pos_mid = 0.5 * (self.lowerLimit + self.upperLimit)
            pos = 2 * (pos - pos_mid) / (self.upperLimit - self.lowerLimit)
        vel *= 0.1 if self.jointType == 0 else 0.5
        return (pos, vel)

    def get_position(self):
        return self.get_state()[0]

    def get_velocity(self):
        return self.get_state()[1]


class XmlBasedRobot:
    self_collision = True

    def __init__(self, env_id):
        self.spec: EnvSpec = SPECS[env_id]
        self.robot_name = self.spec.robot_name
        self.model_xml = self.spec.xml
        self.bullet = mj.parse_mjcf(self.spec.xml)
        self.nd = len(self.bullet.dof_links())
        self.action_space = Box(-np.ones(self.spec.action_dim), np.ones(self.spec.action_dim))
        self.observation_space = Box(-np.inf * np.ones(self.spec.obs_dim), np.inf * np.ones(self.spec.obs_dim))
        dof_of = {li: k for k, li in enumerate(self.bullet.dof_links())}
        self.parts = {}
        for i, l in enumerate(self.bullet.links):
            if i > 0 or self.bullet.floating:
                self.parts[l.name] = BodyPart(self, l.name, i)
        self.jdict, self.ordered_joints = {}, []
        for i in self.bullet.ordered_joints():
            j = Joint(self, self.bullet.links[i].joint_name, i, dof_of[i])
            self.jdict[j.joint_name] = j
            self.ordered_joints.append(j)
        for n, c in self.spec.power_coef.items():
            if n in self.jdict:
                self.jdict[n].power_coef = float(c)
        self.robot_body = self.parts.get(self.robot_name)
        self.np_random = np.random.RandomState()
        self.scene = None
        self._env = None
        self._cache = None

    # host-side views of the device state
    def _state(self):
        if self._cache is None:
            self._cache = {"state": self._env._backend.get_state()[0].double().cpu().numpy()}
        return self._cache["state"]

    def _frames(self):
        s = self._state()
        if "frames" not in self._cache:
            if self.bullet.floating:
                self._cache["frames"] = mj.link_world_frames(self.bullet, s[13:13 + self.nd], s[0:3], s[3:7])
            else:
                self._cache["frames"] = mj.link_world_frames(self.bullet, s[:self.nd])
        return self._cache["frames"]

    def _com_velocities(self):
        s = self._state()
        if "vel" not in self._cache:
            nd = self.nd
            if self.bullet.floating:
                self._cache["vel"] = mj.link_com_velocities(self.bullet, s[13:13 + nd], s[13 + nd:13 + 2 * nd], s[0:3], s[3:7],
                                                            s[7:10], s[10:13])
            else:
                self._cache["vel"] = mj.link_com_velocities(self.bullet, s[:nd], s[nd:2 * nd])
        return self._cache["vel"]

    # body ids as the reference's scene would hand them out: the stadium floor is loaded first (scene_stadium.py:28), then the robot
    FLOOR_BODY, ROBOT_BODY, CUBE_BODY = 0, 1, 2

    def _contacts_of(self, link_name):
        be = self._env._backend if self._env is not None else None
        if be is None or be._cand is None:
            return []
        if "cand" not in self._cache:
            self._state()
            self._cache["cand"] = be.contact_candidates()[0].cpu().numpy()
            self._cache["slots"] = be.tables.contact_slots()
        idx = {l.name: i - 1 for i, l in enumerate(self.bullet.links)}
        out = []
        for (a, b), d in zip(self._cache["slots"], self._cache["cand"]):
            if not np.isfinite(d) or link_name not in (a, b):
                continue
            other = b if a == link_name else a
            if other == "floor":
                body_b, link_b = self.FLOOR_BODY, -1
            elif other == "cube":
                body_b, link_b = self.CUBE_BODY, -1
            else:
                body_b, link_b = self.ROBOT_BODY, idx.get(other, -1)
            me = (self.FLOOR_BODY, -1) if link_name == "floor" else ((self.CUBE_BODY, -1) if link_name == "cube"
                                                                      else (self.ROBOT_BODY, idx.get(link_name, -1)))
            out.append((0, me[0], body_b, me[1], link_b, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), float(d), 0.0))
        return out

    def _invalidate(self):
        self._cache = None


class MJCFBasedRobot(XmlBasedRobot):
    pass


class WalkerBase(MJCFBasedRobot):
    foot_list = []

    def __init__(self, env_id):
        MJCFBasedRobot.__init__(self, env_id)
        self.power = self.spec.power
        self.foot_list = list(self.spec.foot_list)
        self.camera_x = 0
        self.start_pos_x, self.start_pos_y, self.start_pos_z = 0, 0, 0
        self.walk_target_x, self.walk_target_y = self.spec.walk_target
        self.body_xyz = [0, 0, 0]
        self.feet = [self.parts[f] for f in self.foot_list]
        self.feet_contact = np.zeros(len(self.foot_list), dtype=np.float32)
        self.initial_z = None
        self.joint_speeds = np.zeros(self.spec.action_dim, dtype=np.float32)
        self.joints_at_limit = 0
        self.body_rpy = (0.0, 0.0, 0.0)
        self.walk_target_theta, self.walk_target_dist = 0.0, 0.0

    def _update_views(self, obs):
        """Mirror the side outputs of WalkerBase.calc_state (robot_locomotors.py:31-64) from the state."""
        self._invalidate()
        j = np.array([jt.current_relative_position() for jt in self.ordered_joints], dtype=np.float32).flatten()
        self.joint_speeds = j[1::2]
        self.joints_at_limit = int(np.count_nonzero(np.abs(j[0::2]) > 0.99))
        xyz = np.array([p.pose().xyz() for p in self.parts.values()])
        body_pose = self.robot_body.pose()
        self.body_xyz = (xyz[:, 0].mean(), xyz[:, 1].mean(), body_pose.xyz()[2])
        self.body_rpy = body_pose.rpy()
        if self.initial_z is None:
            self.initial_z = self.body_xyz[2] if self.spec.initial_z is None else self.spec.initial_z
        self.walk_target_theta = math.atan2(self.walk_target_y - self.body_xyz[1], self.walk_target_x - self.body_xyz[0])
        self.walk_target_dist = math.hypot(self.walk_target_y - self.body_xyz[1], self.walk_target_x - self.body_xyz[0])

    def calc_potential(self):
        return -self.walk_target_dist / self.scene.dt


class Hopper(WalkerBase):
    def __init__(self):
        WalkerBase.__init__(self, "HopperPyBulletEnv-v0")

    def alive_bonus(self, z, pitch):
        return +1 if z > 0.8 and abs(pitch) < 1.0 else -1


class Walker2D(WalkerBase):
    def __init__(self):
        WalkerBase.__init__(self, "Walker2DPyBulletEnv-v0")

    alive_bonus = Hopper.alive_bonus


class HalfCheetah(WalkerBase):
    def __init__(self):
        WalkerBase.__init__(self, "HalfCheetahPyBulletEnv-v0")

    def alive_bonus(self, z, pitch):
        fc = self.feet_contact
        return +1 if abs(pitch) < 1.0 and not fc[1] and not fc[2] and not fc[4] and not fc[5] else -1


class Ant(WalkerBase):
    def __init__(self):
        WalkerBase.__init__(self, "AntPyBulletEnv-v0")

    def alive_bonus(self, z, pitch):
        return +1 if z > 0.26 else -1


class Humanoid(WalkerBase):
    def __init__(self, env_id="HumanoidPyBulletEnv-v0"):
        WalkerBase.__init__(self, env_id)
        self.motor_names = list(self.spec.power_coef.keys())
        self.motor_power = [self.spec.power_coef[n] for n in self.motor_names]
        self.motors = [self.jdict[n] for n in self.motor_names]

    def alive_bonus(self, z, pitch):
        return +2 if z > 0.78 else -1


class HumanoidFlagrun(Humanoid):
    def __init__(self, env_id="HumanoidFlagrunPyBulletEnv-v0"):
        Humanoid.__init__(self, env_id)
        self.flag = None
        self.flag_timeout = 0

    def _mirror_task(self, view):
        """The flag lives on the device (flag_reposition, robot_locomotors.py:204-218): mirror its position / timeout."""
        self.walk_target_x, self.walk_target_y = float(view["walk_target_x"][0]), float(view["walk_target_y"][0])
        self.flag_timeout = float(view["flag_timeout"][0])


class HumanoidFlagrunHarder(HumanoidFlagrun):
    def __init__(self):
        HumanoidFlagrun.__init__(self, "HumanoidFlagrunHarderPyBulletEnv-v0")
        self.aggressive_cube = None
        self.frame = 0
        self.on_ground_frame_counter = 0

    def _mirror_task(self, view):
        HumanoidFlagrun._mirror_task(self, view)
        self.frame = int(view["frame"][0])
        self.on_ground_frame_counter = int(view["on_ground_frame_counter"][0])


class InvertedPendulum(MJCFBasedRobot):
    swingup = False

    def __init__(self, env_id="InvertedPendulumPyBulletEnv-v0"):
        MJCFBasedRobot.__init__(self, env_id)
        self.pole = self.parts["pole"]
        self.slider = self.jdict["slider"]
        self.j1 = self.jdict["hinge"]
        self.theta = 0.0


class InvertedDoublePendulum(MJCFBasedRobot):
    """robot_pendula.py:57-87."""

    def __init__(self):
        MJCFBasedRobot.__init__(self, "InvertedDoublePendulumPyBulletEnv-v0")
        self.pole2 = self.parts["pole2"]
        self.slider = self.jdict["slider"]
        self.j1 = self.jdict["hinge"]
        self.j2 = self.jdict["hinge2"]
        self.pos_x, self.pos_y = 0.0, 0.0


class Reacher(MJCFBasedRobot):
    """robot_manipulators.py:5-50."""
    TARG_LIMIT = 0.27

    def __init__(self):
        MJCFBasedRobot.__init__(self, "ReacherPyBulletEnv-v0")
        self.fingertip = self.parts["fingertip"]
        self.target = self.parts["target"]
        self.central_joint = self.jdict["joint0"]
        self.elbow_joint = self.jdict["joint1"]
        self.theta_dot = self.gamma = self.gamma_dot = 0.0
        self.to_target_vec = np.zeros(3)

    def _update_views(self):
        self._invalidate()
        _, self.theta_dot = self.central_joint.current_relative_position()
        self.gamma, self.gamma_dot = self.elbow_joint.current_relative_position()
        self.to_target_vec = np.array(self.fingertip.pose().xyz()) - np.array(self.target.pose().xyz())

    def calc_potential(self):
        return -100 * np.linalg.norm(self.to_target_vec)


class InvertedPendulumSwingup(InvertedPendulum):
    swingup = True

    def __init__(self):
        InvertedPendulum.__init__(self, "InvertedPendulumSwingupPyBulletEnv-v0")
