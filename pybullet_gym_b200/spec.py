"""Frozen per-environment constants, transcribed from the reference.

Sources (all under /root/reference/pybulletgym/envs/):
  ids, max_episode_steps, reward_threshold     __init__.py:4-103
  robot ctor arguments, power, foot lists      roboschool/robot_locomotors.py:82-302, robot_pendula.py:5-51
  per-joint power_coef                         roboschool/robot_bases.py:89, robot_locomotors.py:103-127,152-164
  scene parameters                             roboschool/gym_locomotion_envs.py:18-20, gym_pendulum_envs.py:13-14,
                                               scene_bases.py:60-73, scene_stadium.py:13-14,33
  reward coefficients                          roboschool/gym_locomotion_envs.py:48-52,150-151
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

KIND_PENDULUM, KIND_PENDULUM_SWINGUP, KIND_HOPPER, KIND_WALKER2D, KIND_HALFCHEETAH, KIND_ANT, KIND_HUMANOID, \
    KIND_FLAGRUN, KIND_FLAGRUN_HARDER, KIND_DOUBLE_PENDULUM, KIND_REACHER, KIND_DOUBLE_PENDULUM_MJ, KIND_HOPPER_MJ, KIND_WALKER2D_MJ, KIND_ANT_MJ, KIND_HUMANOID_MJ, KIND_HALFCHEETAH_MJ = range(17)


@dataclass(frozen=True)
class SceneSpec:
    """World.clean_everything parameters (scene_bases.py:60-73) and the stadium ground (scene_stadium.py:33).

    The [EXT] block holds the pybullet/Bullet solver defaults the reference never overrides
    (SURVEY.md Appendix C4); they are restated from upstream Bullet and are part of the pin list.
    """
    gravity: float = 9.8
    timestep: float = 0.0165 / 4
    frame_skip: int = 4
    num_solver_iterations: int = 5
    contact_erp: float = 0.9            # setDefaultContactERP(0.9) -> solverInfo.m_erp2
    ground_friction: float = 0.8
    ground_restitution: float = 0.5     # combined with link restitution 0 -> 0
    stadium_halflen: float = 105 * 0.25
    stadium_halfwidth: float = 50 * 0.25
    # [EXT]
    erp: float = 0.2                    # solverInfo.m_erp, used by joint-limit rows
    linear_slop: float = 1e-5
    warmstarting_factor: float = 0.1
    max_coordinate_velocity: float = 100.0
    limit_max_impulse: float = 100.0
    # btMultiBodyJointLimitConstraint with solverInfo.m_splitImpulse (default on): a limit violated by more than 0.04 rad gets
    # its position error routed to m_rhsPenetration, which the multibody solver never applies -- the row only stops the
    # joint from moving further out.  The pretrained Ant policy prefers this variant (2250 vs 2020).
    limit_split_impulse: bool = True
    split_impulse_threshold: float = -0.04
    torsional_friction: bool = False     # spinning / rolling friction rows (C1.11, C6-9)
    ground_spinning_friction: float = 0.0    # the floor's own coefficients (changeDynamics(spinningFriction=, rollingFriction=))
    ground_rolling_friction: float = 0.0

    @property
    def dt(self) -> float:
        return self.timestep * self.frame_skip


@dataclass(frozen=True)
class CubeSpec:
    """HumanoidFlagrunHarder's `aggressive_cube` (robot_locomotors.py:236-243, gym_utils.py:9-16).

    assets/things/cube_small.urdf: box 0.05 m, lateral friction 1.0; `changeDynamics(mass=1.2)`.
    [EXT] changeDynamics(mass=...) recomputes the inertia from the collision box (m/12 (ly^2+lz^2)), which
    discards the URDF's inertia_scaling; the contact breaking threshold follows the compiler's relative rule
    (0.02 x angular-motion disc of the box)."""
    half_extent: float = 0.025
    mass: float = 1.2
    friction: float = 1.0
    pos0: Tuple[float, float, float] = (-1.5, 0.0, 0.05)
    breaking_threshold: float = 0.02

    @property
    def inertia(self) -> float:
        side = 2.0 * self.half_extent
        return self.mass / 12.0 * 2.0 * side * side

    @property
    def contact_threshold(self) -> float:
        return self.breaking_threshold * (3.0 ** 0.5) * self.half_extent


@dataclass(frozen=True)
class EnvSpec:
    id: str
    kind: int
    xml: str
    robot_name: str
    action_dim: int
    obs_dim: int
    power: float
    power_coef: Dict[str, float] = field(default_factory=dict)   # overrides of the default 100.0
    foot_list: Tuple[str, ...] = ()
    initial_z: Optional[float] = None        # None: latch torso z at the first calc_state after reset
    electricity_cost: float = -2.0
    stall_torque_cost: float = -0.1
    joints_at_limit_cost: float = -0.1
    scene: SceneSpec = SceneSpec()
    max_episode_steps: int = 1000
    reward_threshold: Optional[float] = None
    walk_target: Tuple[float, float] = (1e3, 0.0)
    entry_point: str = ""
    cube: Optional[CubeSpec] = None
    aux_links: Tuple[str, ...] = ()          # links whose COM the task layer reads besides robot_name (Reacher: fingertip, target)

    @property
    def noise_dim(self) -> int:
        """Reset draws per episode that the caller may inject (pbg_reset_with): one per actuated joint for the
        walkers (robot_locomotors.py:18-19), the hinge for the pendulum (robot_pendula.py:16), both hinges for the
        double pendulum (robot_pendula.py:66-68)."""
        if self.kind in (KIND_DOUBLE_PENDULUM, KIND_DOUBLE_PENDULUM_MJ):
            return 2
        if self.kind == KIND_REACHER:
            return 4            # target_x, target_y, joint0, joint1 (robot_manipulators.py:12-21)
        if self.kind in (KIND_HOPPER_MJ, KIND_WALKER2D_MJ, KIND_HALFCHEETAH_MJ):
            # add_ignored_joints=True puts the three root joints into ordered_joints, and WalkerBase.robot_specific_reset
            # draws for every ordered joint (mujoco/robot_bases.py:85-90, mujoco/robot_locomotors.py:16-19)
            return self.action_dim + 3
        return self.action_dim

    @property
    def noise_ranges(self) -> List[Tuple[float, float]]:
        """(low, high) of every reset draw, in draw order."""
        if self.kind == KIND_REACHER:
            return [(-0.27, 0.27), (-0.27, 0.27), (-3.14, 3.14), (-3.14, 3.14)]
        return [(-0.1, 0.1)] * self.noise_dim

    def torque_scale(self, ordered_joint_names: List[str]) -> List[float]:
        """tau_max per ordered joint = power * power_coef (robot_locomotors.py:29,189)."""
        if self.kind in (KIND_PENDULUM, KIND_PENDULUM_SWINGUP):
            # robot_pendula.py:25: only the slider is driven, 100 * clip(a)
            return [100.0 if n == "slider" else 0.0 for n in ordered_joint_names]
        if self.kind in (KIND_DOUBLE_PENDULUM, KIND_DOUBLE_PENDULUM_MJ):
            # robot_pendula.py:73 (roboschool and mujoco variants alike): 200 * clip(a) on the slider
            return [200.0 if n == "slider" else 0.0 for n in ordered_joint_names]
        if self.kind == KIND_REACHER:
            # robot_manipulators.py:23-26: 0.05 * clip(a) on the two arm hinges
            return [0.05 if n in ("joint0", "joint1") else 0.0 for n in ordered_joint_names]
        return [self.power * self.power_coef.get(n, 100.0) for n in ordered_joint_names]


_PENDULUM_SCENE = SceneSpec(timestep=0.0165, frame_skip=1)
_HUMANOID_POWER = {
    "abdomen_z": 100, "abdomen_y": 100, "abdomen_x": 100,
    "right_hip_x": 100, "right_hip_z": 100, "right_hip_y": 300, "right_knee": 200,
    "left_hip_x": 100, "left_hip_z": 100, "left_hip_y": 300, "left_knee": 200,
    "right_shoulder1": 75, "right_shoulder2": 75, "right_elbow": 75,
    "left_shoulder1": 75, "left_shoulder2": 75, "left_elbow": 75,
}
_RS = "pybulletgym.envs.roboschool."

SPECS: Dict[str, EnvSpec] = {s.id: s for s in [
    EnvSpec("InvertedPendulumPyBulletEnv-v0", KIND_PENDULUM, "inverted_pendulum.xml", "cart", 1, 5, 1.0,
            scene=_PENDULUM_SCENE, reward_threshold=950.0,
            entry_point=_RS + "gym_pendulum_envs:InvertedPendulumBulletEnv"),
    EnvSpec("InvertedPendulumSwingupPyBulletEnv-v0", KIND_PENDULUM_SWINGUP, "inverted_pendulum.xml", "cart", 1, 5,
            1.0, scene=_PENDULUM_SCENE, reward_threshold=800.0,
            entry_point=_RS + "gym_pendulum_envs:InvertedPendulumSwingupBulletEnv"),
    EnvSpec("InvertedDoublePendulumPyBulletEnv-v0", KIND_DOUBLE_PENDULUM, "inverted_double_pendulum.xml", "cart", 1, 9,
            1.0, scene=_PENDULUM_SCENE, reward_threshold=9100.0,
            entry_point=_RS + "gym_pendulum_envs:InvertedDoublePendulumBulletEnv"),
    # MuJoCo-style variant (pybulletgym/envs/mujoco/): same physics, gym-mujoco observation / reward layout
    EnvSpec("InvertedDoublePendulumMuJoCoEnv-v0", KIND_DOUBLE_PENDULUM_MJ, "inverted_double_pendulum.xml", "cart", 1, 11,
            1.0, scene=_PENDULUM_SCENE, reward_threshold=9100.0,
            entry_point="pybulletgym.envs.mujoco.gym_pendulum_envs:InvertedDoublePendulumMuJoCoEnv"),
    EnvSpec("HopperMuJoCoEnv-v0", KIND_HOPPER_MJ, "hopper.xml", "torso", 3, 11, 0.75, foot_list=("foot",),
            reward_threshold=2500.0, entry_point="pybulletgym.envs.mujoco.gym_locomotion_envs:HopperMuJoCoEnv"),
    EnvSpec("Walker2DMuJoCoEnv-v0", KIND_WALKER2D_MJ, "walker2d.xml", "torso", 6, 17, 0.40,
            power_coef={"foot_joint": 30.0, "foot_left_joint": 30.0}, foot_list=("foot", "foot_left"),
            reward_threshold=2500.0, entry_point="pybulletgym.envs.mujoco.gym_locomotion_envs:Walker2DMuJoCoEnv"),
    # HalfCheetah.robot_specific_reset (mujoco/robot_locomotors.py:207-210) calls changeDynamics(part.bodyIndex, part.bodyPartIndex,
    # lateralFriction=0.8, spinningFriction=0.1, rollingFriction=0.1, restitution=0.5) for every part.  part.bodyIndex is the
    # index into the robot's own body list (0), i.e. pybullet body 0 = the stadium floor, so what the call changes is the
    # FLOOR (its base link for the torso's bodyPartIndex -1; [EXT] link indices that do not exist on the floor are taken to be
    # ignored): lateral friction 0.8 (unchanged), restitution 0.5 (x link restitution 0 = 0) and spinning / rolling friction 0.1,
    # which Bullet combines with every link's lateral friction -> torsional friction rows on all ground contacts.
    EnvSpec("HalfCheetahMuJoCoEnv-v0", KIND_HALFCHEETAH_MJ, "half_cheetah.xml", "torso", 6, 17, 1.0,
            power_coef={"bthigh": 120.0, "bshin": 90.0, "bfoot": 60.0, "fthigh": 140.0, "fshin": 60.0, "ffoot": 30.0},
            foot_list=("ffoot", "fshin", "fthigh", "bfoot", "bshin", "bthigh"), reward_threshold=3000.0,
            scene=SceneSpec(torsional_friction=True, ground_spinning_friction=0.1, ground_rolling_friction=0.1),
            entry_point="pybulletgym.envs.mujoco.gym_locomotion_envs:HalfCheetahMuJoCoEnv"),
    EnvSpec("AntMuJoCoEnv-v0", KIND_ANT_MJ, "ant.xml", "torso", 8, 111, 2.5,
            foot_list=("front_left_foot", "front_right_foot", "left_back_foot", "right_back_foot"),
            reward_threshold=2500.0, entry_point="pybulletgym.envs.mujoco.gym_locomotion_envs:AntMuJoCoEnv"),
    EnvSpec("HumanoidMuJoCoEnv-v0", KIND_HUMANOID_MJ, "humanoid_symmetric.xml", "torso", 17, 376, 0.41,
            power_coef=_HUMANOID_POWER, foot_list=("right_foot", "left_foot"), initial_z=0.8,
            entry_point="pybulletgym.envs.mujoco.gym_locomotion_envs:HumanoidMuJoCoEnv"),
    EnvSpec("ReacherPyBulletEnv-v0", KIND_REACHER, "reacher.xml", "body0", 2, 9, 1.0,
            scene=SceneSpec(gravity=0.0, timestep=0.0165, frame_skip=1), max_episode_steps=150, reward_threshold=18.0,
            aux_links=("fingertip", "target"), entry_point=_RS + "gym_manipulator_envs:ReacherBulletEnv"),
    EnvSpec("HopperPyBulletEnv-v0", KIND_HOPPER, "hopper.xml", "torso", 3, 15, 0.75, foot_list=("foot",),
            reward_threshold=2500.0, entry_point=_RS + "gym_locomotion_envs:HopperBulletEnv"),
    EnvSpec("Walker2DPyBulletEnv-v0", KIND_WALKER2D, "walker2d.xml", "torso", 6, 22, 0.40,
            power_coef={"foot_joint": 30.0, "foot_left_joint": 30.0}, foot_list=("foot", "foot_left"),
            reward_threshold=2500.0, entry_point=_RS + "gym_locomotion_envs:Walker2DBulletEnv"),
    EnvSpec("HalfCheetahPyBulletEnv-v0", KIND_HALFCHEETAH, "half_cheetah.xml", "torso", 6, 26, 0.90,
            power_coef={"bthigh": 120.0, "bshin": 90.0, "bfoot": 60.0, "fthigh": 140.0, "fshin": 60.0, "ffoot": 30.0},
            foot_list=("ffoot", "fshin", "fthigh", "bfoot", "bshin", "bthigh"), reward_threshold=3000.0,
            entry_point=_RS + "gym_locomotion_envs:HalfCheetahBulletEnv"),
    EnvSpec("AntPyBulletEnv-v0", KIND_ANT, "ant.xml", "torso", 8, 28, 2.5,
            foot_list=("front_left_foot", "front_right_foot", "left_back_foot", "right_back_foot"),
            reward_threshold=2500.0, entry_point=_RS + "gym_locomotion_envs:AntBulletEnv"),
    EnvSpec("HumanoidPyBulletEnv-v0", KIND_HUMANOID, "humanoid_symmetric.xml", "torso", 17, 44, 0.41,
            power_coef=_HUMANOID_POWER, foot_list=("right_foot", "left_foot"), initial_z=0.8,
            electricity_cost=4.25 * -2.0, stall_torque_cost=4.25 * -0.1,
            entry_point=_RS + "gym_locomotion_envs:HumanoidBulletEnv"),
    EnvSpec("HumanoidFlagrunPyBulletEnv-v0", KIND_FLAGRUN, "humanoid_symmetric.xml", "torso", 17, 44, 0.41,
            power_coef=_HUMANOID_POWER, foot_list=("right_foot", "left_foot"), initial_z=0.8,
            electricity_cost=4.25 * -2.0, stall_torque_cost=4.25 * -0.1, reward_threshold=2000.0,
            entry_point=_RS + "gym_locomotion_envs:HumanoidFlagrunBulletEnv"),
    EnvSpec("HumanoidFlagrunHarderPyBulletEnv-v0", KIND_FLAGRUN_HARDER, "humanoid_symmetric.xml", "torso", 17, 44,
            0.41, power_coef=_HUMANOID_POWER, foot_list=("right_foot", "left_foot"), initial_z=0.8,
            electricity_cost=4.25 * -2.0, stall_torque_cost=4.25 * -0.1,   # quirk Q6: the `/= 4` is dead
            cube=CubeSpec(),
            entry_point=_RS + "gym_locomotion_envs:HumanoidFlagrunHarderBulletEnv"),
]}

# ids the reference registers (envs/__init__.py) that this backend does not implement
UNBACKED_IDS = (
    "PusherPyBulletEnv-v0",
    "ThrowerPyBulletEnv-v0", "StrikerPyBulletEnv-v0", "AtlasPyBulletEnv-v0",
    "InvertedPendulumMuJoCoEnv-v0",
)
