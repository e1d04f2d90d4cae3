"""A stand-in for the pybullet client, so that the reference's *own* Python task layer can be run.

pybullet / gym are not installable in this image, hence the reference package cannot
even be imported.  This module registers stub `pybullet`, `pybullet_data`, `pybullet_envs.bullet.
bullet_client`, `gym` and `pkg_resources` modules whose client object answers the ~30 calls the
hot path makes from the CPU oracle's physics (oracle/oracle.c) and the Bullet-shaped
link list of pybullet_gym_b200/mjcf/compiler.py.  With it, tools/gen_golden_task.py imports the
unmodified /root/reference/pybulletgym/envs/roboschool/*.py and records what the reference's
calc_state / _step / reset code computes -- the golden vectors that pin the oracle's task layer.

The *physics* behind the stub is our own restatement, so nothing here pins stepSimulation itself.
Test infrastructure only.
"""
from __future__ import annotations

import math
import os
import sys
import types

import numpy as np

_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, _ROOT)

from oracle.oracle import OracleEnv  # noqa: E402
from pybullet_gym_b200.mjcf import compiler as mj  # noqa: E402
from pybullet_gym_b200.spec import SPECS, EnvSpec  # noqa: E402

POSITION_CONTROL, VELOCITY_CONTROL, TORQUE_CONTROL = 2, 0, 1
JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_FIXED = 0, 1, 4


class RecordingRandom:
    """np_random replacement that logs every uniform() draw (the reset noise)."""

    def __init__(self, seed):
        self._rs = np.random.RandomState(seed)
        self.log = []

    def uniform(self, low=0.0, high=1.0, size=None):
        v = self._rs.uniform(low=low, high=high, size=size)
        self.log.append(("uniform", low, high, np.array(v, dtype=np.float64).reshape(-1).tolist()))
        return v

    def randint(self, *a, **k):
        v = self._rs.randint(*a, **k)
        self.log.append(("randint", v))
        return v


class FakeBulletClient:
    """One world.  Body 0..: whatever loadMJCF/loadSDF/loadURDF returned, in load order."""

    current_spec: EnvSpec = None       # set by the generator before the env is constructed
    max_contacts = 0

    def __init__(self, connection_mode=None):
        self._client = 0
        self.bodies = []               # dicts: kind = "robot" | "floor" | "misc"
        self.orc = None
        self.bm = None
        self.tau = None
        self.snapshots = []
        self.engine = {}
        self.calls = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    # ---- world parameters (scene_bases.py:70-73)
    def setGravity(self, x, y, z):
        self.engine["gravity"] = (x, y, z)

    def setDefaultContactERP(self, v):
        self.engine["contact_erp"] = v

    def setPhysicsEngineParameter(self, fixedTimeStep=None, numSolverIterations=None, numSubSteps=None, **kw):
        self.engine.update(fixedTimeStep=fixedTimeStep, numSolverIterations=numSolverIterations, numSubSteps=numSubSteps)

    def configureDebugVisualizer(self, *a, **k):
        pass

    def changeVisualShape(self, *a, **k):
        pass

    def changeDynamics(self, body, link, **kw):
        self.bodies[body].setdefault("dynamics", {}).update(kw)
        self.dynamics_calls = getattr(self, "dynamics_calls", [])
        self.dynamics_calls.append((int(body), int(link), self.bodies[body]["kind"], dict(kw)))

    def disconnect(self):
        pass

    # ---- loading
    def loadMJCF(self, path, flags=0):
        spec = FakeBulletClient.current_spec
        assert os.path.basename(path) == spec.xml, (path, spec.xml)
        self.bm = mj.parse_mjcf(path)      # parses the *reference's* file
        self.orc = OracleEnv(spec, bm=self.bm, max_contacts=FakeBulletClient.max_contacts)
        self.orc.reset(noise=np.zeros(spec.noise_dim))
        self.tau = np.zeros(self.orc.model.nd)
        self.dof_of_link = {li: k for k, li in enumerate(self.bm.dof_links())}
        ids = []
        if os.path.basename(path) in ("inverted_pendulum.xml", "inverted_double_pendulum.xml"):
            # world-level geoms ("floor" plane of the double pendulum, "rail"): separate static bodies, loaded
            # first
            if os.path.basename(path) == "inverted_double_pendulum.xml":
                self.bodies.append({"kind": "misc", "base": b"floor", "name": b"floor"})
                ids.append(len(self.bodies) - 1)
            self.bodies.append({"kind": "misc", "base": b"rail", "name": b"rail"})
            ids.append(len(self.bodies) - 1)
        if os.path.basename(path) == "reacher.xml":
            # six world-level geoms (ground, four sides, root) -> static bodies, loaded first
            for nm in (b"ground", b"sideS", b"sideE", b"sideN", b"sideW", b"root"):
                self.bodies.append({"kind": "misc", "base": nm, "name": nm})
                ids.append(len(self.bodies) - 1)
        # one pybullet multibody per top-level <body> (Reacher: arm, target); joint j of a body = link links[j]
        for a, b in self.bm.multibody_links:
            self.bodies.append({"kind": "robot", "base": self.bm.links[0].name.encode(), "name": self.bm.name.encode(),
                                "links": list(range(a, b))})
            ids.append(len(self.bodies) - 1)
        self.robot_id = ids[-1] if len(self.bm.multibody_links) == 1 else ids[-len(self.bm.multibody_links)]
        return tuple(ids)

    def loadSDF(self, path):
        assert os.path.basename(path) == "plane_stadium.sdf"
        self.bodies.append({"kind": "floor", "base": b"floor", "name": b"floor_obj"})
        self.floor_id = len(self.bodies) - 1
        return (self.floor_id,)

    def loadURDF(self, path, basePosition=None, **kw):
        self.bodies.append({"kind": "misc", "base": b"baseLink", "name": os.path.basename(path).encode(),
                            "pos": list(basePosition or [0, 0, 0])})
        # HumanoidFlagrunHarder's aggressive cube (gym_utils.py:9-16) is a simulated body of the oracle
        if os.path.basename(path) == "cube_small.urdf" and FakeBulletClient.current_spec.cube is not None:
            self.bodies[-1]["kind"] = "cube"
            self.cube_id = len(self.bodies) - 1
            self.orc.set_cube(pos=list(basePosition), quat=[0, 0, 0, 1], omega=[0, 0, 0], vel=[0, 0, 0])
        return len(self.bodies) - 1

    # ---- structure queries
    def getNumJoints(self, body):
        return len(self.bodies[body]["links"]) if self.bodies[body]["kind"] == "robot" else 0

    def _li(self, body, j):
        """oracle link index of joint / link j of a robot body"""
        return self.bodies[body]["links"][j]

    def getBodyInfo(self, body):
        b = self.bodies[body]
        return (b["base"], b["name"])

    def getJointInfo(self, body, j):
        l = self.bm.links[self._li(body, j)]
        jt = {mj.JT_REVOLUTE: JOINT_REVOLUTE, mj.JT_PRISMATIC: JOINT_PRISMATIC, mj.JT_FIXED: JOINT_FIXED}[l.jtype]
        return (j, l.joint_name.encode(), jt, -1, -1, 0, l.damping, 0.0, l.lower, l.upper, 0.0, 0.0, l.name.encode(),
                tuple(l.axis), (0, 0, 0), (0, 0, 0, 1), l.parent - 1)

    # ---- control
    def setJointMotorControl2(self, bodyIndex=None, jointIndex=None, controlMode=None, *args, **kw):
        self._count("setJointMotorControl2")
        if controlMode == TORQUE_CONTROL:
            self.tau[self.dof_of_link[self._li(bodyIndex, jointIndex)]] += kw.get("force", 0.0)

    def resetJointState(self, body, j, targetValue=0.0, targetVelocity=0.0):
        self.orc.set_joint(self.dof_of_link[self._li(body, j)], targetValue, targetVelocity)

    def resetBasePositionAndOrientation(self, body, pos, orn):
        if self.bodies[body]["kind"] == "misc":
            self.bodies[body]["pos"] = list(pos)
            return
        if self.bodies[body]["kind"] == "cube":
            self.orc.set_cube(pos=list(pos), quat=list(orn))       # velocities are left as they are
            return
        raise NotImplementedError

    def resetBaseVelocity(self, body, linearVelocity=(0, 0, 0), angularVelocity=(0, 0, 0)):
        if self.bodies[body]["kind"] == "cube":
            self.orc.set_cube(omega=list(angularVelocity), vel=list(linearVelocity))

    def stepSimulation(self):
        self._count("stepSimulation")
        self.orc.physics_step_torque(self.tau)
        self.tau[:] = 0.0           # forces are cleared after the substep loop

    # ---- state queries
    def getJointState(self, body, j):
        self._count("getJointState")
        li = self._li(body, j)
        if li not in self.dof_of_link:         # fixed joint (jointfix_*): pybullet reports zeros
            return (0.0, 0.0, (0,) * 6, 0.0)
        q, qd = self.orc.get_joint(self.dof_of_link[li])
        return (q, qd, (0,) * 6, 0.0)

    def _link(self, idx):
        return self.orc.link_state()[idx]

    def getBasePositionAndOrientation(self, body):
        self._count("getBasePositionAndOrientation")
        b = self.bodies[body]
        if b["kind"] == "floor":
            return (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0)
        if b["kind"] == "misc":
            return tuple(b.get("pos", [0, 0, 0])), (0.0, 0.0, 0.0, 1.0)
        if b["kind"] == "cube":
            p, q, _, _ = self.orc.get_cube()
            return tuple(p), tuple(q)
        if self.bm.floating:
            # the multibody's own (integrated, sign-continuous) base quaternion, as pybullet reports it
            st = self.orc.get_state()
            return tuple(st[0:3]), tuple(st[3:7])
        s = self._link(0)
        return tuple(s[0:3]), tuple(s[3:7])

    def getLinkState(self, body, link, computeLinkVelocity=0):
        self._count("getLinkState")
        s = self._link(self._li(body, link))
        base = (tuple(s[0:3]), tuple(s[3:7]), (0, 0, 0), (0, 0, 0, 1), tuple(s[0:3]), tuple(s[3:7]))
        if computeLinkVelocity:
            return base + (tuple(s[7:10]), (0.0, 0.0, 0.0))
        return base

    def getBaseVelocity(self, body):
        self._count("getBaseVelocity")
        s = self._link(0)
        ang = tuple(self.orc.get_state()[7:10]) if self.bm.floating else (0.0, 0.0, 0.0)    # base angular velocity
        return tuple(s[7:10]), ang

    def getDynamicsInfo(self, body, link):
        """(mass, lateral_friction, local_inertia_diagonal, local_inertial_pos, local_inertial_orn, restitution, rolling_friction,
        spinning_friction, contact_damping, contact_stiffness, body_type, collision_margin) of the Bullet-shaped link list"""
        l = self.bm.links[0 if link < 0 else self._li(body, link)]
        fr = l.geoms[0].friction if l.geoms else 0.5
        return (l.mass, fr, tuple(l.inertia), tuple(l.com), (0.0, 0.0, 0.0, 1.0), 0.0, 0.0, 0.0, -1.0, -1.0, 2, 0.0)

    def getPhysicsEngineParameters(self):
        return dict(self.engine)

    def getContactPoints(self, bodyA=-1, bodyB=-1, linkIndexA=-2, linkIndexB=-2):
        self._count("getContactPoints")
        la, lb, dist = self.orc.contacts()
        out = []
        if linkIndexA == -2:      # every contact point of the body (tools/pin_pybullet.py)
            for a, b_, d in zip(la, lb, dist):
                out.append((0, bodyA, self.floor_id if b_ == -1 else bodyA, int(a) - 1, int(b_) - 1 if b_ >= 0 else -1,
                            (0, 0, 0), (0, 0, 0), (0, 0, 1), d, 0.0))
            return tuple(out)
        for a, b, d in zip(la, lb, dist):
            if bodyA == self.robot_id and a - 1 == linkIndexA:
                if b == -2:      # the cube (oracle.c LINK_CUBE)
                    out.append((0, bodyA, getattr(self, "cube_id", -1), linkIndexA, -1, (0, 0, 0), (0, 0, 0), (0, 0, 1), d, 0.0))
                elif b < 0:
                    out.append((0, bodyA, self.floor_id, linkIndexA, -1, (0, 0, 0), (0, 0, 0), (0, 0, 1), d, 0.0))
                else:
                    out.append((0, bodyA, bodyA, linkIndexA, int(b) - 1, (0, 0, 0), (0, 0, 0), (0, 0, 1), d, 0.0))
            elif bodyA == self.robot_id and b - 1 == linkIndexA and b >= 0:
                out.append((0, bodyA, bodyA, linkIndexA, int(a) - 1, (0, 0, 0), (0, 0, 0), (0, 0, 1), d, 0.0))
        return tuple(out)

    # ---- snapshots (gym_locomotion_envs.py:23-25,35-36)
    def saveState(self):
        self.snapshots.append(self.orc.get_state().copy())
        return len(self.snapshots) - 1

    def restoreState(self, sid):
        self.orc.set_state(self.snapshots[sid])
        self.tau[:] = 0.0


def _euler_from_quaternion(q):
    x, y, z, w = q
    sarg = -2.0 * (x * z - w * y)
    roll = math.atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z)
    pitch = -0.5 * math.pi if sarg <= -1.0 else (0.5 * math.pi if sarg >= 1.0 else math.asin(sarg))
    yaw = math.atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)
    return (roll, pitch, yaw)


def install():
    """Register the stub modules.  Call before importing pybulletgym."""
    pb = types.ModuleType("pybullet")
    pb.IS_PBG_STUB = True          # so that nothing mistakes this for a real wheel (tests/test_pin_tool.py, bench.py)
    pb.POSITION_CONTROL, pb.VELOCITY_CONTROL, pb.TORQUE_CONTROL = POSITION_CONTROL, VELOCITY_CONTROL, TORQUE_CONTROL
    pb.GUI, pb.DIRECT = 1, 2
    pb.URDF_USE_SELF_COLLISION, pb.URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS = 8, 32
    for n in ("COV_ENABLE_GUI", "COV_ENABLE_RENDERING", "COV_ENABLE_PLANAR_REFLECTION"):
        setattr(pb, n, 0)
    pb.ER_BULLET_HARDWARE_OPENGL = 0
    pb.getEulerFromQuaternion = _euler_from_quaternion
    pb.error = RuntimeError
    sys.modules["pybullet"] = pb

    pdata = types.ModuleType("pybullet_data")
    pdata.getDataPath = lambda: "/nonexistent/pybullet_data"
    sys.modules["pybullet_data"] = pdata

    pe = types.ModuleType("pybullet_envs")
    peb = types.ModuleType("pybullet_envs.bullet")
    bc = types.ModuleType("pybullet_envs.bullet.bullet_client")
    bc.BulletClient = FakeBulletClient
    pe.bullet, peb.bullet_client = peb, bc
    sys.modules.update({"pybullet_envs": pe, "pybullet_envs.bullet": peb, "pybullet_envs.bullet.bullet_client": bc})

    gym = types.ModuleType("gym")
    gym.__version__ = "0.10.5"

    class Env:
        pass

    class Box:
        def __init__(self, low, high, **kw):
            self.low, self.high = np.asarray(low), np.asarray(high)
            self.shape = self.low.shape

        def sample(self):
            return np.random.uniform(self.low, self.high)

    gym.Env = Env
    spaces = types.ModuleType("gym.spaces")
    spaces.Box = Box
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (RecordingRandom(seed), seed)
    utils.seeding = seeding
    gym.spaces, gym.utils = spaces, utils
    envs = types.ModuleType("gym.envs")
    reg = types.ModuleType("gym.envs.registration")
    gym.registered = []
    reg.register = lambda **kw: gym.registered.append(kw)
    envs.registration = reg
    gym.envs = envs
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding,
                        "gym.envs": envs, "gym.envs.registration": reg})
    if "pkg_resources" not in sys.modules:
        try:
            import pkg_resources  # noqa: F401
        except Exception:
            pr = types.ModuleType("pkg_resources")
            pr.parse_version = lambda v: tuple(int(x) for x in v.split(".")[:3])
            sys.modules["pkg_resources"] = pr
    return gym
