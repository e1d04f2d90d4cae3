"""Physics invariants of the CPU oracle (the reference ships no golden vectors for stepSimulation, so
the oracle's dynamics are checked against first principles): inverse mass matrix symmetry / positive
definiteness, agreement with an independent composite-rigid-body mass matrix, energy conservation of
the undamped contact-free pendulum, free fall, static contact force = m g, joint-limit hold."""
import numpy as np
import pytest

from pybullet_gym_b200.mjcf import compiler as mj
from pybullet_gym_b200.spec import SPECS

IDS = ["InvertedPendulumPyBulletEnv-v0", "HopperPyBulletEnv-v0", "Walker2DPyBulletEnv-v0", "HalfCheetahPyBulletEnv-v0",
       "AntPyBulletEnv-v0", "HumanoidPyBulletEnv-v0"]


def crba_mass_matrix(bm, q, base_pos=None, base_quat=None):
    """Independent joint-space inertia: M = sum_links J_i^T diag(m, I) J_i with numerical-free Jacobians."""
    R, p = mj.link_world_frames(bm, q, base_pos, base_quat)
    dofs = bm.dof_links()
    nu = len(dofs) + (6 if bm.floating else 0)
    off = 6 if bm.floating else 0
    M = np.zeros((nu, nu))
    for i, l in enumerate(bm.links):
        if l.mass == 0 and not l.inertia.any():
            continue
        c = p[i] + R[i] @ l.com
        Jv, Jw = np.zeros((3, nu)), np.zeros((3, nu))
        a = i
        while a >= 0:
            la = bm.links[a]
            if la.jtype == mj.JT_REVOLUTE:
                z = R[a] @ la.axis
                k = off + dofs.index(a)
                Jw[:, k] = z
                Jv[:, k] = np.cross(z, c - p[a])
            elif la.jtype == mj.JT_PRISMATIC:
                Jv[:, off + dofs.index(a)] = R[a] @ la.axis
            elif la.jtype == mj.JT_FREE:
                c0 = p[a] + R[a] @ la.com
                Jw[:, 0:3] = np.eye(3)
                for e in range(3):
                    Jv[:, e] = np.cross(np.eye(3)[e], c - c0)
                Jv[:, 3:6] = np.eye(3)
            a = la.parent
        Iw = R[i] @ np.diag(l.inertia) @ R[i].T
        M += l.mass * Jv.T @ Jv + Jw.T @ Iw @ Jw
    return M


@pytest.mark.parametrize("env_id", IDS)
def test_inverse_mass_matrix_matches_independent_crba(env_id, oracle_lib):
    env = oracle_lib.OracleEnv(env_id)
    rng = np.random.default_rng(1)
    env.reset(noise=rng.uniform(-0.1, 0.1, env.nact))
    s = env.get_state()
    nd = env.model.nd
    if env.model.bm.floating:
        qt = rng.normal(size=4); qt /= np.linalg.norm(qt)
        s[3:7] = qt
        s[13:13 + nd] = rng.uniform(-0.5, 0.5, nd)
        env.set_state(s)
        M = crba_mass_matrix(env.model.bm, s[13:13 + nd], s[0:3], s[3:7])
    else:
        s[:nd] = rng.uniform(-0.5, 0.5, nd)
        env.set_state(s)
        M = crba_mass_matrix(env.model.bm, s[:nd])
    Minv = env.mass_matrix_inv()
    assert np.abs(Minv - Minv.T).max() < 1e-10
    assert np.linalg.eigvalsh(0.5 * (Minv + Minv.T)).min() > 0
    assert np.abs(Minv @ M - np.eye(M.shape[0])).max() < 1e-8


def test_pendulum_energy_drift_without_damping(oracle_lib):
    """Contact-free, undamped, unactuated cart-pole: semi-implicit Euler keeps the energy error O(h)."""
    rules = mj.ImporterRules(link_damping=0.0)
    bm = mj.parse_mjcf("inverted_pendulum.xml", rules)
    env = oracle_lib.OracleEnv("InvertedPendulumSwingupPyBulletEnv-v0", bm=bm)
    env.reset(noise=[0.05])          # swing-up starts hanging down: bounded oscillation
    e0 = env.energy()
    es = []
    for _ in range(600):
        env.physics_step([0.0])
        es.append(env.energy())
    assert np.abs(np.array(es) - e0).max() < 0.02 * (abs(e0) + 1.0)


def test_free_fall_and_static_contact(oracle_lib):
    env = oracle_lib.OracleEnv("AntPyBulletEnv-v0")
    env.reset(noise=np.zeros(8))
    s0 = env.get_state()
    # put the ankles inside their ranges (at q = 0 they violate them, quirk Q8, and the limit rows push)
    s0[13 + 1], s0[13 + 3], s0[13 + 5], s0[13 + 7] = 1.0, -1.0, -1.0, 1.0
    s0[2] = 1.5
    env.set_state(s0)
    env.physics_step(np.zeros(8))
    s1 = env.get_state()
    # in the air the base accelerates at -g (minus the tiny link damping): v_z ~ -g * 0.0165
    assert abs(s1[12] + 9.8 * 0.0165) < 2e-3
    assert abs((s1[2] - s0[2]) + 0.5 * 9.8 * 0.0165 ** 2 * (1 + 1 / 4)) < 1e-3
    assert np.abs(s1[13:21] - s0[13:21]).max() < 1e-3       # nothing else moves
    for _ in range(500):
        env.physics_step(np.zeros(8))
    s = env.get_state()
    assert np.abs(s[7:13]).max() < 0.05, "ant should come to rest on its feet"
    assert 0.25 < s[2] < 0.75 and env.num_contacts() >= 3


def test_joint_limit_hold(oracle_lib):
    """A limited joint driven into its stop stays within a small ERP band of the limit."""
    env = oracle_lib.OracleEnv("HopperPyBulletEnv-v0")
    env.reset(noise=np.zeros(3))
    bm = env.model.bm
    foot = [l for l in bm.links if l.joint_name == "foot_joint"][0]
    for _ in range(30):
        env.physics_step([0.0, 0.0, 1.0])
    nd = env.model.nd
    q = env.get_state()[:nd]
    assert q[5] < foot.upper + 0.05 and q[5] > foot.lower - 0.05


@pytest.mark.parametrize("env_id", IDS)
def test_random_rollouts_stay_finite(env_id, oracle_lib):
    env = oracle_lib.OracleEnv(env_id, seed=3)
    n, ret, eps = env.rollout(1500, action_seed=5)
    assert n == 1500 and np.isfinite(ret)
    if SPECS[env_id].kind in (2, 3):     # hopper / walker fall quickly under random actions
        assert eps > 20
