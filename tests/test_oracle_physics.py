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
    """A limited joint driven into its stop is stopped there.  With Bullet's split-impulse rule (spec.SceneSpec) an overshoot
    of more than 0.04 rad is not pushed back, only held: the joint stays where the sub-step that crossed the stop left it
    (up to qdot * h beyond) and its velocity is zero; without the rule the ERP term brings it back into a small band."""
    import dataclasses
    for split, band in ((True, 0.15), (False, 0.05)):
        spec = SPECS["HopperPyBulletEnv-v0"]
        spec = dataclasses.replace(spec, scene=dataclasses.replace(spec.scene, limit_split_impulse=split))
        env = oracle_lib.OracleEnv(spec)
        env.reset(noise=np.zeros(3))
        bm = env.model.bm
        foot = [l for l in bm.links if l.joint_name == "foot_joint"][0]
        for _ in range(20):
            env.physics_step([0.0, 0.0, 1.0])
        nd = env.model.nd
        s = env.get_state()
        assert foot.upper - 0.01 < s[5] < foot.upper + band, (split, s[5])
        assert abs(s[nd + 5]) < 0.5, (split, s[nd + 5])


@pytest.mark.parametrize("env_id", IDS)
def test_random_rollouts_stay_finite(env_id, oracle_lib):
    env = oracle_lib.OracleEnv(env_id, seed=3)
    n, ret, eps = env.rollout(1500, action_seed=5)
    assert n == 1500 and np.isfinite(ret)
    if SPECS[env_id].kind in (2, 3):     # hopper / walker fall quickly under random actions
        assert eps > 20


def test_airborne_centre_of_mass_is_ballistic(oracle_lib):
    """Internal joint torques cannot move the centre of mass: an airborne Ant keeps its horizontal COM fixed and its vertical
    COM on the discrete ballistic curve -- exactly when nothing moves internally, and to the integrator's O(h * qdot^2)
    momentum error under moderate random actions (link damping and joint limits off: a limit row that throws an ankle to
    25 rad/s inside one sub-step makes that error visible, in Bullet's semi-implicit scheme as well)."""
    # unit joint axes here: with the Ant's raw (-1,1,0) ankle axes Bullet's velocities and positions are not consistent with each
    # other (ImporterRules.normalize_joint_axes) and momentum is not a conserved quantity of that model
    rules = mj.ImporterRules(link_damping=0.0, normalize_joint_axes=True)
    mass = None

    def run(scale):
        nonlocal mass
        bm = mj.parse_mjcf("ant.xml", rules)
        for l in bm.links:
            l.lower, l.upper = 0.0, -1.0
        env = oracle_lib.OracleEnv("AntPyBulletEnv-v0", bm=bm)
        env.reset(noise=np.zeros(8))
        s = env.get_state(); s[2] = 50.0; env.set_state(s)
        mass = np.array([l.mass for l in bm.links])

        def com():
            ls = env.link_state()
            return (mass[:, None] * ls[:, 0:3]).sum(0) / mass.sum(), (mass[:, None] * ls[:, 7:10]).sum(0) / mass.sum()

        c0, _ = com()
        rng = np.random.default_rng(2)
        h, n, w = 0.0165 / 4, 0, np.zeros(4)
        for t in range(40):
            env.physics_step(scale * rng.uniform(-1, 1, 8))
            n += 4
            c, v = com()
            assert env.num_contacts() == 0
            w = np.maximum(w, [np.abs(c[:2] - c0[:2]).max(), np.abs(v[:2]).max(), abs(v[2] + 9.8 * h * n),    # v_n = -g h n
                               abs(c[2] - (c0[2] - 9.8 * h * h * n * (n + 1) / 2))])                       # z_n = z_0 - g h^2 n (n+1) / 2
        return w

    w0 = run(0.0)
    assert w0.max() < 1e-12, w0
    w1 = run(0.1)                     # joint speeds up to ~3 rad/s
    assert w1[0] < 1e-4 and w1[1] < 1e-3 and w1[2] < 1e-3 and w1[3] < 2e-4, w1


def test_static_contact_force_equals_weight(oracle_lib):
    """An Ant resting on its feet: the contact normal impulses average m g h per sub-step (5 PGS sweeps with a 0.1 warm start
    leave a small jitter, so single sub-steps scatter by ~10 % around it)."""
    import dataclasses
    spec = SPECS["AntPyBulletEnv-v0"]
    spec1 = dataclasses.replace(spec, scene=dataclasses.replace(spec.scene, frame_skip=1))
    env = oracle_lib.OracleEnv(spec1)
    env.reset(noise=np.zeros(8))
    s0 = env.get_state()
    s0[13 + 1], s0[13 + 3], s0[13 + 5], s0[13 + 7] = 1.0, -1.0, -1.0, 1.0
    env.set_state(s0)
    for _ in range(2400):
        env.physics_step(np.zeros(8))
    weight = sum(l.mass for l in env.model.bm.links) * 9.8 * (0.0165 / 4)
    tot = []
    for _ in range(400):
        env.physics_step(np.zeros(8))
        nl, nc, rows = env.rows()
        assert nc >= 3
        tot.append(rows[nl:nl + nc, 2].sum())
    assert abs(np.mean(tot) - weight) < 0.01 * weight, (np.mean(tot), weight)
    assert np.abs(env.get_state()[7:13]).max() < 0.05


def test_sliding_box_decelerates_at_mu_g(oracle_lib):
    """Coulomb friction on the cube of HumanoidFlagrunHarder (mu = 1.0 x 0.8): sliding along x on the floor it decelerates
    uniformly with an effective coefficient between mu and 1.3 mu, then sticks.  Not exactly mu: friction shifts the load to the
    two leading corners, and a friction row whose normal impulse has dropped to 0 is *skipped*, not reset
    (btMultiBodyConstraintSolver::solveSingleIteration, `if (totalImpulse > 0)`), so the unloaded corners keep the friction
    impulse they picked up in the first sweeps -- the sequential-impulse artifact is part of the restated algorithm."""
    env = oracle_lib.OracleEnv("HumanoidFlagrunHarderPyBulletEnv-v0", bm=mj.parse_mjcf("humanoid_symmetric.xml", mj.ImporterRules(link_damping=0.0)))
    env.reset(noise=np.zeros(17))
    x0, v0, mu_g = 30.0, 3.0, 0.8 * 9.8
    env.set_cube(pos=[x0, 5.0, 0.025], quat=[0, 0, 0, 1], omega=[0, 0, 0], vel=[v0, 0, 0])
    h = 0.0165 / 4
    vs = []
    for t in range(40):
        env.physics_step(np.zeros(17))
        p, q, w, v = env.get_cube()
        vs.append(v[0])
    vs = np.array(vs)
    sliding = vs > 0.2
    dec = -np.diff(vs[sliding]) / (4 * h)              # deceleration while sliding
    assert sliding.sum() > 10 and dec.min() > 0.99 * mu_g and dec.max() < 1.3 * mu_g, (dec, mu_g)
    assert np.abs(dec - dec.mean()).max() < 0.02 * dec.mean()        # uniform
    assert abs(vs[-1]) < 1e-3                          # stuck
    dist = p[0] - x0
    assert v0 * v0 / (2 * 1.3 * mu_g) - 0.03 < dist < v0 * v0 / (2 * mu_g) + 0.03
    assert abs(p[2] - 0.025) < 1e-3 and abs(p[1] - 5.0) < 5e-3


def _link_jacobians(bm, q, base_pos, base_quat):
    """Per link: world COM, linear / angular Jacobians w.r.t. u = [base omega, base v, qdot], world inertia (independent of the
    oracle's recursions, same construction as crba_mass_matrix)."""
    R, p = mj.link_world_frames(bm, q, base_pos, base_quat)
    dofs = bm.dof_links()
    nu = len(dofs) + 6
    out = []
    for i, l in enumerate(bm.links):
        c = p[i] + R[i] @ l.com
        Jv, Jw = np.zeros((3, nu)), np.zeros((3, nu))
        a = i
        while a >= 0:
            la = bm.links[a]
            if la.jtype == mj.JT_REVOLUTE:
                z = R[a] @ la.axis
                k = 6 + dofs.index(a)
                Jw[:, k] = z
                Jv[:, k] = np.cross(z, c - p[a])
            elif la.jtype == mj.JT_FREE:
                c0 = p[a] + R[a] @ la.com
                Jw[:, 0:3] = np.eye(3)
                for e in range(3):
                    Jv[:, e] = np.cross(np.eye(3)[e], c - c0)
                Jv[:, 3:6] = np.eye(3)
            a = la.parent
        out.append((c, Jv, Jw, R[i] @ np.diag(l.inertia) @ R[i].T))
    return out


@pytest.mark.parametrize("env_id", ["AntPyBulletEnv-v0", "HumanoidPyBulletEnv-v0"])
@pytest.mark.parametrize("scale,spin", [(0.0, (1.0, 0.5, 2.0)), (0.3, (0.0, 0.0, 0.0))])
def test_airborne_angular_momentum_converges_first_order(env_id, scale, spin, oracle_lib):
    """Gravity and internal torques exert no moment about the centre of mass: the angular momentum of an airborne Ant (tumbling
    with limp joints, or driven by random torques) is conserved by the continuous dynamics.  Bullet's semi-implicit Euler step
    conserves it to first order only, so the drift over a fixed time span must halve when the time step is halved -- which it
    would not if a Coriolis / gyroscopic term or the floating-base coupling of the articulated-body recursion were wrong."""
    import dataclasses
    rules = mj.ImporterRules(link_damping=0.0, normalize_joint_axes=True)
    drift = []
    for div in (1, 2, 4):
        sp = SPECS[env_id]
        bm = mj.parse_mjcf(sp.xml, rules)
        for l in bm.links:
            l.lower, l.upper = 0.0, -1.0
        nj = len(bm.dof_links())
        sp = dataclasses.replace(sp, scene=dataclasses.replace(sp.scene, timestep=sp.scene.timestep / div))
        env = oracle_lib.OracleEnv(sp, bm=bm)
        env.reset(noise=np.zeros(sp.noise_dim))
        s = env.get_state(); s[2] = 50.0; s[7:10] = spin; env.set_state(s)
        mass = np.array([l.mass for l in bm.links])

        def H():
            s = env.get_state()
            u = np.concatenate([s[7:13], s[13 + nj:13 + 2 * nj]])
            J = _link_jacobians(bm, s[13:13 + nj], s[0:3], s[3:7])
            C = sum(m * c for m, (c, _, _, _) in zip(mass, J)) / mass.sum()
            return sum(m * np.cross(c - C, Jv @ u) + Iw @ (Jw @ u) for m, (c, Jv, Jw, Iw) in zip(mass, J))

        h0 = H()
        rng = np.random.default_rng(2)
        w = 0.0
        for t in range(40):
            a = scale * (0.2 if 'Humanoid' in env_id else 1.0) * rng.uniform(-1, 1, sp.action_dim)   # limp enough not to hit itself
            for _ in range(div):
                env.physics_step(a)
            assert env.num_contacts() == 0
            w = max(w, np.abs(H() - h0).max())
        drift.append(w)
    assert drift[0] > 1e-3, drift                                        # a visible first-order error (Ant: ~0.3 % of |H| in the spin case)
    assert 0.45 < drift[1] / drift[0] < 0.55 and 0.45 < drift[2] / drift[1] < 0.55, drift


def test_ground_manifold_switch_agrees_with_candidates_at_rest(oracle_lib):
    """`ground_manifold` (oracle-only probe of SURVEY C5.2: one new floor point per collision pass and geom, up to four
    cached points, dropped beyond the breaking threshold) must reach the same resting configuration as the instantaneous
    end-sphere candidates the CUDA kernel tests: at rest the cached points ARE the end spheres' lowest points."""
    import ctypes as C
    rest = {}
    for mode in (0, 1):
        env = oracle_lib.OracleEnv("AntPyBulletEnv-v0", seed=1, ground_manifold=mode)
        env.reset()
        for _ in range(150):
            ob, _, _, _ = env.step(np.zeros(env.nact))
            assert oracle_lib.lib().orc_num_contacts(env._h) <= 4 * int(env.model.c.ng)
        rest[mode] = (env.get_state().copy(), oracle_lib.lib().orc_num_contacts(env._h))
    assert rest[0][1] == rest[1][1] > 0
    assert np.abs(rest[0][0][:7] - rest[1][0][:7]).max() < 1e-4
    # and a moving robot stays finite with manifolds on
    env = oracle_lib.OracleEnv("HalfCheetahPyBulletEnv-v0", seed=3, ground_manifold=1)
    n, ret, _ = env.rollout(600, action_seed=5)
    assert n == 600 and np.isfinite(ret)
