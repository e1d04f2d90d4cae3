"""HumanoidFlagrunHarder's aggressive cube on the CUDA path vs the CPU oracle.  Run with -m gpu.

The cube (rs/robot_locomotors.py:236-266, gym_utils.py:9-16) is a second free body of the world: box-vs-plane
and box-vs-capsule/sphere contacts, thrown at the robot every 30 frames after frame 100 while it stands.  The
oracle's attack arithmetic is pinned against the reference's Python by tests/golden/task_HumanoidFlagrunHarderHeld.json;
here the kernel is compared with the oracle on the same seeded inputs (same Philox streams).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ENV_ID = "HumanoidFlagrunHarderPyBulletEnv-v0"


def _pair(oracle_lib, n, seed=1):
    from pybullet_gym_b200 import _lib
    from pybullet_gym_b200.spec import SPECS
    from pybullet_gym_b200.vector_env import VectorEnv
    env = VectorEnv(ENV_ID, n, device="cuda:0", seed=seed, auto_reset=False)
    budget = _lib.solver_budget(SPECS[ENV_ID].kind)
    orcs = [oracle_lib.OracleEnv(ENV_ID, seed=seed, env_index=i, **budget) for i in range(n)]
    return env, orcs


def test_state_layout_and_reset(oracle_lib):
    env, orcs = _pair(oracle_lib, 4)
    env.reset(floor_in_parts=True)
    st = env.get_state().cpu().numpy()
    assert st.shape[1] == 13 + 34 + 13 == orcs[0].model.state_size
    ost = np.stack([(o.reset(floor_in_parts=True), o.get_state())[1] for o in orcs])
    assert np.abs(st - ost).max() < 1e-6
    assert np.allclose(st[:, -13:-10], [-1.5, 0.0, 0.05]) and np.allclose(st[:, -7], 1.0)     # cube rest pose, identity quaternion


def test_cube_settles_on_the_floor(oracle_lib):
    """Box-vs-plane: dropped from 2.5 cm, the cube comes to rest on its face on both paths."""
    env, orcs = _pair(oracle_lib, 4)
    env.reset(floor_in_parts=True)
    for o in orcs:
        o.reset(floor_in_parts=True)
    a = np.zeros((4, 17), np.float32)
    for t in range(40):
        env.physics_step(torch.from_numpy(a))
        for o in orcs:
            o.physics_step(a[0].astype(np.float64))
    g = env.get_state().cpu().numpy()[:, -13:]
    o = np.stack([x.get_state() for x in orcs])[:, -13:]
    assert np.abs(g[:, 2] - 0.025).max() < 2e-4 and np.abs(o[:, 2] - 0.025).max() < 2e-4      # resting on a face
    assert np.abs(g[:, :3] - o[:, :3]).max() < 1e-4
    assert np.abs(g[:, 7:]).max() < 1e-2                                                      # at rest


def test_thrown_cube_hits_the_robot_like_the_oracle(oracle_lib):
    """Cube thrown at different body heights from different directions: free flight agrees to fp32 round-off,
    the momentum it hands to the robot agrees within the contact tier's tolerance."""
    n = 32
    env, orcs = _pair(oracle_lib, n)
    rng = np.random.default_rng(4)
    noise = rng.uniform(-0.1, 0.1, (n, 17)).astype(np.float32)
    env.reset(joint_noise=torch.from_numpy(noise), floor_in_parts=True)
    for i, o in enumerate(orcs):
        o.reset(noise=noise[i].astype(np.float64), floor_in_parts=True)
    st = np.stack([o.get_state() for o in orcs])
    ang = rng.uniform(-np.pi, np.pi, n)
    height = rng.uniform(-0.35, 0.25, n)              # relative to the torso COM: pelvis .. head
    speed = rng.uniform(20, 30, n)
    target = st[:, 0:3] + np.stack([np.zeros(n), np.zeros(n), height], 1)
    start = target + 1.0 * np.stack([np.cos(ang), np.sin(ang), np.zeros(n)], 1)
    st[:, -13:-10] = start
    st[:, -3:] = (target - start) * speed[:, None]
    st32 = st.astype(np.float32)
    env.set_state(torch.from_numpy(st32))
    for i, o in enumerate(orcs):
        o.set_state(st32[i].astype(np.float64))
    a = np.zeros((n, 17), np.float32)
    hit_g = np.zeros(n, bool); hit_o = np.zeros(n, bool)
    flight_err = 0.0
    for t in range(6):
        env.physics_step(torch.from_numpy(a))
        for o in orcs:
            o.physics_step(a[0].astype(np.float64))
        g = env.get_state().cpu().numpy(); o_ = np.stack([x.get_state() for x in orcs])
        sg, so = np.linalg.norm(g[:, -3:], axis=1), np.linalg.norm(o_[:, -3:], axis=1)
        hit_g |= sg < 0.8 * speed; hit_o |= so < 0.8 * speed
        free = ~(hit_g | hit_o)
        if free.any():
            flight_err = max(flight_err, (np.abs(g[free, -13:] - o_[free, -13:]) / (1 + np.abs(o_[free, -13:]))).max())
    assert flight_err < 2e-5, flight_err
    assert hit_o.sum() >= n // 3, hit_o.sum()                  # the scenario does produce hits
    assert (hit_g == hit_o).mean() >= 0.9, (hit_g.sum(), hit_o.sum())
    both = hit_g & hit_o
    # momentum handed over: cube velocity after the hit, robot base velocity after the hit
    dv_cube = np.abs(g[both, -3:] - o_[both, -3:]).max(axis=1)
    dv_base = np.abs(g[both, 10:13] - o_[both, 10:13]).max(axis=1)
    assert np.median(dv_cube) < 0.05 and np.median(dv_base) < 5e-3, (np.median(dv_cube), np.median(dv_base))
    assert np.isfinite(g).all()


def test_attack_launch_matches_oracle(oracle_lib):
    """Held humanoid (state put back to the reset pose before every step, as in the held golden fixture): the attack
    at frame 120 draws the same Philox numbers on both paths and launches the cube from the same place."""
    n = 8
    env, orcs = _pair(oracle_lib, n, seed=3)
    env.reset(floor_in_parts=True)
    for o in orcs:
        o.reset(floor_in_parts=True)
    s0 = np.stack([o.get_state() for o in orcs])
    a = np.zeros((n, 17), np.float32)
    launched = None
    for t in range(124):
        sg = env.get_state().cpu().numpy()
        so = np.stack([o.get_state() for o in orcs])
        for s in (sg, so):
            s[:, :47] = s0[:, :47]
            s[:, 0] += 0.01 * t
            s[:, 10] = 0.6
        env.set_state(torch.from_numpy(sg.astype(np.float32)))
        for i, o in enumerate(orcs):
            o.set_state(so[i])
        obs, rew, done, _ = env.step(torch.from_numpy(a))
        res = [o.step(a[0].astype(np.float64)) for o in orcs]
        assert not done.any() and not any(r[2] for r in res)
        if t == 120:
            launched = (env.get_state().cpu().numpy()[:, -13:], np.stack([o.get_state() for o in orcs])[:, -13:])
    g, o_ = launched
    assert (np.linalg.norm(o_[:, -3:], axis=1) > 15).all()            # the cube is in flight after frame 120
    assert np.abs(g[:, :3] - o_[:, :3]).max() < 1e-4, np.abs(g[:, :3] - o_[:, :3]).max()
    assert np.abs(g[:, -3:] - o_[:, -3:]).max() < 1e-3, np.abs(g[:, -3:] - o_[:, -3:]).max()
    assert np.abs(obs.cpu().numpy() - np.stack([r[0] for r in res])).max() < 1e-3
