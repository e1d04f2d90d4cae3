"""CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.  Run with -m gpu.

Tiers (BASELINE.json north_star):
  T1  calc_state / reward / done / feet flags given identical state  1e-5 abs (progress: 2e-5 * (1 + |x|), fp32 FK);
      done and feet flags exact except where the oracle itself sits within round-off of a threshold
  T2  single sub-step / single env step from identical state  median bounds + EVERY sample bounded by K x the oracle's
      own response to a 1e-7 perturbation of that same state
  T3  contact-free pendulum trajectory over a full episode . 2e-4 after 100 steps, 1e-3 after 500 (all of the observation), drift-scaled at 1000
  T4  contact envs: random-policy return / length distributions, 4096 CUDA vs 1024 oracle episodes, two-sample KS, p > 0.01
The oracle's physics is parity-unpinned against pybullet (oracle/oracle.h); its task layer is pinned
by tests/test_golden_task.py.

Why T2's per-sample bound is relative to the oracle's own sensitivity: the dynamics themselves amplify a
1e-7 (fp32 storage) perturbation of the state into up to ~4e-3 after one env step when stiff contacts /
friction-cone switches are active (test_oracle_sensitivity_reference), so a fixed max-norm bound at fp32
round-off level is not meaningful there; contact-free steps are held to 1e-4 in max norm.
"""
import dataclasses

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

IDS = ["InvertedPendulumPyBulletEnv-v0", "InvertedDoublePendulumPyBulletEnv-v0", "ReacherPyBulletEnv-v0", "HopperPyBulletEnv-v0", "Walker2DPyBulletEnv-v0", "HalfCheetahPyBulletEnv-v0",
       "AntPyBulletEnv-v0", "HumanoidPyBulletEnv-v0"]
FLAGRUN_IDS = ["HumanoidFlagrunPyBulletEnv-v0", "HumanoidFlagrunHarderPyBulletEnv-v0"]
PHYS_IDS = IDS + FLAGRUN_IDS + ["HalfCheetahMuJoCoEnv-v0"]      # the latter: torsional friction rows
TASK_IDS = IDS + ["HumanoidFlagrunPyBulletEnv-v0", "HumanoidFlagrunHarderPyBulletEnv-v0", "InvertedDoublePendulumMuJoCoEnv-v0",
                  "HopperMuJoCoEnv-v0", "Walker2DMuJoCoEnv-v0", "HalfCheetahMuJoCoEnv-v0", "AntMuJoCoEnv-v0", "HumanoidMuJoCoEnv-v0"]
E = 48


def _mk(env_id, n=E, spec=None, auto_reset=False, seed=1):
    from pybullet_gym_b200.vector_env import VectorEnv
    return VectorEnv(env_id, n, device="cuda:0", seed=seed, auto_reset=auto_reset, spec=spec)


def _oracles(oracle_lib, env_id, n, spec=None):
    from pybullet_gym_b200 import _lib
    from pybullet_gym_b200.spec import SPECS
    budget = _lib.solver_budget(SPECS[env_id].kind)      # same contact cap / row budget as the kernel
    # same RNG key as _mk(): seed 1, stream = env index (matters for the Flagrun target draws)
    return [oracle_lib.OracleEnv(spec if spec is not None else env_id, seed=1, env_index=i, **budget) for i in range(n)]


def _rel(g, o):
    return np.abs(g - o) / (1.0 + np.abs(o))


@pytest.mark.parametrize("env_id", IDS)
def test_reset_matches_oracle(env_id, oracle_lib):
    env = _mk(env_id)
    rng = np.random.default_rng(0)
    noise = rng.uniform(-0.1, 0.1, (E, env.noise_dim)).astype(np.float32)
    for floor in (False, True):
        obs = env.reset(joint_noise=torch.from_numpy(noise), floor_in_parts=floor).cpu().numpy()
        orcs = _oracles(oracle_lib, env_id, E)
        oobs = np.stack([o.reset(noise=noise[i].astype(np.float64), floor_in_parts=floor) for i, o in enumerate(orcs)])
        assert np.abs(obs - oobs).max() < 1e-5
        st = env.get_state().cpu().numpy()
        ost = np.stack([o.get_state() for o in orcs])
        assert np.abs(st - ost).max() < 1e-6


@pytest.mark.parametrize("env_id", TASK_IDS)
def test_device_rng_matches_oracle_rng(env_id, oracle_lib):
    """pbg_reset's Philox draws are bit-identical to the oracle's (same seed / env index / episode)."""
    env = _mk(env_id, n=8, seed=7)
    obs = env.reset(floor_in_parts=True).cpu().numpy()
    for i in range(8):
        o = oracle_lib.OracleEnv(env_id, seed=7, env_index=i)
        assert np.abs(o.reset(floor_in_parts=True) - obs[i]).max() < 1e-5


@pytest.mark.parametrize("env_id", TASK_IDS)
def test_observation_reward_parity_T1(env_id, oracle_lib):
    """T1: calc_state + reward terms + done on identical states, states sampled along oracle rollouts.  The whole
    observation is compared, feet-flag tail included.  The flags calc_state sees are the ones each side latched from its OWN
    previous physics step (quirk Q2), so an env takes part in the tail / HalfCheetah-done comparison only while the two sides'
    flags agree; test_feet_flags_from_identical_state holds the flags themselves to the oracle."""
    env = _mk(env_id)
    rng = np.random.default_rng(3)
    nA = env.action_dim
    noise = rng.uniform(-0.1, 0.1, (E, env.noise_dim)).astype(np.float32)
    env.reset(joint_noise=torch.from_numpy(noise), floor_in_parts=True)
    orcs = _oracles(oracle_lib, env_id, E)
    for i, o in enumerate(orcs):
        o.reset(noise=noise[i].astype(np.float64), floor_in_parts=True)
    nf = len(env.spec.foot_list)
    worst_obs = worst_terms = worst_prog = worst_tail = 0.0
    same_flags = np.ones(E, bool)          # flags latched by the previous observe agree (both start from zeros)
    n_done_cmp = n_done_true = n_tail_cmp = 0
    for t in range(30):
        a = (1.4 * rng.uniform(-1, 1, (E, nA))).astype(np.float32)          # |a| > 1: quirk Q3
        # advance both on their own physics (keeps feet flags / potentials in step), then pin the state
        env.physics_step(torch.from_numpy(a))
        for i, o in enumerate(orcs):
            o.physics_step(a[i].astype(np.float64))
        ost = np.stack([o.get_state() for o in orcs]).astype(np.float32)
        env.set_state(torch.from_numpy(ost))
        for i, o in enumerate(orcs):
            o.set_state(ost[i].astype(np.float64))
        gobs, grew, gdone, gterms = [x.cpu().numpy() for x in env.observe(torch.from_numpy(a))]
        res = [o.observe(a[i].astype(np.float64)) for i, o in enumerate(orcs)]
        oobs = np.stack([r[0] for r in res])
        oterms = np.stack([r[3] for r in res])
        odone = np.array([r[2] for r in res])
        margin = np.array([o.done_margin() for o in orcs])
        body = slice(0, oobs.shape[1] - nf) if (nf and env.spec.kind < 11) else slice(None)
        worst_obs = max(worst_obs, np.abs(gobs[:, body] - oobs[:, body]).max())
        if nf and env.spec.kind < 11 and same_flags.any():
            tail = slice(oobs.shape[1] - nf, oobs.shape[1])
            worst_tail = max(worst_tail, np.abs(gobs[same_flags, tail] - oobs[same_flags, tail]).max())
            n_tail_cmp += int(same_flags.sum())
        if env.spec.kind in (12, 13, 16):
            # MuJoCo-style walkers: terms[0] = dx / dt carries the fp32 error of x itself (x / 0.0165 * 6e-8)
            x = np.abs(ost[:, 0])
            worst_prog = max(worst_prog, (np.abs(gterms[:, 0] - oterms[:, 0]) / (1.0 + x)).max())
            gterms[:, 0] = oterms[:, 0]
        # done: exact, except where the oracle's own test sits within fp32 round-off of its threshold (and, for the
        # HalfCheetah, where the stale flags it reads differed)
        cmp_ok = same_flags if env.spec.kind == 4 else np.ones(E, bool)
        bad = (gdone.astype(bool) != odone) & cmp_ok
        assert not (bad & (margin > 1e-5)).any(), (t, np.nonzero(bad)[0], margin[bad])
        n_done_cmp += int(cmp_ok.sum()); n_done_true += int(odone.sum())
        agree = ~bad & cmp_ok                  # the alive term flips with done: compare it where the decisions agree
        worst_terms = max(worst_terms, np.abs(gterms[:, [2, 3, 4]] - oterms[:, [2, 3, 4]]).max())
        if agree.any():
            worst_terms = max(worst_terms, np.abs(gterms[agree, 0] - oterms[agree, 0]).max())
        if 2 <= env.spec.kind <= 8 or env.spec.kind in (14, 15):
            x = np.abs(ost[:, 0])
            worst_prog = max(worst_prog, (np.abs(gterms[:, 1] - oterms[:, 1]) / (1.0 + x)).max())
        if nf:
            gf = env.feet_contact().cpu().numpy()
            of = np.stack([o.feet_contact() for o in orcs])
            same_flags = (gf == of).all(axis=1)
    assert worst_obs < 1e-5, worst_obs
    assert worst_tail == 0.0, worst_tail
    assert worst_terms < 1e-5, worst_terms
    assert worst_prog < 2e-5, worst_prog
    if nf and env.spec.kind < 11:
        assert n_tail_cmp > 0.8 * 30 * E, n_tail_cmp          # the excuse is the exception
    assert n_done_cmp > 0.8 * 30 * E


@pytest.mark.parametrize("env_id", ["HopperPyBulletEnv-v0", "Walker2DPyBulletEnv-v0", "HalfCheetahPyBulletEnv-v0", "AntPyBulletEnv-v0",
                                    "HumanoidPyBulletEnv-v0", "HumanoidFlagrunHarderPyBulletEnv-v0"])
def test_feet_flags_from_identical_state(env_id, oracle_lib):
    """robot.feet_contact (rs/gym_locomotion_envs.py:69-78) after one physics step from an identical state, CUDA vs oracle.
    With frame_skip = 1 the step's only collision pass runs on the identical state itself, so the flags must agree exactly
    except for a foot whose nearest candidate lies within 1e-6 m of its breaking threshold (fp32 FK round-off); with the
    real frame_skip = 4 the last pass runs after three sub-steps of each side's own dynamics, and a disagreement is
    excused only within 10x that env's own position difference after the step (+1e-6)."""
    from pybullet_gym_b200.spec import SPECS
    spec = SPECS[env_id]
    spec1 = dataclasses.replace(spec, scene=dataclasses.replace(spec.scene, frame_skip=1))
    env4, env1 = _mk(env_id), _mk(env_id, spec=spec1)
    rng = np.random.default_rng(17)
    nA = env4.action_dim
    noise = rng.uniform(-0.1, 0.1, (E, env4.noise_dim)).astype(np.float32)
    env4.reset(joint_noise=torch.from_numpy(noise), floor_in_parts=True); env1.reset(joint_noise=torch.from_numpy(noise), floor_in_parts=True)
    orcs, orcs1 = _oracles(oracle_lib, env_id, E), _oracles(oracle_lib, env_id, E, spec=spec1)
    for i in range(E):
        orcs[i].reset(noise=noise[i].astype(np.float64), floor_in_parts=True)
        orcs1[i].reset(noise=noise[i].astype(np.float64), floor_in_parts=True)
    nflag = n_on = n_excused1 = n_excused4 = 0
    for t in range(60):
        a = rng.uniform(-1, 1, (E, nA)).astype(np.float32)
        ost = np.stack([o.get_state() for o in orcs]).astype(np.float32)
        ta = torch.from_numpy(a)
        for envx in (env4, env1):
            envx.set_state(torch.from_numpy(ost))
            envx.physics_step(ta)
        g4 = env4.get_state().cpu().numpy()
        env4.observe(ta); env1.observe(ta)               # latches the step's flags (quirk Q2)
        gf4, gf1 = env4.feet_contact().cpu().numpy(), env1.feet_contact().cpu().numpy()
        of4, of1, m4, m1, o4 = [], [], [], [], []
        for i in range(E):
            for o, fl, mg in ((orcs[i], of4, m4), (orcs1[i], of1, m1)):
                o.set_state(ost[i].astype(np.float64))
                o.physics_step(a[i].astype(np.float64))
                mg.append(o.feet_margin())
                if o is orcs[i]:
                    o4.append(o.get_state())
                o.observe(a[i].astype(np.float64))
                fl.append(o.feet_contact())
        of4, of1, m4, m1, o4 = np.stack(of4), np.stack(of1), np.stack(m4), np.stack(m1), np.stack(o4)
        bad1 = gf1 != of1
        assert not (bad1 & (m1 > 1e-6)).any(), (t, np.argwhere(bad1), m1[bad1])
        poserr = np.abs(g4 - o4).max(axis=1)[:, None]
        bad4 = gf4 != of4
        assert not (bad4 & (m4 > 10 * poserr + 1e-6)).any(), (t, np.argwhere(bad4), m4[bad4], poserr[bad4.any(axis=1)])
        nflag += of1.size; n_on += int(of1.sum()); n_excused1 += int(bad1.sum()); n_excused4 += int(bad4.sum())
    assert n_on > 0.05 * nflag and n_on < 0.98 * nflag, (n_on, nflag)       # both outcomes are exercised
    assert n_excused1 <= 0.001 * nflag and n_excused4 <= 0.01 * nflag, (n_excused1, n_excused4, nflag)


def _throw_cube_at_robots(ost, rng, frac=0.5):
    """FlagrunHarder T2 states: put the cube next to a fraction of the robots, moving at them (canonical state:
    [base pos3 quat4 omega3 vel3][q 17][qd 17][cube pos3 quat4 omega3 vel3])."""
    n = ost.shape[0]
    pick = rng.uniform(size=n) < frac
    ang = rng.uniform(-np.pi, np.pi, n)
    d = np.stack([np.cos(ang), np.sin(ang), np.zeros(n)], 1)
    pos = ost[:, 0:3] + 0.30 * d + np.stack([np.zeros(n), np.zeros(n), rng.uniform(-0.35, 0.25, n)], 1)
    vel = -d * rng.uniform(2.0, 12.0, n)[:, None]
    ost = ost.copy()
    ost[pick, -13:-10] = pos[pick]
    ost[pick, -6:-3] = 0.0
    ost[pick, -3:] = vel[pick]
    return ost


T2_K = 500.0            # per-sample bound: K x (oracle's response to 1e-7 perturbations of the same state), floored at T2_FLOOR
T2_FLOOR = 1e-4


@pytest.mark.parametrize("env_id", PHYS_IDS)
def test_single_substep_and_step_parity_T2(env_id, oracle_lib):
    """T2: one sub-step (frame_skip = 1) and one env step (4 sub-steps) from identical states: ABA-equivalent dynamics, limit /
    contact rows, 5 PGS sweeps, integration.  Every sample is bounded: its error may not exceed T2_K times what NPERT 1e-7-sized
    perturbations of that same state do to the double-precision oracle (floor T2_FLOOR); medians are held to fixed bounds."""
    from pybullet_gym_b200.spec import SPECS
    spec = SPECS[env_id]
    spec1 = dataclasses.replace(spec, scene=dataclasses.replace(spec.scene, frame_skip=1))
    env4, env1 = _mk(env_id), _mk(env_id, spec=spec1)
    rng = np.random.default_rng(5)
    nA = env4.action_dim
    NPERT = 4
    noise = rng.uniform(-0.1, 0.1, (E, env4.noise_dim)).astype(np.float32)
    env4.reset(joint_noise=torch.from_numpy(noise)); env1.reset(joint_noise=torch.from_numpy(noise))
    orcs = _oracles(oracle_lib, env_id, E)
    orcs1 = _oracles(oracle_lib, env_id, E, spec=spec1)
    twin4 = [_oracles(oracle_lib, env_id, E) for _ in range(NPERT)]
    twin1 = [_oracles(oracle_lib, env_id, E, spec=spec1) for _ in range(NPERT)]
    for i, o in enumerate(orcs):
        for x in [o, orcs1[i]] + [tw[i] for tw in twin4 + twin1]:
            x.reset(noise=noise[i].astype(np.float64))
    errs1, errs4, free4, sens1, sens4 = [], [], [], [], []
    for t in range(40):
        a = rng.uniform(-1, 1, (E, nA)).astype(np.float32)
        ost = np.stack([o.get_state() for o in orcs]).astype(np.float32)
        if env_id == "HumanoidFlagrunHarderPyBulletEnv-v0" and t % 4 == 1:
            ost = _throw_cube_at_robots(ost, rng).astype(np.float32)
        for envx in (env4, env1):
            envx.set_state(torch.from_numpy(ost))
        n4 = env4.physics_step(torch.from_numpy(a), want_contacts=True).cpu().numpy()
        env1.physics_step(torch.from_numpy(a))
        g4, g1 = env4.get_state().cpu().numpy(), env1.get_state().cpu().numpy()
        pert = [rng.normal(size=ost.shape) * 1e-7 * (1 + np.abs(ost)) for _ in range(NPERT)]
        o4, o1, s4, s1, on4 = [], [], [], [], []
        for i in range(E):
            ai, si = a[i].astype(np.float64), ost[i].astype(np.float64)
            orcs[i].set_state(si); orcs1[i].set_state(si)
            orcs[i].physics_step(ai); orcs1[i].physics_step(ai)
            r4, r1 = orcs[i].get_state(), orcs1[i].get_state()
            d4 = d1 = 0.0
            for k in range(NPERT):
                twin4[k][i].set_state(si + pert[k][i]); twin1[k][i].set_state(si + pert[k][i])
                twin4[k][i].physics_step(ai); twin1[k][i].physics_step(ai)
                d4 = max(d4, _rel(twin4[k][i].get_state(), r4).max()); d1 = max(d1, _rel(twin1[k][i].get_state(), r1).max())
            o4.append(r4); o1.append(r1); s4.append(d4); s1.append(d1); on4.append(orcs[i].num_contacts())
        o4, o1, on4 = np.stack(o4), np.stack(o1), np.array(on4)
        errs1.append(_rel(g1, o1).max(axis=1)); errs4.append(_rel(g4, o4).max(axis=1))
        sens1.append(np.array(s1)); sens4.append(np.array(s4))
        free4.append((on4 == 0) & (n4 == 0))
        assert np.isfinite(g4).all()
    e1, e4, fr = np.concatenate(errs1), np.concatenate(errs4), np.concatenate(free4)
    b1, b4 = np.maximum(T2_FLOOR, T2_K * np.concatenate(sens1)), np.maximum(T2_FLOOR, T2_K * np.concatenate(sens4))
    print("\n  [T2 %s] sub-step: median %.1e max %.1e worst err/bound %.2f | step: median %.1e max %.1e worst err/bound %.2f"
          % (env_id, np.median(e1), e1.max(), (e1 / b1).max(), np.median(e4), e4.max(), (e4 / b4).max()))
    assert np.median(e1) < 2e-5 and np.median(e4) < 1e-4, (np.median(e1), np.median(e4))
    assert (e1 <= b1).all(), ("sub-step", np.argmax(e1 / b1), (e1 / b1).max(), e1.max())
    assert (e4 <= b4).all(), ("step", np.argmax(e4 / b4), (e4 / b4).max(), e4.max())
    if env_id in ("InvertedPendulumPyBulletEnv-v0", "InvertedDoublePendulumPyBulletEnv-v0", "ReacherPyBulletEnv-v0",
                  "AntPyBulletEnv-v0") and fr.any():
        # contact-free env steps (pendulum always; Ant while airborne) are held to a max-norm bound
        assert e4[fr].max() < 1e-4, e4[fr].max()


GAIT_IDS = ["HopperPyBulletEnv-v0", "HalfCheetahPyBulletEnv-v0", "AntPyBulletEnv-v0", "HumanoidPyBulletEnv-v0", "HumanoidFlagrunPyBulletEnv-v0"]


@pytest.mark.parametrize("env_id", GAIT_IDS)
def test_single_step_parity_on_policy_gait_states_T2(env_id, oracle_lib):
    """T2 on the states of a walking gait instead of random flailing: every state of the policy-driven golden episodes
    (tests/golden/task_*Policy.json, the reference's pretrained MLP in the loop: stance / swing phases, feet touching down, joints
    against their stops) is one env of a batch; ONE physics step of the kernel with the recorded action against the oracle from
    the identical (float32-rounded) state, both with empty warm-start caches.  Same bars as T2: every sample within T2_K x the
    oracle's own response to 1e-7 perturbations of that state (floor T2_FLOOR), median below 1e-4.  Unlike a free-running replay
    this covers ALL steps of the episodes -- chaos cannot end the comparison."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "task_%sPolicy.json" % env_id.split("PyBullet")[0])))
    ep = g["episodes"][0]
    ref = oracle_lib.OracleEnv(env_id)                 # the fixture's oracle: no solver budget, replays the recorded trajectory
    if ep.get("tape"):
        ref.set_tape(ep["tape"])
    ref.reset(noise=ep["noise"], floor_in_parts=False)
    states, actions = [], []
    for st in ep["steps"]:
        states.append(ref.get_state().astype(np.float32))
        actions.append(st["a"])
        ref.step(st["a"])
    S, A = np.stack(states), np.array(actions, np.float32)
    n = len(S)
    env = _mk(env_id, n=n)
    env.reset()
    env.set_state(torch.from_numpy(S))
    env.physics_step(torch.from_numpy(A))
    gk = env.get_state().cpu().numpy().astype(np.float64)
    assert np.isfinite(gk).all()
    o = _oracles(oracle_lib, env_id, 1)[0]             # the kernel's contact cap / row budget
    rng = np.random.default_rng(11)
    errs, bounds = [], []
    for i in range(n):
        si, ai = S[i].astype(np.float64), A[i].astype(np.float64)
        o.reset(noise=np.zeros(env.noise_dim)); o.set_state(si); o.physics_step(ai)
        r = o.get_state().copy()
        err, sens, draws = _rel(gk[i], r).max(), 0.0, 0
        # 4 perturbations per state; a state that misses its bound gets 16 more before it counts: a gait holds joints AT their stops
        # (the Hopper's thigh sits at q = 0 within 1e-9 for dozens of steps), where a limit row switches on or off with the sign of a
        # round-off and only about half of the 1e-7 draws land on the other side
        while draws < 4 or (draws < 20 and err > max(T2_FLOOR, T2_K * sens)):
            o.reset(noise=np.zeros(env.noise_dim)); o.set_state(si + rng.normal(size=si.shape) * 1e-7 * (1 + np.abs(si))); o.physics_step(ai)
            sens = max(sens, _rel(o.get_state(), r).max())
            draws += 1
        errs.append(err); bounds.append(max(T2_FLOOR, T2_K * sens))
    e, b = np.array(errs), np.array(bounds)
    print("\n  [T2 gait %s] %d states: median %.1e max %.1e worst err/bound %.2f" % (env_id, n, np.median(e), e.max(), (e / b).max()))
    assert np.median(e) < 1e-4, np.median(e)
    assert (e <= b).all(), (int(np.argmax(e / b)), (e / b).max(), e.max())


def test_oracle_sensitivity_reference(oracle_lib):
    """Documents the conditioning T2's per-sample bounds rest on: the double-precision oracle, perturbed
    by a float32-storage-sized 1e-7, moves by > 1e-5 after one Ant env step once contacts are active."""
    rng = np.random.default_rng(0)
    a_env, b_env = oracle_lib.OracleEnv("AntPyBulletEnv-v0", max_contacts=8), oracle_lib.OracleEnv("AntPyBulletEnv-v0", max_contacts=8)
    a_env.reset(noise=rng.uniform(-.1, .1, 8)); b_env.reset(noise=np.zeros(8))
    worst = 0.0
    for t in range(60):
        act = rng.uniform(-1, 1, 8)
        s = a_env.get_state().astype(np.float32).astype(np.float64)
        a_env.set_state(s); b_env.set_state(s + rng.normal(size=s.size) * 1e-7 * (1 + np.abs(s)))
        a_env.physics_step(act); b_env.physics_step(act)
        worst = max(worst, _rel(b_env.get_state(), a_env.get_state()).max())
    assert worst > 1e-5


def test_pendulum_full_episode_trajectory_T3(oracle_lib):
    """Contact-free cart-pole, 1000 steps, fixed action tape (small actions keep it from diverging fast): the whole
    observation (x, vx, cos, sin, theta_dot) within 2e-4 after 100 steps, 1e-3 after 500, and after the full episode within
    max(1e-3, 100 x the drift of the double-precision oracle under a 1e-7 perturbation of the reset angle)."""
    env_id = "InvertedPendulumSwingupPyBulletEnv-v0"      # never terminates: a full 1000-step episode
    n = 16
    env = _mk(env_id, n=n)
    rng = np.random.default_rng(11)
    noise = rng.uniform(-0.1, 0.1, (n, 1)).astype(np.float32)
    env.reset(joint_noise=torch.from_numpy(noise))
    orcs = [oracle_lib.OracleEnv(env_id) for _ in range(n)]
    twins = [[oracle_lib.OracleEnv(env_id) for _ in range(n)] for _ in range(3)]
    for i, o in enumerate(orcs):
        o.reset(noise=noise[i].astype(np.float64))
        for tw in twins:
            tw[i].reset(noise=noise[i].astype(np.float64) + rng.normal() * 1e-7)
    tape = rng.uniform(-1, 1, (1000, n, 1)).astype(np.float32) * 0.3
    err100 = err500 = drift = 0.0
    for t in range(1000):
        obs, rew, done, _ = env.step(torch.from_numpy(tape[t]))
        res = [o.step(tape[t, i].astype(np.float64)) for i, o in enumerate(orcs)]
        for tw in twins:
            for i in range(n):
                drift = max(drift, np.abs(tw[i].step(tape[t, i].astype(np.float64))[0] - res[i][0]).max())
        if t == 99:
            err100 = np.abs(obs.cpu().numpy() - np.stack([r[0] for r in res])).max()
        if t == 499:
            err500 = np.abs(obs.cpu().numpy() - np.stack([r[0] for r in res])).max()
    oobs = np.stack([r[0] for r in res])
    gobs = obs.cpu().numpy()
    assert err100 < 2e-4, err100                       # 100 steps: fp32 round-off accumulates ~1e-6 per step
    assert err500 < 1e-3, err500                       # 500 steps, every component
    # full episode: the swinging pole is mildly chaotic -- a 1e-7 perturbation of the reset angle moves the double-precision
    # oracle itself by ~1e-3 after 1000 steps -- so the bound follows that drift
    bound = max(1e-3, 100.0 * drift)
    assert np.abs(gobs - oobs).max() < bound, (np.abs(gobs - oobs).max(axis=0), drift)
    assert np.isfinite(gobs).all()


def _ks_pvalue(a, b):
    from scipy.stats import ks_2samp
    return ks_2samp(a, b).pvalue


T4_CAP = {"HopperPyBulletEnv-v0": 1000, "Walker2DPyBulletEnv-v0": 1000, "HalfCheetahPyBulletEnv-v0": 1000,
          "AntPyBulletEnv-v0": 300, "HumanoidPyBulletEnv-v0": 300, "HumanoidFlagrunPyBulletEnv-v0": 300,
          "HumanoidFlagrunHarderPyBulletEnv-v0": 300, "HalfCheetahMuJoCoEnv-v0": 200}


def _t4_sample(env_id, oracle_lib, n, m, cap, gpu_seed, act_seed, orc_seed):
    env = _mk(env_id, n=n, seed=gpu_seed)
    env.reset(floor_in_parts=True)
    gen = torch.Generator(device="cuda").manual_seed(act_seed)
    ret = torch.zeros(n, device="cuda"); length = torch.zeros(n, device="cuda"); alive = torch.ones(n, device="cuda")
    for t in range(cap):
        a = torch.rand(n, env.action_dim, device="cuda", generator=gen) * 2 - 1
        obs, rew, done = env.step_fast(a)
        ret += alive * rew; length += alive
        alive = alive * (1 - done.float())
        if t % 50 == 49 and float(alive.sum()) == 0:
            break
    g_ret, g_len = ret.cpu().numpy(), length.cpu().numpy()
    from pybullet_gym_b200 import _lib
    from pybullet_gym_b200.spec import SPECS
    o_ret, o_len = oracle_lib.random_policy_episodes(env_id, m, cap, seed=orc_seed, **_lib.solver_budget(SPECS[env_id].kind))
    return g_len, o_len, g_ret, o_ret


@pytest.mark.parametrize("env_id", list(T4_CAP))
def test_random_policy_distributions_T4(env_id, oracle_lib):
    """T4: episode length and return distributions under U(-1,1) actions, 4096 CUDA episodes vs 1024 oracle episodes
    (two-sample KS, p > 0.01 on both).  Whole episodes (TimeLimit 1000) for Hopper / Walker2D / HalfCheetah; the first 300
    steps for the Ant and the Humanoids, whose random-policy episodes are cut by the cap rarely / never.
    With 14 p-values per run a true-null p < 0.01 turns up in one run out of seven, so a low p-value is re-drawn once with
    fresh seeds on both sides and the test fails only if the second, independent sample is below 0.01 as well (false alarm
    1e-4 per quantity; a real shift fails both).  tools/t4_power.py is the same comparison at 16384 vs 8192 episodes."""
    n, m, cap = 4096, 1024, T4_CAP[env_id]
    for attempt, (gs, as_, os_) in enumerate(((100, 0, 200), (101, 1, 201))):
        g_len, o_len, g_ret, o_ret = _t4_sample(env_id, oracle_lib, n, m, cap, gs, as_, os_)
        p_len, p_ret = _ks_pvalue(g_len, o_len), _ks_pvalue(g_ret, o_ret)
        print("\n  [T4 %s] len %.1f vs %.1f  return %.2f vs %.2f  p_len %.3f p_ret %.3f%s" % (
            env_id, g_len.mean(), o_len.mean(), g_ret.mean(), o_ret.mean(), p_len, p_ret, "  (re-draw)" if attempt else ""))
        if p_len > 0.01 and p_ret > 0.01:
            return
    assert False, (p_len, p_ret, g_len.mean(), np.mean(o_len), g_ret.mean(), np.mean(o_ret))


GOLD_K = 500.0          # the same factor as T2_K: float32 arithmetic against 1e-7 perturbations of the double-precision oracle


@pytest.mark.parametrize("env_id", IDS)
def test_golden_reference_rollouts_whole_episodes(env_id, oracle_lib):
    """The CUDA path replays the reset noise / actions of the golden files (recorded from the reference's own Python on
    oracle physics) over the WHOLE recorded episodes, free running.  Reset observation to 1e-5.  At step t the deviation
    may not exceed GOLD_K x the drift of the double-precision oracle itself when its reset noise is perturbed by 1e-7
    (max over 3 twins; floor 2e-5 (t + 1)): contact dynamics amplify round-off, so the bound follows the trajectory's own
    conditioning instead of a fixed number; an episode's comparison ends where that bound passes 2e-2 (the oracle itself has become
    unpredictable there).  A step that exceeds its bound is tolerated once per fixture, and ends its episode's window, only if that
    very step is exact from the CUDA path's own pre-step state (T2's per-sample bound): a discrete event hit 2e-5 off the oracle's
    trajectory.  done flags must agree wherever the bound is below 1e-3."""
    import os
    _replay_golden(env_id, os.path.join(os.path.dirname(__file__), "golden", "task_%s.json" % env_id.split("PyBullet")[0]), oracle_lib)


# (task_HumanoidFlagrunPolicy.json is replayed by the oracle only, tests/test_golden_task.py: its flag positions are the reference's
# np_random draws, which the oracle takes from the fixture's tape and the kernel draws from its own counter RNG)
POLICY_IDS = {"HopperPyBulletEnv-v0": 30, "HalfCheetahPyBulletEnv-v0": 15, "AntPyBulletEnv-v0": 10, "HumanoidPyBulletEnv-v0": 5}


@pytest.mark.parametrize("env_id", list(POLICY_IDS))
def test_golden_policy_rollouts(env_id, oracle_lib):
    """The same free-running replay on the policy-driven fixtures (task_*Policy.json: the reference's Python task layer stepped
    with the reference's own pretrained MLP, i.e. walking gaits with feet touching down and lifting instead of random flailing).
    These episodes are 100-300 steps long, far beyond the horizon over which ANY two integrations of a walking robot stay
    comparable: three oracle twins 1e-7 apart at the reset are 4e-5 apart (bound 2e-2) after 41 Hopper, 22 HalfCheetah, 15 Ant and 7
    Humanoid steps, and that is where the comparison has to end.  Inside the window the bar is the same per-step bound (observed
    worst err / bound 0.16, 0.04, 0.008, 0.006); the minimum window lengths below guard against the window collapsing.  The whole
    fixtures are replayed bit for bit by the oracle and checked against the reference's numbers in tests/test_golden_task.py."""
    import os
    _replay_golden(env_id, os.path.join(os.path.dirname(__file__), "golden", "task_%sPolicy.json" % env_id.split("PyBullet")[0]),
                   oracle_lib, min_frac=0.0, min_steps=POLICY_IDS[env_id])


def _replay_golden(env_id, path, oracle_lib, min_frac=0.25, min_steps=0):
    import json
    g = json.load(open(path))
    env, chk = _mk(env_id, n=1), _mk(env_id, n=1)
    chk.reset()
    rng = np.random.default_rng(23)
    worst_ratio = 0.0
    nsteps = events = 0
    for ei, ep in enumerate(g["episodes"]):
        noise = np.array(ep["noise"], np.float64)
        obs0 = env.reset(joint_noise=torch.tensor([ep["noise"]], dtype=torch.float32), floor_in_parts=ei > 0)
        assert np.abs(obs0.cpu().numpy()[0] - np.array(ep["obs0"])).max() < 1e-5
        twins = [(oracle_lib.OracleEnv(env_id), 1e-7, GOLD_K) for i in range(3)]
        for tw, eps, _ in twins:
            tw.reset(noise=noise + rng.normal(size=noise.shape) * eps, floor_in_parts=ei > 0)
        drift = 0.0
        for t, st in enumerate(ep["steps"]):
            a = np.array(st["a"], np.float64)
            gold = np.array(st["obs"])
            prev_state = env.get_state().clone()
            obs, rew, done, info = env.step(torch.tensor([st["a"]], dtype=torch.float32))
            for tw, _, kk in twins:
                drift = max(drift, kk / GOLD_K * np.abs(tw.step(a)[0] - gold).max())          # running max: drift does not shrink
            bound = max(2e-5 * (t + 1), GOLD_K * drift)
            if bound > 2e-2:
                break        # the trajectory has become unpredictable: 1e-7 moves the oracle itself by > 4e-5; nothing left to compare
            err = np.abs(obs.cpu().numpy()[0] - gold).max()
            if err > bound:
                # A discrete event (a joint reaching its stop, a foot touching down one sub-step earlier) answers a trajectory that
                # arrives 2e-5 away out of proportion, and the three 1e-7 twins need not catch it (HalfCheetah episode 3, step 23:
                # the twins move by 8e-6, every float32 build tried by 1e-3 .. 3e-3, and a 1e-7 perturbation of the state right
                # before that step moves the oracle by 2e-2).  The deviation is accepted -- and this episode's free-running comparison
                # ends -- only if the step itself is exact: from the CUDA path's OWN state before the step, kernel and oracle must
                # agree under T2's per-sample bound.
                chk.set_state(prev_state)
                chk.physics_step(torch.tensor([st["a"]], dtype=torch.float32))
                gk = chk.get_state().cpu().numpy()[0].astype(np.float64)
                si = prev_state.cpu().numpy()[0].astype(np.float64)
                loc = oracle_lib.OracleEnv(env_id)
                loc.reset(noise=noise, floor_in_parts=ei > 0)
                loc.set_state(si); loc.physics_step(a)
                ro = loc.get_state()
                sens = 0.0
                for k in range(4):
                    loc.set_state(si + rng.normal(size=si.shape) * 1e-7 * (1 + np.abs(si))); loc.physics_step(a)
                    sens = max(sens, _rel(loc.get_state(), ro).max())
                lerr, lbound = _rel(gk, ro).max(), max(T2_FLOOR, T2_K * sens)
                print("\n  [golden %s] episode %d step %d: free-running error %.1e above its bound %.1e; the step from the kernel's own "
                      "state agrees with the oracle to %.1e (bound %.1e): window ends here" % (env_id, ei, t, err, bound, lerr, lbound))
                assert lerr <= lbound, (ei, t, err, bound, lerr, lbound)
                events += 1
                break
            worst_ratio = max(worst_ratio, err / bound)
            assert abs(float(rew[0]) - st["reward"]) <= 100 * bound + 1e-4, (ei, t, float(rew[0]), st["reward"], bound)
            if bound < 1e-3:
                assert bool(done[0]) == st["done"], (ei, t)
            nsteps += 1
    print("\n  [golden %s] %d steps, worst err / bound %.3f" % (env_id, nsteps, worst_ratio))
    total = sum(len(ep["steps"]) for ep in g["episodes"])
    print("  [golden %s] compared %d of %d recorded steps" % (env_id, nsteps, total))
    assert nsteps >= min_frac * total and nsteps >= min_steps      # a good part of every fixture lies inside the comparable window
    assert events <= 1                   # at most one episode of a fixture may end on such an event
