#!/usr/bin/env python
"""Generate tests/golden/task_<env>.json by running the reference's UNMODIFIED Python task layer.

The reference modules under /root/reference/pybulletgym/envs/roboschool/ are imported as they are,
on top of the stub pybullet client of tools/fake_pybullet.py (physics = our CPU oracle).  What gets
recorded is everything the reference computes *above* the pybullet boundary for a scripted rollout:

  per episode : the reset noise the reference drew (np_random.uniform log), the reset observation
  per step    : action, physics state after the step (canonical layout), observation, reward,
                done, the reward-term list `env.rewards`, robot.feet_contact, joints_at_limit

tests/test_golden_task.py replays the same noise/actions through the C oracle (whose physics is
the same code, so states agree bit for bit) and requires the oracle's task layer to reproduce the
reference's numbers; this pins SURVEY.md Appendix A incl. quirks Q1-Q3, Q6, Q10 to reference source.

Run in the build container only (needs /root/reference).  Usage: python tools/gen_golden_task.py [--out DIR]
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fake_pybullet as fp  # noqa: E402

REF = "/root/reference"
GOLDEN = os.path.join(HERE, "..", "tests", "golden")
# `--out DIR` writes the fixtures somewhere else (tests/test_golden_task.py regenerates them into a scratch directory and compares bytes)
OUT = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else GOLDEN

CASES = {
    # env id -> (module, class, episodes, max steps per episode, action scale)
    "InvertedPendulumPyBulletEnv-v0": ("gym_pendulum_envs", "InvertedPendulumBulletEnv", 3, 40, 1.0),
    "InvertedPendulumSwingupPyBulletEnv-v0": ("gym_pendulum_envs", "InvertedPendulumSwingupBulletEnv", 2, 40, 1.0),
    "InvertedDoublePendulumPyBulletEnv-v0": ("gym_pendulum_envs", "InvertedDoublePendulumBulletEnv", 3, 60, 0.3),
    "InvertedDoublePendulumMuJoCoEnv-v0": ("mujoco.gym_pendulum_envs", "InvertedDoublePendulumMuJoCoEnv", 3, 60, 0.3),
    "HopperMuJoCoEnv-v0": ("mujoco.gym_locomotion_envs", "HopperMuJoCoEnv", 6, 40, 0.2),
    "Walker2DMuJoCoEnv-v0": ("mujoco.gym_locomotion_envs", "Walker2DMuJoCoEnv", 4, 40, 1.3),
    "HalfCheetahMuJoCoEnv-v0": ("mujoco.gym_locomotion_envs", "HalfCheetahMuJoCoEnv", 3, 60, 1.3),
    "AntMuJoCoEnv-v0": ("mujoco.gym_locomotion_envs", "AntMuJoCoEnv", 3, 40, 1.3),
    "HumanoidMuJoCoEnv-v0": ("mujoco.gym_locomotion_envs", "HumanoidMuJoCoEnv", 3, 40, 1.3),
    "ReacherPyBulletEnv-v0": ("gym_manipulator_envs", "ReacherBulletEnv", 3, 60, 1.3),
    "HopperPyBulletEnv-v0": ("gym_locomotion_envs", "HopperBulletEnv", 4, 40, 1.3),
    "Walker2DPyBulletEnv-v0": ("gym_locomotion_envs", "Walker2DBulletEnv", 4, 40, 1.3),
    "HalfCheetahPyBulletEnv-v0": ("gym_locomotion_envs", "HalfCheetahBulletEnv", 4, 40, 1.3),
    "AntPyBulletEnv-v0": ("gym_locomotion_envs", "AntBulletEnv", 3, 60, 1.3),
    "HumanoidPyBulletEnv-v0": ("gym_locomotion_envs", "HumanoidBulletEnv", 3, 40, 1.3),
    "HumanoidFlagrunPyBulletEnv-v0": ("gym_locomotion_envs", "HumanoidFlagrunBulletEnv", 4, 40, 1.3),
    "HumanoidFlagrunHarderPyBulletEnv-v0": ("gym_locomotion_envs", "HumanoidFlagrunHarderBulletEnv", 2, 200, 1.0),
    # "#held": before every step the robot part of the physics state is put back to the reset pose with a
    # small scripted drift, so that the humanoid is still standing at frame 120 / 150 / ... and the reference's
    # cube attack (rs/robot_locomotors.py:251-266) fires; the cube itself keeps flying / colliding freely
    "HumanoidFlagrunHarderPyBulletEnv-v0#held": ("gym_locomotion_envs", "HumanoidFlagrunHarderBulletEnv", 1, 230, 0.2),
    # "#policy": the actions come from the reference's own pretrained MLP for that env (tests/golden/policy_*.npz, extracted from
    # its enjoy_TF_* scripts) evaluated on the reference's observation: long walking episodes -- feet touching down and lifting,
    # joints at their stops, progress, a moving flag -- instead of a few dozen random-action steps before the robot falls
    "HopperPyBulletEnv-v0#policy": ("gym_locomotion_envs", "HopperBulletEnv", 1, 300, 1.0),
    "HalfCheetahPyBulletEnv-v0#policy": ("gym_locomotion_envs", "HalfCheetahBulletEnv", 1, 250, 1.0),
    "AntPyBulletEnv-v0#policy": ("gym_locomotion_envs", "AntBulletEnv", 1, 250, 1.0),
    "HumanoidPyBulletEnv-v0#policy": ("gym_locomotion_envs", "HumanoidBulletEnv", 1, 200, 1.0),
    "HumanoidFlagrunPyBulletEnv-v0#policy": ("gym_locomotion_envs", "HumanoidFlagrunBulletEnv", 1, 320, 1.0),
}


def policy_action(w, ob):
    """The demo scripts' forward pass (enjoy_TF_*: two ReLU layers, linear head; the Humanoid demos shift ob[0] first)."""
    ob = np.asarray(ob, dtype=np.float64) + w["obs_shift"]
    x = np.maximum(ob @ w["dense1_w"] + w["dense1_b"], 0)
    x = np.maximum(x @ w["dense2_w"] + w["dense2_b"], 0)
    return x @ w["final_w"] + w["final_b"]


def main():
    from pybullet_gym_b200.spec import SPECS
    fp.install()
    sys.path.insert(0, REF)
    import importlib
    os.makedirs(OUT, exist_ok=True)
    for case_id, (mod, cls, episodes, max_steps, ascale) in CASES.items():
        env_id, _, variant = case_id.partition("#")
        held = variant == "held"
        policy = np.load(os.path.join(GOLDEN, "policy_%s.npz" % env_id.split("PyBullet")[0])) if variant == "policy" else None
        spec = SPECS[env_id]
        fp.FakeBulletClient.current_spec = spec
        fp.FakeBulletClient.max_contacts = 0
        m = importlib.import_module("pybulletgym.envs." + (mod if "." in mod else "roboschool." + mod))
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):     # the reference prints "WalkerBase::__init__"
            if cls == "HumanoidBulletEnv" and variant:
                # `def __init__(self, robot=Humanoid())` (rs/gym_locomotion_envs.py:147): the default robot is ONE object shared by
                # every HumanoidBulletEnv of the process, already loaded into the first case's client -- a second env of this class
                # gets its own robot through the constructor argument
                env = getattr(m, cls)(robot=m.Humanoid())
            else:
                env = getattr(m, cls)()
        # gym.make() patches reset/step/seed to the most-derived _reset/_step/_seed
        # (gym.envs.registration.patch_deprecated_methods); call those directly
        env._seed(1234)
        rng = np.random.RandomState(99)
        nA = env.action_space.shape[0]
        eps = []
        for ep in range(episodes):
            nlog = len(env.np_random.log)
            obs0 = env._reset()
            draws = env.np_random.log[nlog:]
            flat = [v for d in draws if d[0] == "uniform" for v in d[3]]
            nN = spec.noise_dim
            noise, rec_tape = flat[:nN], None
            rec = {"noise": noise, "obs0": np.asarray(obs0, dtype=np.float64).tolist(), "steps": []}
            s_reset = env._p.orc.get_state().copy()
            obs = obs0
            for t in range(max_steps):
                a = (ascale * rng.uniform(-1, 1, nA)).astype(np.float64)   # |a| > 1 exercises quirk Q3
                if policy is not None:
                    a = np.asarray(policy_action(policy, obs), dtype=np.float64)
                if held:
                    s_now = env._p.orc.get_state()
                    nrob = s_now.size - 13
                    s_now[:nrob] = s_reset[:nrob]
                    s_now[0] += 0.01 * t                  # base x drifts, base velocity 0.6 m/s: the attack leads its target
                    s_now[10] = 0.6
                    env._p.orc.set_state(s_now)
                obs, rew, done, info = env._step(a)
                st = {"a": a.tolist(), "obs": np.asarray(obs, dtype=np.float64).tolist(), "reward": float(rew),
                      "done": bool(done), "rewards": [float(r) for r in env.rewards]}
                if t < 40 or held:
                    st["state"] = env._p.orc.get_state().tolist()
                if hasattr(env.robot, "feet_contact") and hasattr(env.robot, "joints_at_limit"):
                    st["feet_contact"] = [float(f) for f in env.robot.feet_contact]
                    st["joints_at_limit"] = int(env.robot.joints_at_limit)
                    st["body_xyz"] = [float(v) for v in env.robot.body_xyz]
                rec["steps"].append(st)
                if done:
                    break
            # every task-layer draw after the joint noise (flag positions, cube attacks), in order
            rec["tape"] = [v for d in env.np_random.log[nlog:] if d[0] == "uniform" for v in d[3]][nN:]
            eps.append(rec)
        calls = dict(env._p.calls)
        out = {"env_id": env_id, "reference_class": "pybulletgym.envs.roboschool.%s:%s" % (mod, cls),
               "parts": sorted(env.robot.parts.keys()), "ordered_joints": [j.joint_name for j in env.robot.ordered_joints],
               "episodes": eps}
        if held:
            out["held"] = True
        if policy is not None:
            out["policy"] = True
        # changeDynamics calls that switch torsional friction on, with the body they land on (HalfCheetahMuJoCoEnv's reset
        # addresses pybullet body `part.bodyIndex` = 0, the stadium floor: mujoco/robot_locomotors.py:207-210)
        tors = [[b, l, kind, kw] for b, l, kind, kw in getattr(env._p, "dynamics_calls", []) if "spinningFriction" in kw]
        if tors:
            out["torsional_change_dynamics"] = tors
        stem = env_id.split("PyBullet")[0] if "PyBullet" in env_id else env_id.split("Env-")[0]
        path = os.path.join(OUT, "task_%s%s.json" % (stem, "Held" if held else ("Policy" if policy is not None else "")))
        with open(path, "w") as f:
            json.dump(out, f)
        nsteps = sum(len(e["steps"]) for e in eps)
        print("%-40s episodes=%d steps=%d parts=%d calls/step=%.1f -> %s" % (
            case_id, episodes, nsteps, len(out["parts"]),
            sum(v for k, v in calls.items()) / max(1, calls.get("stepSimulation", 1)), os.path.basename(path)))


if __name__ == "__main__":
    main()
