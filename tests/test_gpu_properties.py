"""Size-independent properties of the CUDA path at BASELINE.json's full sizes, and the Gym shell.  -m gpu."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

FULL = {"HalfCheetahPyBulletEnv-v0": 4096, "HopperPyBulletEnv-v0": 4096, "Walker2DPyBulletEnv-v0": 4096,
        "AntPyBulletEnv-v0": 16384, "HumanoidPyBulletEnv-v0": 2048, "InvertedPendulumPyBulletEnv-v0": 4096,
        "InvertedDoublePendulumPyBulletEnv-v0": 4096, "ReacherPyBulletEnv-v0": 4096, "HumanoidFlagrunHarderPyBulletEnv-v0": 2048}


def _mk(env_id, n, **kw):
    from pybullet_gym_b200.vector_env import VectorEnv
    return VectorEnv(env_id, n, device="cuda:0", **kw)


@pytest.mark.parametrize("env_id", sorted(FULL))
def test_full_size_rollout_is_deterministic_and_finite(env_id):
    n = FULL[env_id]
    outs = []
    for rep in range(2):
        env = _mk(env_id, n, seed=5)
        env.reset()
        gen = torch.Generator(device="cuda").manual_seed(3)
        acc = torch.zeros(n, device="cuda", dtype=torch.float64)
        for t in range(60):
            a = torch.rand(n, env.action_dim, device="cuda", generator=gen) * 2 - 1
            obs, rew, done, info = env.step(a)
            assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
            acc += rew.double() + obs.double().sum(dim=1)
        outs.append((acc.cpu().numpy(), obs.cpu().numpy().copy(), env.stats()))
        if 2 <= env.spec.kind <= 8:
            assert obs.abs().max() <= 5.0                       # np.clip(..., -5, 5)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])   # bitwise
    assert outs[0][2]["steps"] == 60 * n


@pytest.mark.parametrize("env_id", ["AntPyBulletEnv-v0", "HopperPyBulletEnv-v0", "HumanoidPyBulletEnv-v0"])
def test_envs_are_independent_of_batch_composition(env_id):
    """env i of a big batch == the same env (same RNG stream, same actions) stepped in a small batch."""
    n, k = 1024, 37
    big = _mk(env_id, n, seed=9, auto_reset=True)
    small = _mk(env_id, 64, seed=9, env_offset=k, auto_reset=True)
    big.reset(); small.reset()
    gen = torch.Generator(device="cuda").manual_seed(1)
    for t in range(40):
        a = torch.rand(n, big.action_dim, device="cuda", generator=gen) * 2 - 1
        ob, rb, db, _ = big.step(a)
        os_, rs, ds, _ = small.step(a[k:k + 64].contiguous())
        assert torch.equal(ob[k:k + 64], os_) and torch.equal(rb[k:k + 64], rs) and torch.equal(db[k:k + 64], ds)


@pytest.mark.parametrize("env_id", ["HopperPyBulletEnv-v0", "Walker2DPyBulletEnv-v0", "InvertedPendulumPyBulletEnv-v0"])
def test_auto_reset_semantics(env_id, oracle_lib):
    """After an env finishes, obs is the first observation of a fresh episode (oracle reset with the same
    RNG stream), final_obs the terminal one; episode statistics add up."""
    n = 512
    env = _mk(env_id, n, seed=21, auto_reset=True)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(2)
    episodes = np.ones(n, dtype=np.int64)       # reset() started episode 1 of every env
    checked = 0
    total_done = 0
    for t in range(80):
        a = torch.rand(n, env.action_dim, device="cuda", generator=gen) * 2 - 1
        obs, rew, done, info = env.step(a)
        d = done.cpu().numpy().astype(bool)
        total_done += int(d.sum())
        for i in np.where(d)[0][:3]:
            episodes_i = episodes[i] + 1
            o = oracle_lib.OracleEnv(env_id, seed=21, env_index=int(i))
            for _ in range(episodes_i):
                first = o.reset(floor_in_parts=True)
            assert np.abs(first - obs[i].cpu().numpy()).max() < 1e-5
            checked += 1
        episodes[d] += 1
    st = env.stats()
    assert checked > 0 and st["episodes"] == total_done and st["length_sum"] <= 80 * n


def test_time_limit_truncation():
    from pybullet_gym_b200.spec import SPECS
    import dataclasses
    spec = dataclasses.replace(SPECS["AntPyBulletEnv-v0"], max_episode_steps=7)
    env = _mk("AntPyBulletEnv-v0", 32, spec=spec, auto_reset=True)
    env.reset()
    a = torch.zeros(32, 8, device="cuda")
    for t in range(7):
        obs, rew, done, info = env.step(a)
    assert done.all() and info["truncated"].all()
    obs, rew, done, info = env.step(a)
    assert not done.any()


def test_smoke_every_backed_id():
    """The reference's gym_sanity_check.py pattern: make, reset, one random step, for every backed id."""
    from pybullet_gym_b200.envs import make, registry
    from pybullet_gym_b200.spec import SPECS
    for env_id, spec in SPECS.items():
        env = make(env_id)
        obs = env.reset()
        assert obs.shape == (spec.obs_dim,) and np.isfinite(obs).all()
        obs, r, done, info = env.step(env.action_space.sample())
        assert obs.shape == (spec.obs_dim,) and np.isfinite(obs).all() and np.isfinite(r) and info == {}
        assert isinstance(done, bool)
        env.close()
    assert "AntPyBulletEnv-v0" in registry


def test_gym_shell_attribute_surface():
    from pybullet_gym_b200.envs import make
    env = make("AntPyBulletEnv-v0")
    env.seed(3)
    obs0 = env.reset()
    r = env.unwrapped.robot
    assert len(r.parts) == 22 and "floor" in r.parts and len(r.ordered_joints) == 8 and len(r.feet) == 4
    assert env.unwrapped.scene.dt == pytest.approx(0.0165) and env.unwrapped.scene.frame_skip == 4
    obs, rew, done, info = env.step(np.zeros(8, dtype=np.float32))
    # the host-side views agree with what the kernel reported
    assert abs((r.body_xyz[2] - r.initial_z) - obs[0]) < 1e-5
    assert len(env.unwrapped.rewards) == 5 and abs(sum(env.unwrapped.rewards) - rew) < 1e-5
    assert abs(r.body_rpy[0] - obs[6]) < 1e-5 and abs(r.body_rpy[1] - obs[7]) < 1e-5
    # same seed -> same reset noise -> same first observation
    env2 = make("AntPyBulletEnv-v0"); env2.seed(3)
    assert np.array_equal(env2.reset(), obs0)


def test_flagrun_shell_mirrors_the_device_flag():
    """HumanoidFlagrun's flag lives on the device (flag_reposition, rs/robot_locomotors.py:204-218): after every reset / step the
    shell's walk_target_x/y, flag_timeout and env.potential are the kernel's own values, so env.potential equals
    -walk_target_dist / dt also right after the flag has moved; env.seed(s) keys the flag draws."""
    from pybullet_gym_b200.envs import make
    env = make("HumanoidFlagrunPyBulletEnv-v0"); env.seed(11)
    env.reset()
    u = env.unwrapped
    r = u.robot
    t0 = (r.walk_target_x, r.walk_target_y)
    assert t0 != (1e3, 0.0) and abs(t0[0]) <= 25.0 and abs(t0[1]) <= 25.0          # the first calc_state repositions the flag
    assert u.potential == pytest.approx(-r.walk_target_dist / u.scene.dt, rel=1e-6)
    timeouts, moved = [r.flag_timeout], False
    for t in range(160):                                                              # flag_timeout = 600 / frame_skip = 150 steps
        obs, rew, done, info = env.step(np.zeros(17, dtype=np.float32))
        timeouts.append(r.flag_timeout)
        assert (u.walk_target_x, u.walk_target_y) == (r.walk_target_x, r.walk_target_y)
        assert u.potential == pytest.approx(-r.walk_target_dist / u.scene.dt, rel=1e-5, abs=1e-3)
        if (r.walk_target_x, r.walk_target_y) != t0:
            moved = True
            break
    assert moved and timeouts[1] == timeouts[0] - 1                                   # counts down, then the flag moves
    # same seed -> same flags; another seed -> other flags
    again = make("HumanoidFlagrunPyBulletEnv-v0"); again.seed(11); again.reset()
    other = make("HumanoidFlagrunPyBulletEnv-v0"); other.seed(12); other.reset()
    assert (again.unwrapped.robot.walk_target_x, again.unwrapped.robot.walk_target_y) == t0
    assert (other.unwrapped.robot.walk_target_x, other.unwrapped.robot.walk_target_y) != t0
    harder = make("HumanoidFlagrunHarderPyBulletEnv-v0"); harder.seed(1); harder.reset()
    for t in range(5):
        harder.step(np.zeros(17, dtype=np.float32))
    assert harder.unwrapped.robot.frame == 5 and harder.unwrapped.robot.on_ground_frame_counter == 0


def test_body_part_speed_and_contact_list():
    """BodyPart.speed() (rs/robot_bases.py:243-248) and contact_list() (:280-281) on the shell: the Ant dropped on the floor
    reports floor contacts for its feet in the shape rs/gym_locomotion_envs.py:73 reads (fields [2] = bodyB, [4] = linkB), and
    feet_contact is exactly "the foot has a floor contact"."""
    from pybullet_gym_b200.envs import make
    env = make("AntPyBulletEnv-v0"); env.seed(0); env.reset()
    u = env.unwrapped
    r = u.robot
    prev_z = r.parts["torso"].pose().xyz()[2]
    for t in range(40):
        env.step(np.zeros(8, dtype=np.float32))
        z = r.parts["torso"].pose().xyz()[2]
        vz = r.parts["torso"].speed()[2]
        if 1 < t < 8 and r.feet_contact.sum() == 0:
            assert vz < -0.1 and abs(vz - (z - prev_z) / u.scene.dt) < 0.25     # free fall: end-of-step velocity ~ finite difference of the pose
        prev_z = z
    for i, f in enumerate(r.feet):
        contact_ids = set((x[2], x[4]) for x in f.contact_list())         # the reference's own expression
        assert bool(u.ground_ids & contact_ids) == bool(r.feet_contact[i]), (f.name, contact_ids, r.feet_contact)
    assert r.feet_contact.sum() >= 3                                  # settled on its feet
    assert np.allclose(r.parts["floor"].speed(), 0.0)


def test_step_host_round_trip():
    env = _mk("AntPyBulletEnv-v0", 256, seed=4)
    ref = _mk("AntPyBulletEnv-v0", 256, seed=4)
    env.reset(); ref.reset()
    a = (torch.rand(256, 8) * 2 - 1).pin_memory()
    obs = torch.empty(256, 28).pin_memory(); rew = torch.empty(256).pin_memory(); done = torch.empty(256, dtype=torch.uint8).pin_memory()
    env.step_host(a, obs, rew, done)
    assert env.last_host_path() == "zero-copy"            # pinned buffers: the kernel reads / writes them directly
    o2, r2, d2, _ = ref.step(a.cuda())
    assert torch.equal(obs, o2.cpu()) and torch.equal(rew, r2.cpu()) and torch.equal(done, d2.cpu())
    # staged-copy transport (pageable buffers or zero-copy switched off) gives the same bytes
    env.set_zero_copy(False)
    a2 = (torch.rand(256, 8) * 2 - 1).pin_memory()
    env.step_host(a2, obs, rew, done)
    assert env.last_host_path() == "staged"
    o3, r3, d3, _ = ref.step(a2.cuda())
    assert torch.equal(obs, o3.cpu()) and torch.equal(rew, r3.cpu()) and torch.equal(done, d3.cpu())
    env.set_zero_copy(True)
    a3, obs_p, rew_p, done_p = torch.rand(256, 8) * 2 - 1, torch.empty(256, 28), torch.empty(256), torch.empty(256, dtype=torch.uint8)
    env.step_host(a3, obs_p, rew_p, done_p)               # pageable buffers fall back to staged copies
    assert env.last_host_path() == "staged"
    o4, r4, d4, _ = ref.step(a3.cuda())
    assert torch.equal(obs_p, o4.cpu()) and torch.equal(rew_p, r4.cpu()) and torch.equal(done_p, d4.cpu())


@pytest.mark.parametrize("n", [1, 2, 27, 29, 57])
def test_ragged_batch_sizes_match_the_big_batch(n):
    """Batch sizes that leave warps / CTAs partly empty (1 env, odd counts, one more than a CTA's 28 envs) step
    exactly like the first n envs of a large batch."""
    big = _mk("AntPyBulletEnv-v0", 256, seed=11, auto_reset=True)
    small = _mk("AntPyBulletEnv-v0", n, seed=11, auto_reset=True)
    big.reset(); small.reset()
    gen = torch.Generator(device="cuda").manual_seed(5)
    for t in range(25):
        a = torch.rand(256, 8, device="cuda", generator=gen) * 2 - 1
        ob, rb, db, _ = big.step(a)
        os_, rs, ds, _ = small.step(a[:n].contiguous())
        assert torch.equal(ob[:n], os_) and torch.equal(rb[:n], rs) and torch.equal(db[:n], ds)
    assert small.stats()["steps"] == 25 * n


def test_nonfinite_action_ends_only_that_env():
    """SURVEY 8b error contract: a NaN in one env ends that env's episode (rs/gym_locomotion_envs.py:63-65), is counted in
    pbg_episode_stats.nonfinite, the env restarts clean, and its neighbours (same warp, same CTA) are untouched."""
    n = 64
    env = _mk("AntPyBulletEnv-v0", n, seed=2, auto_reset=True)
    ref = _mk("AntPyBulletEnv-v0", n, seed=2, auto_reset=True)
    env.reset(); ref.reset()
    gen = torch.Generator(device="cuda").manual_seed(8)
    for t in range(6):
        a = torch.rand(n, 8, device="cuda", generator=gen) * 2 - 1
        b = a.clone()
        if t == 3:
            b[5, 2] = float("nan")
        ob, rb, db, _ = env.step(b)
        orf, rr, dr, _ = ref.step(a)
        others = torch.ones(n, dtype=torch.bool, device="cuda"); others[5] = False
        if t < 3:
            assert torch.equal(ob, orf)
        else:
            assert torch.equal(ob[others], orf[others]) and torch.equal(rb[others], rr[others])
        if t == 3:
            assert db[5].item() == 1 and torch.isfinite(ob[5]).all()       # ended, and the returned obs is the fresh episode's
        assert torch.isfinite(ob).all()
    st = env.stats()
    assert st["nonfinite"] == 1 and st["episodes"] >= 1


def test_error_contract_on_a_live_handle():
    """Bad calls on a live handle return an error (raised as PbgError / ValueError by the host mirror) and leave the handle
    usable: the next good step is bit-identical to an undisturbed twin's."""
    import ctypes
    from pybullet_gym_b200 import _lib
    env = _mk("HopperPyBulletEnv-v0", 64, seed=3)
    twin = _mk("HopperPyBulletEnv-v0", 64, seed=3)
    env.reset(); twin.reset()
    L, h = env._L, env._h
    null = ctypes.c_void_p()
    with pytest.raises(ValueError):
        env.step(torch.zeros(64, 4, device="cuda"))                     # wrong action width
    with pytest.raises(ValueError):
        env.step(torch.zeros(63, 3, device="cuda"))                     # wrong batch
    with pytest.raises(ValueError):
        env.reset(joint_noise=torch.zeros(64, 2, device="cuda"))
    assert L.pbg_step(h, null, null, null, null, null, null, null, null) == -1 and b"actions_dev" in L.pbg_last_error(h)
    assert L.pbg_step_host(h, null, null, null, null) == -1
    with pytest.raises(_lib.PbgError, match="no policy set"):
        env.rollout_policy(4)
    big = np.zeros((15, 4096), np.float32)
    with pytest.raises(_lib.PbgError, match="do not fit"):
        env.set_policy(big, np.zeros(4096, np.float32), np.zeros((4096, 64), np.float32), np.zeros(64, np.float32),
                       np.zeros((64, 3), np.float32), np.zeros(3, np.float32))
    with pytest.raises(_lib.PbgError):
        env.rollout_policy(0)
    a = torch.rand(64, 3, device="cuda") * 2 - 1
    o1, r1, d1, _ = env.step(a)
    o2, r2, d2, _ = twin.step(a)
    assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(d1, d2)


@pytest.mark.parametrize("env_id", ["AntPyBulletEnv-v0", "HumanoidFlagrunHarderPyBulletEnv-v0", "InvertedPendulumPyBulletEnv-v0"])
def test_snapshot_restore_resumes_bit_identically(env_id):
    """pbg_snapshot / pbg_restore (the reference's saveState / restoreState, rs/gym_pendulum_envs.py:20-27): the continuation
    after a restore -- auto-resets and their RNG draws, targets, the thrown cube included -- repeats the original bit for bit, in
    the same handle and in a fresh one, from a device blob and from a pinned-host blob; a wrong blob is refused."""
    from pybullet_gym_b200 import _lib
    n = 256
    env = _mk(env_id, n, seed=13, auto_reset=True)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(4)
    # FlagrunHarder episodes end after 170 frames on the ground at the earliest: put the window where they do
    pre, post = (165, 60) if "Harder" in env_id else (30, 60)
    acts = [torch.rand(n, env.action_dim, device="cuda", generator=gen) * 2 - 1 for _ in range(pre + post)]
    for a in acts[:pre]:
        env.step(a)
    blob_d, blob_h = env.snapshot(), env.snapshot(pinned_host=True)
    stats0 = env.stats()
    first = [tuple(x.clone() for x in env.step(a)[:3]) for a in acts[pre:]]
    assert sum(int(d.sum()) for _, _, d in first) > 0 or env_id.startswith("Ant")      # resets happen inside the window
    stats1 = env.stats()
    fresh = _mk(env_id, n, seed=13, auto_reset=True)
    for target, blob in ((env, blob_d), (fresh, blob_h)):
        target.restore(blob)
        assert target.stats() == stats0
        for a, (o, r, d) in zip(acts[pre:], first):
            o2, r2, d2, _ = target.step(a)
            assert torch.equal(o, o2) and torch.equal(r, r2) and torch.equal(d, d2)
        assert target.stats() == stats1
    other = _mk(env_id, n, seed=14)
    with pytest.raises(_lib.PbgError, match="seed"):
        other.restore(blob_d)
    with pytest.raises(ValueError):
        _mk(env_id, n // 2, seed=13).restore(blob_d)


@pytest.mark.parametrize("env_id,n", [("AntPyBulletEnv-v0", 16384), ("HumanoidPyBulletEnv-v0", 2048), ("Walker2DPyBulletEnv-v0", 4096)])
def test_state_invariants_and_get_set_idempotence(env_id, n):
    """Full-size physical invariants after 150 random steps (unit base quaternion, nothing below the floor, finite state), and
    set_state(get_state()) is the identity on the canonical state while leaving the next step within fp32 rounding of an
    undisturbed twin's (warm-start impulses are not part of the canonical state, so not bitwise)."""
    env = _mk(env_id, n, seed=6, auto_reset=True)
    twin = _mk(env_id, n, seed=6, auto_reset=True)
    env.reset(); twin.reset()
    gen = torch.Generator(device="cuda").manual_seed(12)
    for t in range(150):
        a = torch.rand(n, env.action_dim, device="cuda", generator=gen) * 2 - 1
        env.step(a); twin.step(a)
    st = env.get_state()
    assert torch.isfinite(st).all()
    if env_id.startswith(("Ant", "Humanoid")):                          # floating base: [pos3 quat4 omega3 vel3 | q | qd]
        qn = st[:, 3:7].norm(dim=1)
        assert (qn - 1).abs().max().item() < 1e-5
        assert st[:, 2].min().item() > 0.0                              # base COM never under the floor
    env.set_state(st)
    assert torch.equal(env.get_state(), st)
    a = torch.rand(n, env.action_dim, device="cuda", generator=gen) * 2 - 1
    o1, r1, d1, _ = env.step(a)
    o2, r2, d2, _ = twin.step(a)
    same = (d1 == d2)
    assert same.float().mean().item() > 0.999
    err = (o1[same] - o2[same]).abs().flatten()
    assert err.median().item() < 1e-3 and torch.quantile(err[:1000000].float(), 0.99).item() < 5e-2


@pytest.mark.parametrize("env_id", ["InvertedPendulumPyBulletEnv-v0", "AntPyBulletEnv-v0"])
def test_steps_capture_into_a_cuda_graph(env_id):
    """pbg_step enqueues exactly one kernel on the caller's stream (no allocation, no synchronisation), so a block of steps can
    be captured into a CUDA graph -- what makes the launch-bound pendulum envs fast (tools/graph_probe.py).  The replay is
    bit-identical to eager stepping and the device-side statistics (env steps, episodes) keep counting."""
    n, K = 512, 24
    eager = _mk(env_id, n, seed=1, auto_reset=True)
    graphed = _mk(env_id, n, seed=1, auto_reset=True)
    eager.reset(); graphed.reset()
    acts = torch.rand(K, n, eager.action_dim, device="cuda") * 2 - 1
    log_o = torch.empty(K, n, eager.obs_dim, device="cuda"); log_r = torch.empty(K, n, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        snap = graphed.snapshot()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for k in range(K):
                o, r, d = graphed.step_fast(acts[k])
                log_o[k].copy_(o); log_r[k].copy_(r)
        graphed.restore(snap)
        graphed.stats(reset=True)
        for rep in range(2):
            g.replay()
            torch.cuda.synchronize()
            for k in range(K):
                o, r, d = eager.step_fast(acts[k])
                assert torch.equal(o, log_o[k]) and torch.equal(r, log_r[k]), (rep, k)
    st_g, st_e = graphed.stats(), eager.stats()
    assert st_g["steps"] == 2 * K * n
    assert st_g["episodes"] == st_e["episodes"] and st_g["length_sum"] == st_e["length_sum"]


def test_cross_stream_ordering_is_the_library_s_job():
    """pbg_step_host runs on a private stream and pbg_stats blocks on the legacy stream; neither used to be ordered
    after a reset / step the caller had enqueued on ANOTHER stream.  The handle now remembers the stream of its last
    stream-ordered call and waits for it: reset on a busy side stream followed directly by step_host / stats gives the same
    result as the fully synchronised sequence."""
    n = 512
    env, twin = _mk("AntPyBulletEnv-v0", n, seed=21), _mk("AntPyBulletEnv-v0", n, seed=21)
    a = (torch.rand(n, env.action_dim) * 2 - 1).pin_memory()
    bufs = [(torch.empty(n, env.obs_dim).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory())
            for _ in range(2)]
    side = torch.cuda.Stream()
    busy = torch.empty(64 * 1024 * 1024, device="cuda")
    with torch.cuda.stream(side):
        for _ in range(20):
            busy.normal_()                      # keeps the side stream busy for a few ms before the reset runs
        env.reset()
        env.step_fast(torch.zeros(n, env.action_dim, device="cuda"))
    env.step_host(a, *bufs[0])                  # no synchronisation by the caller
    st = env.stats()
    twin.reset(); torch.cuda.synchronize()
    twin.step_fast(torch.zeros(n, twin.action_dim, device="cuda")); torch.cuda.synchronize()
    twin.step_host(a, *bufs[1])
    assert torch.equal(bufs[0][0], bufs[1][0]) and torch.equal(bufs[0][1], bufs[1][1]) and torch.equal(bufs[0][2], bufs[1][2])
    assert st["steps"] == 2 * n == twin.stats()["steps"]


def test_raw_pointer_paths_check_their_arguments():
    """step_fast / step_host take raw data pointers; a wrong device / dtype / shape / stride or a step before the
    first reset must raise instead of reading garbage."""
    env = _mk("HopperPyBulletEnv-v0", 64)
    good = torch.zeros(64, env.action_dim, device="cuda")
    with pytest.raises(RuntimeError):
        env.step_fast(good)                                     # before reset()
    from pybullet_gym_b200 import _lib
    with pytest.raises(_lib.PbgError):
        env.step(good)                                          # the library refuses as well (PBG_ERR_INVALID)
    env.reset()
    env.step_fast(good)
    for bad in (good.cpu(), good.double(), torch.zeros(63, env.action_dim, device="cuda"), torch.zeros(64, 2 * env.action_dim, device="cuda")[:, ::2]):
        with pytest.raises(ValueError):
            env.step_fast(bad)
    with pytest.raises(ValueError):
        env.step_host(good, torch.empty(64, env.obs_dim).pin_memory(), torch.empty(64).pin_memory(), torch.empty(64, dtype=torch.uint8).pin_memory())


def test_restore_brings_back_the_step_outputs():
    """observe() is the task half of a step, not a way to get the observation back; snapshot() therefore carries the
    output buffers and restore() puts them back."""
    env = _mk("HumanoidFlagrunPyBulletEnv-v0", 128, seed=3, auto_reset=True)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(8)
    for _ in range(12):
        obs, rew, done, info = env.step(torch.rand(128, env.action_dim, device="cuda", generator=gen) * 2 - 1)
    keep = (obs.clone(), rew.clone(), done.clone(), info["reward_terms"].clone())
    blob = env.snapshot()
    for _ in range(5):
        env.step(torch.rand(128, env.action_dim, device="cuda", generator=gen) * 2 - 1)
    assert not torch.equal(env.obs, keep[0])
    env.restore(blob)
    assert torch.equal(env.obs, keep[0]) and torch.equal(env.reward, keep[1]) and torch.equal(env.done, keep[2]) and torch.equal(env.terms, keep[3])
