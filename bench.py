#!/usr/bin/env python
"""Throughput benchmark of the hot path: env-steps/s of batched AntPyBulletEnv-v0 stepping.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the CPU restatement of the reference path

A "step" is one pbg_step over one batch of `--envs` (default 4096) environments per GPU with
synthetic U(-1,1) actions, auto-reset on; envs are sharded across ranks with no data-path collective
(weak scaling).  The batch is first pre-rolled (untimed) to its stationary mix of episode phases, so the
number does not depend on --steps / --warmup; the K-step block is repeated until the timed region is
>= 0.3 s.  One JSON line is printed by rank 0; its "configs" list carries short runs of BASELINE.json's
other single-GPU configurations.  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
ENV_ID = "AntPyBulletEnv-v0"
# SURVEY.md 8(d) / BASELINE.md section 3: algorithmic bytes and flops per env-step
ALG = {
    "AntPyBulletEnv-v0": (461, 1.04e5), "HalfCheetahPyBulletEnv-v0": (373, 5.6e4), "HopperPyBulletEnv-v0": (229, 2.7e4),
    "Walker2DPyBulletEnv-v0": (325, 4.9e4), "HumanoidPyBulletEnv-v0": (849, 3.0e5),
    "InvertedPendulumPyBulletEnv-v0": (109, 1.5e3), "HumanoidFlagrunHarderPyBulletEnv-v0": (849, 3.0e5),
    "HumanoidFlagrunPyBulletEnv-v0": (849, 3.0e5),
}


# the other single-GPU configurations BASELINE.json names (C2, C3, C4's upper size, C5) and the Humanoid at the metric's
# 4096 envs/GPU: short runs reported under "configs"
EXTRA_CONFIGS = [("HalfCheetahPyBulletEnv-v0", 4096), ("HopperPyBulletEnv-v0", 4096), ("Walker2DPyBulletEnv-v0", 4096),
                 ("AntPyBulletEnv-v0", 16384), ("HumanoidFlagrunHarderPyBulletEnv-v0", 2048), ("HumanoidPyBulletEnv-v0", 4096)]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--env", default=ENV_ID)
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--preroll", type=int, default=1000,
                    help="untimed env steps before the warm-up that bring the batch to its stationary episode-phase mix")
    ap.add_argument("--min-time", type=float, default=0.3, help="the K-step block is repeated until the timed region is this long (s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of BASELINE.json's other configurations")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    return ap.parse_args()


def config_of(env_id, envs):
    """The `config` object, identical for both arms (the driver compares them)."""
    return {"workload": "%s, %d envs/GPU, U(-1,1) actions, auto-reset, frame_skip 4 x 5 PGS iterations" % (env_id, envs),
            "envs_per_gpu": envs,
            "l2": "GPU arm: flushed between steps (256 MiB memset outside the per-step CUDA-event pair); CPU arm: not applicable",
            "parallelism": "env-sharded over the ranks (GPU arm) / host threads (CPU arm), no data-path collective",
            "phase": "stationary: measured after pre-rolling the batch through whole episodes with staggered resets"}


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_rollout(env_id, steps_per_thread, threads, envs=None):
    """Oracle (CPU restatement of the reference path) on `threads` host threads; returns env-steps/s.
    ctypes releases the GIL inside orc_rollout, so the threads run in parallel."""
    from oracle.oracle import OracleEnv
    envs = envs or [OracleEnv(env_id, seed=0, env_index=i) for i in range(threads)]
    t0 = time.perf_counter()
    th = [threading.Thread(target=e.rollout, args=(steps_per_thread, 2)) for e in envs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return threads * steps_per_thread / dt, dt, envs


def _settled_oracles(env_id, threads):
    """One oracle env per thread, rolled 400 random-action steps first: like the GPU arm, the sample is taken in the
    stationary phase of the rollout (robots on the ground), not during the first airborne frames after the very first reset."""
    from oracle.oracle import OracleEnv
    envs = [OracleEnv(env_id, seed=0, env_index=i) for i in range(threads)]
    cpu_rollout(env_id, 400, threads, envs)
    return envs


def real_reference_available():
    """Pin day: the moment the reference's own stack (pybullet + gym + pybulletgym) imports, the CPU arm is the real thing."""
    try:
        import pybullet
        import gym  # noqa: F401
        return hasattr(pybullet, "connect") and not getattr(pybullet, "IS_PBG_STUB", False)
    except Exception:
        return False


def real_reference_rate(env_id, seconds, procs):
    """The UNMODIFIED reference (gym.make(env_id) from pybulletgym, pybullet DIRECT client) as one process per host core
    (BASELINE.md section 4.1): env-steps/s summed over the processes, each running `seconds` of random-action steps."""
    code = ("import sys,time,json\n"
            "sys.path[:0]=[%r,%r]\n"
            "import gym, numpy as np\nimport pybulletgym\n"
            "env=gym.make(%r); env.reset(); n=0; t0=time.perf_counter()\n"
            "while time.perf_counter()-t0<%f:\n"
            "    o,r,d,_=env.step(env.action_space.sample()); n+=1\n"
            "    if d: env.reset()\n"
            "print(json.dumps({'n':n,'dt':time.perf_counter()-t0}))\n"
            % (os.path.join(ROOT, "baseline", "_ref"), "/root/reference", env_id, seconds))
    ps = [subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for _ in range(procs)]
    rate = 0.0
    for p in ps:
        out = p.communicate()[0].strip().splitlines()
        r = json.loads(out[-1])
        rate += r["n"] / r["dt"]
    return rate


def cpu_baseline(env_id, budget_s):
    cores = os.cpu_count() or 1
    if real_reference_available():
        rate = real_reference_rate(env_id, budget_s, cores)
        return {"value": rate, "unit": UNIT, "cores": cores, "kind": "pybullet",
                "sample": "%.0f s of random-action steps of gym.make(%s) (unmodified reference on pybullet) in each of %d processes"
                          % (budget_s, env_id, cores)}
    envs = _settled_oracles(env_id, cores)
    rate1, _, _ = cpu_rollout(env_id, 300, cores, envs)
    n = max(200, int(rate1 / cores * budget_s * 0.8))
    rate, dt, _ = cpu_rollout(env_id, n, cores, envs)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d random-action env steps of %s on each of %d host threads (%.1f s) after 700 settling steps, C oracle in "
                      "double precision, auto-reset; CPU restatement, not pybullet (not installable here)" % (n, env_id, cores, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    real = real_reference_available()
    t0 = time.perf_counter()
    total = 0.0
    if real:
        per = 1.0                                     # seconds of stepping per "step"
        for _ in range(min(args.warmup, 1)):
            real_reference_rate(args.env, 0.5, cores)
        t0 = time.perf_counter()
        tsum = 0.0
        for _ in range(args.steps):
            total += real_reference_rate(args.env, per, cores) * per
            tsum += per
            if time.perf_counter() - t0 > 150:
                break
        dt = tsum
        sample = "%.1f s of gym.make(%s).step(random action) in each of %d processes per step" % (per, args.env, cores)
        kind = "pybullet"
    else:
        envs = _settled_oracles(args.env, cores)
        rate1, _, _ = cpu_rollout(args.env, 300, cores, envs)
        per_step = max(100, int(rate1 / cores * 1.0))        # ~1 s of CPU work per "step"
        for _ in range(min(args.warmup, 3)):
            cpu_rollout(args.env, per_step // 4, cores, envs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_rollout(args.env, per_step, cores, envs)
            total += per_step * cores
            if time.perf_counter() - t0 > 150:
                break
        dt = time.perf_counter() - t0
        sample = "%d env steps per thread per step x %d threads, after 700 settling steps" % (per_step, cores)
        kind = "port"
    value = total / dt
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": config_of(args.env, args.envs),
            "reference_arm": ("the unmodified reference (pybulletgym on pybullet), one process per host core" if real else
                              "same env / action distribution / auto-reset, stepped by the CPU restatement of the reference "
                              "path (oracle/, double precision) as one independent env per host thread -- pybullet itself "
                              "is not installable in this image; each step is a bounded sample of ~1 s of CPU work"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], None, set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def load_traffic(env_id, envs):
    """DRAM bytes per launch of the step kernel from the committed ncu captures (profiles/traffic.json), or None."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tab.get("%s@%d" % (env_id, envs))
        return (ent["dram_bytes_per_launch"], ent["source"]) if ent else (None, None)
    except Exception:
        return None, None


class Workload:
    """One env kind x E envs on this rank's GPU, pre-rolled to its stationary episode-phase mix."""

    def __init__(self, env_id, E, dev, rank, world, preroll):
        import torch
        from pybullet_gym_b200.sharding import shard
        from pybullet_gym_b200.vector_env import VectorEnv
        self.torch, self.env_id, self.E, self.dev = torch, env_id, E, dev
        env_offset, _ = shard(rank, world, E)
        self.env = env = VectorEnv(env_id, E, device=dev, seed=0, env_offset=env_offset, auto_reset=True)
        self.nA, self.D = env.action_dim, env.obs_dim
        env.reset()
        gen = torch.Generator(device=dev).manual_seed(rank)
        # a ring of pre-generated action batches resident in HBM (synthetic random-policy rollout)
        self.NACT = 64
        self.acts = torch.rand(self.NACT, E, self.nA, device=dev, generator=gen) * 2 - 1
        # Pre-roll (untimed): the throughput depends on where in their episodes the envs are (robots dropping from the spawn
        # height have fewer contacts than walking / lying ones).  Step through `preroll` steps and reset a random 5 % of the
        # batch every 50 steps, so that episode ages are spread over the whole TimeLimit instead of moving as one cohort.
        for i in range(preroll):
            env.step_fast(self.acts[i % self.NACT])
            if i % 50 == 49 and i + 50 < preroll:
                mask = (torch.rand(E, device=dev, generator=gen) < 0.05).to(torch.uint8)
                env.reset(mask=mask)
        env.stats(reset=True)
        self.i = preroll

    def step(self):
        self.env.step_fast(self.acts[self.i % self.NACT])
        self.i += 1


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pybullet_gym_b200 import _lib, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the CPU baseline is an N = 1 figure; it runs BEFORE any GPU work so that no rank spins in a barrier beside it
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.env, args.cpu_seconds)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    K, W = args.steps, max(args.warmup, 3)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)      # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        return sharding.max_over_ranks(x, device=dev)

    def timed_flushed(wl, K, min_time):
        """K steps, L2 flushed before each, each step bracketed by its own CUDA-event pair on the launching stream; the block
        is repeated until the summed step time reaches min_time.  Returns (mean seconds per step, repeats, launches)."""
        tot, reps, launches = 0.0, 0, 0
        while True:
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            l0 = wl.env.launch_count()
            barrier()
            for i in range(K):
                flush.zero_()
                ev[i][0].record()
                wl.step()
                ev[i][1].record()
            barrier()
            launches += wl.env.launch_count() - l0
            tot += sum(a.elapsed_time(b) for a, b in ev) * 1e-3
            reps += 1
            # every rank must take the same decision: decide on the slowest rank's clock
            if maxr(tot) >= min_time or reps >= 200:
                return tot / (reps * K), reps, launches

    def time_host_path(wl, zero_copy, Ke, bufs):
        h_act, h_obs, h_rew, h_done = bufs
        wl.env.set_zero_copy(zero_copy)
        for i in range(5):
            wl.env.step_host(h_act[i % 8], h_obs, h_rew, h_done)
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            wl.env.step_host(h_act[i % 8], h_obs, h_rew, h_done)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        return dt, wl.env.last_host_path()

    def host_bufs(wl):
        return ((torch.rand(8, wl.E, wl.nA) * 2 - 1).pin_memory(), torch.empty(wl.E, wl.D).pin_memory(),
                torch.empty(wl.E).pin_memory(), torch.empty(wl.E, dtype=torch.uint8).pin_memory())

    import ctypes as C
    fp32 = C.c_double(0)
    _lib.lib().pbg_measure_fp32_peak(local, C.byref(fp32))

    def fp32_roof(env_id, E, sec_per_step):
        alg_b, alg_f = ALG.get(env_id, (0, 0))
        ach_tf = alg_f * E / sec_per_step / 1e12
        return alg_b, alg_f, ach_tf, (ach_tf / fp32.value) if fp32.value else None

    # ---------------- headline workload
    wl = Workload(args.env, args.envs, dev, rank, world, args.preroll)
    E = wl.E
    for _ in range(W):
        wl.step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    # timed region A: the metric
    sec_step, reps, launches = timed_flushed(wl, K, args.min_time)
    # timed region B: the same K steps back to back (state stays L2-resident, as in a real rollout loop)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a0.record()
    for i in range(K * reps):
        wl.step()
    a1.record()
    barrier()
    t_resident = a0.elapsed_time(a1) * 1e-3 / (K * reps)
    clocks = sampler.stop() if sampler else None
    stats = sharding.reduce_stats(wl.env.stats(), device=dev)
    # end to end through the host-buffer C-ABI call: pinned host buffers, H2D + step + D2H (or zero-copy PCIe) + sync per step
    Ke = max(min(K, 300), min(300, int(0.3 / sec_step) + 1))
    bufs = host_bufs(wl)
    t_staged, _ = time_host_path(wl, False, Ke, bufs)           # H2D copy -> kernel -> D2H copies -> sync
    t_e2e, host_path = time_host_path(wl, True, Ke, bufs)       # default transport of pbg_step_host (zero-copy for pinned buffers)
    sec_step, t_resident, t_e2e, t_staged = maxr(sec_step), maxr(t_resident), maxr(t_e2e), maxr(t_staged)

    # ---------------- BASELINE.json's other single-GPU configurations, short runs (every rank runs its own shard)
    wl_dims = (wl.nA, wl.D)
    extra = []
    if not args.no_configs:
        del wl, bufs
        for env_id, En in EXTRA_CONFIGS:
            w2 = Workload(env_id, En, dev, rank, world, args.preroll)
            for _ in range(W):
                w2.step()
            s2, r2, _ = timed_flushed(w2, min(K, 100), 0.25)
            st2 = sharding.reduce_stats(w2.env.stats(), device=dev)
            b2 = host_bufs(w2)
            Ke2 = min(300, int(0.2 / s2) + 20)
            te2, _ = time_host_path(w2, True, Ke2, b2)
            s2, te2 = maxr(s2), maxr(te2)
            alg_b, alg_f, ach_tf, frac = fp32_roof(env_id, En, s2)
            extra.append({"workload": config_of(env_id, En)["workload"], "envs_per_gpu": En, "value": world * En / s2, "unit": UNIT,
                          "ms_per_step": 1e3 * s2, "timed_steps": min(K, 100) * r2, "e2e": world * En * Ke2 / te2,
                          "fp32_frac": frac, "hbm_gbs": alg_b * En / s2 / 1e9,
                          "resets_per_env_step": st2["episodes"] / max(1, st2["steps"]),
                          "solver_budget_binds_per_env_step": st2.get("contact_overflow", 0) / max(1, st2["steps"]),
                          "mean_episode_len": st2["length_sum"] / max(1, st2["episodes"])})
            del w2, b2

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        alg_b, alg_f, ach_tf, frac = fp32_roof(args.env, E, sec_step)
        ach_gbs = alg_b * E / sec_step / 1e9
        traffic, traffic_src = load_traffic(args.env, E)
        line = {
            "metric": METRIC, "value": world * E / sec_step, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * sec_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_of(args.env, E),
            "timed": {"block_steps": K, "block_repeats": reps, "timed_steps": K * reps, "preroll_steps": args.preroll,
                      "note": "the K-step block is repeated until the timed region is >= %.2f s; ms_per_step is the mean over all of them" % args.min_time},
            "value_l2_resident": world * E / t_resident,
            "roofline": {"bound": "fp32", "achieved": ach_tf, "peak": fp32.value, "unit": "TFLOP/s", "frac": frac,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "FFMA microbenchmark in this run (pbg_measure_fp32_peak); MEASURED_PEAKS.json carries no FP32 "
                                        "CUDA-core figure; nominal 74.5",
                         "alg_flops_per_env_step": alg_f, "alg_bytes_per_env_step": alg_b,
                         "note": "issue/latency-bound FP32 small-matrix kernel (no tensor-core-shaped work, SURVEY 8d); achieved = "
                                 "alg_flops_per_env_step x envs / mean step time",
                         "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "peak_source": peak_src}},
            "e2e": {"value": world * E * Ke / t_e2e, "unit": UNIT, "h2d_bytes_per_step": E * wl_dims[0] * 4,
                    "d2h_bytes_per_step": E * (wl_dims[1] + 1) * 4 + E, "steps": Ke,
                    "path": "pbg_step_host (%s): pinned host actions -> step kernel -> host obs/reward/done -> sync; "
                            "zero-copy = the kernel reads / writes the mapped pinned buffers over PCIe itself" % host_path,
                    "value_staged_copies": world * E * Ke / t_staged},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "episodes": {"finished": stats["episodes"], "mean_len": stats["length_sum"] / max(1, stats["episodes"]),
                         "mean_return": stats["return_sum"] / max(1, stats["episodes"]),
                         "resets_per_env_step": stats["episodes"] / max(1, stats["steps"]),
                         "solver_budget_binds_per_env_step": stats.get("contact_overflow", 0) / max(1, stats["steps"])},
            "configs": extra,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
