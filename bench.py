#!/usr/bin/env python
"""Throughput benchmark of the hot path: env-steps/s of batched AntPyBulletEnv-v0 stepping.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  the CPU restatement of the reference path

A "step" is one pbg_step over one batch of `--envs` (default 4096) environments per GPU with
synthetic U(-1,1) actions, auto-reset on; envs are sharded across ranks with no data-path collective
(weak scaling).  One JSON line is printed by rank 0.  See DESIGN.md "Measurement".
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--env", default=ENV_ID)
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_rollout(env_id, steps_per_thread, threads):
    """Oracle (CPU restatement of the reference path) on `threads` host threads; returns env-steps/s.
    ctypes releases the GIL inside orc_rollout, so the threads run in parallel."""
    from oracle.oracle import OracleEnv
    envs = [OracleEnv(env_id, seed=0, env_index=i) for i in range(threads)]
    for e in envs:
        e.rollout(50, action_seed=1)
    t0 = time.perf_counter()
    th = [threading.Thread(target=e.rollout, args=(steps_per_thread, 2)) for e in envs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    return threads * steps_per_thread / dt, dt


def cpu_baseline(env_id, budget_s):
    cores = os.cpu_count() or 1
    rate1, _ = cpu_rollout(env_id, 300, 1)
    n = max(200, int(rate1 * budget_s * 0.7))
    rate, dt = cpu_rollout(env_id, n, cores)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d random-action env steps of %s on each of %d host threads (%.1f s), C oracle in double "
                      "precision, auto-reset; CPU restatement, not pybullet (not installable here)" % (n, env_id, cores, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rate1, _ = cpu_rollout(args.env, 300, 1)
    per_step = max(100, int(rate1 * 1.0))        # ~1 s of CPU work per "step"
    for _ in range(min(args.warmup, 3)):
        cpu_rollout(args.env, per_step // 4, cores)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        cpu_rollout(args.env, per_step, cores)
        total += per_step * cores
        if time.perf_counter() - t0 > 150:
            break
    dt = time.perf_counter() - t0
    value = total / dt
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "%s, %d envs/GPU, U(-1,1) actions, auto-reset, frame_skip 4 x 5 PGS iterations" % (args.env, args.envs),
                       "envs_per_gpu": args.envs,
                       "reference_arm": "same env / action distribution / auto-reset, stepped by the CPU restatement of the reference "
                                        "path (oracle/, double precision) as one independent env per host thread -- pybullet itself "
                                        "is not installable in this image; each step is a bounded sample of ~1 s of CPU work"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d env steps per thread per step x %d threads" % (per_step, cores)},
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pybullet_gym_b200 import _lib
    from pybullet_gym_b200.vector_env import VectorEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    E, K, W = args.envs, args.steps, max(args.warmup, 3)
    from pybullet_gym_b200.sharding import shard
    env_offset, _ = shard(rank, world, E)
    env = VectorEnv(args.env, E, device=dev, seed=0, env_offset=env_offset, auto_reset=True)
    nA, D = env.action_dim, env.obs_dim
    env.reset()
    gen = torch.Generator(device=dev).manual_seed(rank)
    # a ring of pre-generated action batches resident in HBM (synthetic random-policy rollout)
    NACT = 64
    acts = torch.rand(NACT, E, nA, device=dev, generator=gen) * 2 - 1
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)      # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for i in range(W):
        env.step_fast(acts[i % NACT])
    barrier()
    # ---- timed region A: K steps, L2 flushed between steps, each step bracketed by CUDA events
    l0 = env.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        env.step_fast(acts[i % NACT])
        ev[i][1].record()
    barrier()
    launches = env.launch_count() - l0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_flushed = sum(step_ms) * 1e-3
    # ---- timed region B: K steps back to back (state stays L2-resident, as in a real rollout loop)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a0.record()
    for i in range(K):
        env.step_fast(acts[i % NACT])
    a1.record()
    barrier()
    t_resident = a0.elapsed_time(a1) * 1e-3
    clocks = sampler.stop() if sampler else None
    # ---- end to end through the host-buffer C-ABI call: pinned numpy-style buffers, H2D + step + D2H per step
    Ke = min(K, 300)
    h_act = (torch.rand(8, E, nA) * 2 - 1).pin_memory()
    h_obs = torch.empty(E, D).pin_memory()
    h_rew = torch.empty(E).pin_memory()
    h_done = torch.empty(E, dtype=torch.uint8).pin_memory()
    def time_host_path(zero_copy):
        env.set_zero_copy(zero_copy)
        for i in range(5):
            env.step_host(h_act[i % 8], h_obs, h_rew, h_done)
        barrier()
        t0 = time.perf_counter()
        for i in range(Ke):
            env.step_host(h_act[i % 8], h_obs, h_rew, h_done)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        return dt, env.last_host_path()

    t_staged, _ = time_host_path(False)           # H2D copy -> kernel -> D2H copies -> sync
    t_e2e, host_path = time_host_path(True)       # default transport of pbg_step_host (zero-copy for pinned buffers)

    from pybullet_gym_b200 import sharding

    def maxr(x):
        return sharding.max_over_ranks(x, device=dev)

    t_flushed, t_resident, t_e2e, t_staged = maxr(t_flushed), maxr(t_resident), maxr(t_e2e), maxr(t_staged)
    stats = sharding.reduce_stats(env.stats(), device=dev)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        import ctypes as C
        fp32 = C.c_double(0)
        _lib.lib().pbg_measure_fp32_peak(local, C.byref(fp32))
        alg_b, alg_f = ALG.get(args.env, (0, 0))
        ms_kernel = 1e3 * t_flushed / K
        ach_gbs = alg_b * E / (ms_kernel * 1e-3) / 1e9
        ach_tf = alg_f * E / (ms_kernel * 1e-3) / 1e12
        value = world * E * K / t_flushed
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_kernel, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s, %d envs/GPU, U(-1,1) actions, auto-reset, frame_skip 4 x 5 PGS iterations" % (args.env, E),
                       "envs_per_gpu": E, "l2": "flushed between steps (256 MiB memset outside the per-step CUDA-event pair)",
                       "parallelism": "env-sharded x%d, no data-path collective" % world},
            "value_l2_resident": world * E * K / t_resident,
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "traffic": (1.72e6 if (args.env == ENV_ID and E == 4096) else None),
                         "traffic_source": "dram__bytes_read+write per launch, ncu --set full, profiles/r01c_ant_step_kernel_ncu.md",
                         "peak_source": peak_src,
                         "note": "latency/issue-bound FP32 small-matrix kernel: the HBM fraction is reported because the "
                                 "schema asks for it; the meaningful ceiling is fp32 below",
                         "fp32": {"achieved": ach_tf, "peak": fp32.value, "unit": "TFLOP/s",
                                  "frac": (ach_tf / fp32.value) if fp32.value else None,
                                  "peak_source": "FFMA microbenchmark in this run (pbg_measure_fp32_peak)",
                                  "alg_flops_per_env_step": alg_f, "alg_bytes_per_env_step": alg_b}},
            "e2e": {"value": world * E * Ke / t_e2e, "unit": UNIT, "h2d_bytes_per_step": E * nA * 4,
                    "d2h_bytes_per_step": E * (D + 1) * 4 + E, "steps": Ke,
                    "path": "pbg_step_host (%s): pinned host actions -> step kernel -> host obs/reward/done -> sync; "
                            "zero-copy = the kernel reads / writes the mapped pinned buffers over PCIe itself" % host_path,
                    "value_staged_copies": world * E * Ke / t_staged},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "episodes": {"finished": stats["episodes"], "mean_len": stats["length_sum"] / max(1, stats["episodes"]),
                         "mean_return": stats["return_sum"] / max(1, stats["episodes"]),
                         "resets_per_env_step": stats["episodes"] / max(1, stats["steps"])},
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.env, args.cpu_seconds)
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
