"""bench.py prints ONE JSON line per arm with the keys the driver reads.  The reference arm (the CPU restatement of the path on
the host threads) needs no GPU; the product arm is a `-m gpu` test with a tiny workload."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
COMMON = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "cpu_baseline", "e2e"]


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT,
                       timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line(oracle_lib):
    l = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--envs", "256"], 300)
    for k in COMMON + ["impl"]:
        assert k in l, k
    assert l["impl"] == "reference" and l["higher_is_better"] is True and l["vs_baseline"] is None
    assert l["metric"] == "env_steps_per_sec" and l["unit"] == "env-steps/s" and l["value"] > 0
    assert "workload" in l["config"] and "model" not in l["config"]
    cb = l["cpu_baseline"]
    assert cb["kind"] in ("port", "pybullet") and cb["cores"] >= 1 and cb["value"] == l["value"] and cb["sample"]
    assert l["e2e"] == {"value": l["value"], "unit": l["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_product_arm_line():
    l = _run(["--steps", "4", "--warmup", "3", "--envs", "512", "--preroll", "60", "--min-time", "0.001", "--no-configs",
              "--cpu-seconds", "1"], 600)
    for k in COMMON + ["roofline", "clocks", "gpu_launches"]:
        assert k in l, k
    assert "impl" not in l or l["impl"] != "reference"
    assert l["n_gpus"] == 1 and l["steps"] == 4 and l["warmup"] == 3 and l["dtype"] == "f32" and l["data"] == "synthetic"
    assert l["value"] > 0 and l["ms_per_step"] > 0 and l["gpu_launches"] >= l["steps"]
    rf = l["roofline"]
    assert rf["bound"] == "fp32" and rf["unit"] == "TFLOP/s" and 0 < rf["frac"] < 1 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = l["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 512 * 8 * 4 and e["d2h_bytes_per_step"] == 512 * (28 * 4 + 4 + 1)
    cb = l["cpu_baseline"]
    assert cb["kind"] in ("port", "pybullet") and cb["cores"] >= 1 and cb["value"] > 0
    assert set(l["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
