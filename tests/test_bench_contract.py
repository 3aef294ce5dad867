"""bench.py's output contract: stdout carries exactly ONE line and that line is the JSON object the driver parses (library
banners -- NCCL prints its version to file descriptor 1 -- must not reach it); the keys the contract names are present."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches")


def run_bench(*flags):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, env=env,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, "stdout must carry the JSON line only, got %d lines" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    for k in BASE_KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "points/s" and d["value"] > 0 and d["dtype"] == "f64" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["gpu_launches"] == 0


@pytest.mark.gpu
def test_gpu_arm_prints_one_json_line_with_roofline_and_launch_count():
    d = run_bench("--steps", "2", "--warmup", "3", "--no-configs", "--no-cpu-baseline", "--no-assembly")
    for k in BASE_KEYS + ("roofline", "clocks", "multi_gpu"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["gpu_launches"] > 0 and d["value"] > 1e9
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert 0.5 < rf["frac"] < 1.3
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] < d["value"]
    mg = d["multi_gpu"]
    assert mg["comm_nranks"] == 1 and mg["checked"] is True
