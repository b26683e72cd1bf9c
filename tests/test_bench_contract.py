"""The JSON line of `bench.py --impl reference` (the CPU arm the driver runs first): keys and meanings of the contract."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/sec on Sponza primary+random" and line["unit"] == "Mrays/s"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] >= 3 and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None and line["gpu_launches"] == 0
    assert abs(line["value"] - 2 * 1048576 / line["ms_per_step"] / 1e3) / line["value"] < 0.02      # rays of one step / its time
    assert line["config"]["workload"].startswith("bench_traversal: Sponza BVH8/Tri4")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "passes" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_0_only():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest


@pytest.mark.gpu
def test_b200_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "3", "--warmup", "3", "--no-path-trace", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks", "per_set", "variants"):
        assert key in line, key
    assert "impl" not in line and line["metric"] == "Mrays/sec on Sponza primary+random" and line["dtype"] == "f32"
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["gpu_launches"] == 2 * line["steps"]                       # two traversal launches per step, nothing else of ours
    assert abs(line["value"] - 2 * 1048576 / line["ms_per_step"] / 1e3) / line["value"] < 0.01
    assert line["value"] > 500 and 0 < line["e2e"]["value"] < line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 2 * 1048576 * 32 and line["e2e"]["d2h_bytes_per_step"] == 2 * 1048576 * 16
    assert line["e2e"]["results_match_device_path"] is True
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3 and rf["traffic"] > 0
    assert line["per_set"]["primary"]["hits"] == 1026430 and line["per_set"]["random"]["hits"] == 959359
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
