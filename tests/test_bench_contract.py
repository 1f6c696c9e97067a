"""The measurement contract of bench.py, checked without a GPU: the reference arm runs here (it is the reference's own
CPU step), and the committed records of the GPU arm under profiles/ carry every key the contract names."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _last_json_line(text):
    lines = [l for l in text.strip().splitlines() if l.startswith("{")]
    assert lines, text
    return json.loads(lines[-1])


def test_reference_arm_line():
    """`bench.py --impl reference`: one JSON line, impl = reference, the metric / unit / direction of the GPU arm, a
    cpu_baseline that describes this run and an e2e object that repeats the line's own value with no copies."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cavity384",
                          "--steps", "3", "--warmup", "3"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = _last_json_line(out.stdout)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "MLUPS" and d["unit"] == "MLUPS" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["steps"] == 3 and d["warmup"] >= 3 and d["value"] > 0 and d["config"]["workload"] == "cavity384"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    """Under torchrun only rank 0 runs the reference arm; the other ranks exit 0 without output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--workload", "cavity384", "--steps", "3", "--warmup", "3"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_n1*.json")) +
                                        glob.glob(os.path.join(ROOT, "profiles", "r02_scale_n*.json"))))
def test_committed_gpu_records_follow_the_contract(path):
    d = _last_json_line(open(path).read())
    side = d["config"]["workload"] == "datagen256"        # side workload (config 4), device-resident line only: no e2e
    assert BASE_KEYS - ({"e2e"} if side else set()) <= set(d), sorted(BASE_KEYS - set(d))
    assert d["metric"] == "MLUPS" and d["unit"] == "MLUPS" and d["higher_is_better"] is True
    assert d["dtype"] in ("f64", "f32") and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["vs_baseline"] is None
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    if not side:
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    if d["n_gpus"] == 1 and d["config"]["workload"] == "cavity4096":
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        # value and roofline agree: MLUPS x bytes per node-launch / steps per launch = achieved GB/s
        gbs = d["value"] * 1e6 * r["algorithmic_bytes_per_node_per_launch"] / r["steps_per_launch"] / 1e9
        assert abs(gbs - r["achieved"]) / r["achieved"] < 0.01
