"""CPU: the reference arm of bench.py (--impl reference: the CPU oracle port on the host cores) prints exactly one JSON
line carrying the keys the driver's contract names; rank > 0 under torchrun prints nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--points", "3000",
                           "--width", "96", "--height", "64", "--steps", "1", "--warmup", "0", "--cpu-rows", "32"],
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run({"OMP_NUM_THREADS": "1"})  # what torchrun exports to its workers: the arm must override it
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("fwd+bwd Mpixels/sec") and d["value"] > 0 and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["dtype"] == "f32" and d["data"] == "synthetic" and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "rows" in cb["sample"]
    assert cb["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_prints_one_contract_line():
    """-m gpu: the product arm at a reduced size: one JSON line with value / e2e / roofline / cpu_baseline / clocks."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--points", "20000", "--width", "320", "--height",
                        "192", "--steps", "3", "--warmup", "3", "--cpu-rows", "32"], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["unit"] == "Mpixel/s" and d["value"] > 0
    assert d["scaling"] in ("weak", "strong") and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["gpu_launches"] >= 3 * 13  # every step launches the ~14 kernels of the path
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 320 * 192 * 3 * 4 and e["d2h_bytes_per_step"] == 4
    assert e["h2d_bytes_measured_per_step"] == e["h2d_bytes_per_step"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["kernel"] in ("raster_backward", "raster_forward")
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0 < rf["share_of_step"] <= 1.0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert "workload" in d["config"] and "l2" in d["config"]
    # the 1-GPU point of the scaling workload and the reference's own CUDA kernels beside our path
    sb = d["scale_base"]
    assert sb["views_per_step"] == 64 and sb["value"] > 0 and sb["ms_per_step"] > 0
    rc = d["reference_cuda"]
    assert rc is not None and ("unavailable" in rc or (rc["ms_per_step"] > 0 and rc["ms"]["raster_bwd"] > 0))
    assert d["allreduce_ms"] is None and "traffic_source" in rf
