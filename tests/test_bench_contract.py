"""The JSON line `bench.py` prints (the driver depends on it): reference arm on CPU, our arm on a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "e2e")


def _line(args, timeout):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout)
    assert proc.returncode == 0, proc.stderr[-3000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, proc.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _line(["--impl", "reference", "--steps", "3", "--warmup", "3"], 600)
    for k in COMMON:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "overlap_latent_px_per_sec" and d["unit"] == "latent-px/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_our_arm_line():
    """Default run = BASELINE config 3 (the north_star target config), with the parity check and the secondary records."""
    d = _line(["--steps", "20", "--warmup", "3"], 1500)
    for k in COMMON + ("clocks", "gpu_launches", "roofline", "cpu_baseline", "parity_check", "extra"):
        assert k in d, k
    assert d["metric"] == "overlap_latent_px_per_sec" and d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 3
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("cfg3") and "l2" in d["config"]
    assert "no eager launch" in d["config"]["launch"]
    assert d["gpu_launches"] >= d["steps"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1.1
    assert r["traffic"] is None or r["traffic"] >= r["bytes_per_launch"] * 0.9
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                      # host copies are inside the timed region
    assert e["runs"] >= 50 and e["ms_per_run_p90"] >= e["ms_per_run"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and 0 < cb["value"] < d["value"]
    c = d["clocks"]
    assert set(c) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    p = d["parity_check"]
    assert p["ok"] is True and p["max_abs_err_f32"] < 1e-4 and p["ranks"] == 1
    x = d["extra"]
    assert x["cfg2"]["config"]["workload"].startswith("cfg2") and x["cfg2"]["cached_plan"]["job"]["steps"] == 20
    assert x["cfg4_bake"]["unit"] == "views/s" and x["cfg4_bake_reference_modes"]["value"] > 0
    fo = x["feature_overlap"]["sizes"]
    assert set(fo) == {"hw64x64_c320", "hw32x32_c640", "hw16x16_c1280"}
    assert all(0 < v["ms_per_call"] < v["ms_per_call_bucketing_every_call"] and v["rows_gathered"] > 0 for v in fo.values())


def test_workload_string_is_shared_by_both_arms():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.workload_string("cfg3", 96)
    assert s.startswith("cfg3: 96 frames of 1024x1024x4 int32 ids -> 128x128x4 bf16 latents")
