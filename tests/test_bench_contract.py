"""The JSON line `bench.py` prints: one line, the keys the driver reads (task contract), for both arms.  The reference arm runs on
the CPU (a small image so that the CPU suite stays short); the product arm needs a B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
          "data", "config", "e2e")


def _run(*args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines  # ONE JSON line on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--size", "64")
    for k in COMMON + ("impl", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "optimizer_frames_per_sec_fwd_bwd_512x512_MANO" and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_product_arm_line():
    d = _run("--steps", "5", "--warmup", "3", "--no-standin")
    for k in COMMON + ("roofline", "cpu_baseline", "clocks", "gpu_launches", "e2e_float32_host", "e2e_resident_epochs", "counters"):
        assert k in d, k
    assert "impl" not in d or d["impl"] != "reference"
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["value"] > 1000
    assert abs(d["value"] - 64 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and r["peak"] > 1000 and r["achieved"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and (r["traffic"] is None or r["traffic"] > 0)
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 64 * 512 * 512 * 4 and e["d2h_bytes_per_step"] > 0
    assert d["e2e_float32_host"]["h2d_bytes_per_step"] == 64 * 512 * 512 * 20
    assert d["gpu_launches"] > 0 and d["counters"]["sil_k_overflow_pixels"] == 0 and d["counters"]["bin_overflow"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
