"""bench.py's JSON-line contract (keys the driver reads), on the CPU reference arm and on the GPU arm."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "env-steps/sec (batched flies)" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
@pytest.mark.parametrize("workload,extra", [("flat", ["--n-flies", "512", "--chunk", "10"]), ("terrain", ["--n-flies", "512", "--chunk", "10"]),
                                            ("vision", ["--n-flies", "32"]), ("olfaction", ["--n-flies", "512"])])
def test_gpu_arm_line(workload, extra):
    d = _run("--workload", workload, "--steps", "20", "--no-cpu", *extra)
    assert (BASE_KEYS | {"clocks", "gpu_launches", "roofline"}) <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] >= 3 and d["dtype"] == "f32" and d["scaling"] == "weak"
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] >= 2 and d["state_finite"] is True
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 1000 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
