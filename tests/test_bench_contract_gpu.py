"""GPU: `python bench.py` prints ONE JSON line that carries every key of the driver's contract (task statement, section 4:
metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling / vs_baseline / dtype / data /
config.workload / clocks / e2e{value, unit, h2d_bytes_per_step, d2h_bytes_per_step} / gpu_launches / roofline{bound,
achieved, peak, unit, frac, traffic}).  Run with --lean (headline + e2e + roofline only: ~25 s) so the test stays short; the
full default line (gpu_eager_fp16, decode, config3, cpu_baseline, two_clips_in_flight) is exercised by the driver itself."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_has_every_contract_key():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--lean", "--steps", "4", "--warmup", "3"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["unit"] == "frames/s" and d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] >= 3
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f16" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and abs(d["value"] - 16 / d["ms_per_step"] * 1e3) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == "frames/s" and 0 < e["value"] <= d["value"] * 1.10      # run-to-run noise is ~1 %
    assert e["h2d_bytes_per_step"] == 16 * 3 * 224 * 224 * 4 + 64 * 8 and e["d2h_bytes_per_step"] == 32000 * 2
    assert d["gpu_launches"] == d["launches_per_step"] * d["steps"] and d["launches_per_step"] > 300
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.3 < r["frac"] < 1.0
    c = d["clocks"]
    assert c["sm_max_mhz"] >= c["sm_mhz"] > 0 and isinstance(c["reasons"], list)   # throttle reasons are reported, not judged here
