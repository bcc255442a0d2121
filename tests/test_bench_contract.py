"""bench.py prints ONE JSON line with the keys the driver reads.  CPU: the reference arm (the unmodified reference
on the host cores, oracle/_ref).  GPU: our arm at a few steps."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref not built here")
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "ctxt-mul+decrypt output blocks/s" and d["unit"] == "blocks/s" and d["higher_is_better"] is True
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] > 1e4 and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "blocks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["config"]["workload"].startswith("cfg2")


@pytest.mark.gpu
def test_our_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-cpu-baseline"], 900)
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["dtype"] == "u64"
    assert d["value"] > 5e9 and d["config"]["workload"].startswith("cfg2")
    roof = d["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["peak"] > 1000
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and roof["achieved"] > 3000
    assert roof["algorithmic_bytes_per_launch"] == 160000000
    e2e = d["e2e"]
    assert e2e["value"] > 5e9 and e2e["h2d_bytes_per_step"] == 16 * 2000 * 160 and e2e["d2h_bytes_per_step"] == 128
    assert d["gpu_launches"] == 3 * 32                       # 16 multiplies + 16 folds per step, nothing else
    extra = d["other_kernels"]                               # permute and add, reported beside the headline
    assert "error" not in extra and extra["permute"]["blocks_per_s"] > 5e9 and extra["add"]["gbs_read_plus_write"] > 3000
    assert d["clocks"]["sm_mhz"] and not (set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
