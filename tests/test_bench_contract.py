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


def test_bench_numpy_restatement_matches_the_oracle():
    """bench.py checks the GPU with its OWN numpy restatement (precheck, host-known truth); that restatement is pinned
    to the oracle here, on CPU, for three contexts incl. odd L."""
    import importlib.util
    import numpy as np
    from oracle.pyoracle import Oracle, random_blocks, random_key
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    o = Oracle()
    rng = np.random.default_rng(31)
    for N, D in ((1247, 3), (191, 2), (16383, 4), (128, 2)):
        L = bench.words_per_block(N)
        s = random_key(rng, N, D)
        mask = bench.np_key_mask(N, s)
        a = bench.planted_blocks(rng, 23, N, mask, k=7)
        b = bench.planted_blocks(rng, 9, N, mask, k=4)
        assert not np.any(a.reshape(-1, L)[:, -1] & ~np.uint64(((1 << 64) - 1) << ((64 - N % 64) % 64) & ((1 << 64) - 1)))   # pad bits stay zero
        prod = bench.np_mul(a, b, L)
        assert np.array_equal(prod, o.mul(a, b, L))
        assert bench.np_count(a, L, mask) == o.count_satisfied(a, N, s) >= 7
        assert bench.np_count(prod, L, mask) == o.count_satisfied(prod, N, s) == bench.np_count(a, L, mask) * bench.np_count(b, L, mask)
        perm = rng.permutation(N).astype(np.uint64)
        assert np.array_equal(bench.np_permute(prod, N, perm), o.permute_all(prod, N, perm))


@pytest.mark.gpu
def test_our_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--sustain-s", "0.3"], 1200)
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["dtype"] == "u64"
    assert d["value"] > 1e10 and d["config"]["workload"].startswith("cfg2") and d["config"]["mode"].startswith("fused")
    roof = d["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["peak"] > 1000
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and roof["achieved"] > 3000
    assert roof["algorithmic_bytes_per_launch"] == 160000000
    e2e = d["e2e"]
    assert e2e["value"] > 1e10 and e2e["h2d_bytes_per_step"] == 16 * 2000 * 160 and e2e["d2h_bytes_per_step"] == 128
    assert d["gpu_launches"] == 3 * 16                       # one fused multiply->decrypt kernel per pair, nothing else
    assert d["precheck"]["passed"] is True and len(d["precheck"]["cases"]) == 3
    two = d["two_pass"]                                      # the round-1 step (separate kernels) on the same buffers
    assert two["gpu_launches"] == 3 * 32 and 5e9 < two["value"] < d["value"]
    assert d["sustained"]["seconds"] >= 0.3 and d["sustained"]["value"] > 1e10
    extra = d["other_kernels"]                               # permute and add, reported beside the headline
    assert "error" not in extra and extra["permute"]["blocks_per_s"] > 5e9 and extra["add"]["gbs_read_plus_write"] > 3000
    ow = d["other_workloads"]                                # BASELINE.json configs[3] and [4]
    for name in ("cfg5_300x300", "cfg5_2000x2000", "cfg4_chain"):
        assert "error" not in ow[name], ow[name]
    assert ow["cfg4_chain"]["blocks_per_gpu"] == 125000000 and ow["cfg4_chain"]["multiply_chain"]["frac_of_peak"] > 0.8
    assert ow["cfg5_2000x2000"]["multiply"]["frac_of_peak"] > 0.8
    cpp = d["e2e_cpp"]                                       # the same step through libcertFHE.so
    assert cpp.get("checked") is True and cpp["value"] > 1e9, cpp
    assert d["clocks"]["sm_mhz"] and not (set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})


@pytest.mark.gpu
def test_two_pass_mode_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-extras", "--mode", "two-pass"], 600)
    assert d["config"]["mode"].startswith("two-pass") and d["gpu_launches"] == 3 * 32
    assert set(d["kernels"]) >= {"multiply", "decrypt"} and d["value"] > 5e9
