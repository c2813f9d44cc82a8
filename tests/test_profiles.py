"""The bench lines committed under profiles/ are what DESIGN.md / README.md quote: each must satisfy bench.py's contract and be
internally consistent (value = global batch / step time, roofline.frac = achieved / peak, parity leg green at N > 1)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "e2e", "gpu_launches", "clocks", "roofline")


@pytest.mark.parametrize("n", [1, 2, 4])
def test_committed_bench_line(n):
    path = os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    for k in KEYS:
        assert k in d, k
    assert d["n_gpus"] == n and d["metric"] == "clip_head_fwd_bwd_pairs_per_sec" and d["dtype"] == "bf16"
    B = d["config"]["global_batch"]
    assert abs(d["value"] - B / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.2 < r["frac"] < 1.0
    e = d["e2e"]
    assert 0 < e["value"] <= d["value"] * 1.001 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n > 1:
        assert d["parity"]["ok"] is True and d["parity"]["loss_rel"] <= 1e-5
    else:
        assert "cpu_baseline" in d and {"cfg2", "cfg4_zeroshot"} <= set(d["extra"])
