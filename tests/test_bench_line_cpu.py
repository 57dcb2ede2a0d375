"""The committed bench lines (profiles/r1_bench_line*.json, written by bench.py on the GPU box) carry every key of the
measurement contract: the base line, `e2e`, `gpu_launches`, `roofline`, `cpu_baseline`, `clocks`; the reference arm's
line carries `impl`, `cpu_baseline` and a zero-copy `e2e`."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _load(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed yet")
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r1_bench_line.json", "r1_bench_line_n2.json", "r1_bench_line_n8.json"])
def test_bench_line_schema(name):
    d = _load(name)
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                       # BASELINE.md holds no published number for this metric
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 1 and d["warmup"] >= 3
    e2e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e2e)
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and e2e["unit"] == d["unit"]
    assert 0 < e2e["value"] < d["value"]                  # copies inside the timed region: never the resident number
    assert e2e["matches_resident_run"] is True
    assert d["gpu_launches"] == 6 * d["steps"]            # prep_cols, prep_rows, seg_count, pairs, sums, finalize per step
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.0 < r["frac"] < 1.2
    shares = r["step_share"]
    assert 0.9 < sum(shares.values()) <= 1.0 + 1e-6
    clocks = d["clocks"]
    assert clocks and clocks["sm_mhz"] > 0.8 * clocks["sm_max_mhz"]
    assert not set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
        assert c["gpu_matches_oracle_on_sample"]["counts_exact"] and c["gpu_matches_oracle_on_sample"]["stats_within_1e-12"]


def test_reference_arm_line_schema():
    d = _load("r1_bench_reference_line.json")
    assert d["impl"] == "reference" and BASE_KEYS - {"scaling"} <= set(d) | {"scaling"}
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] == 0
