"""The bench line contract (no GPU): the last committed line of each arm carries the keys the driver and the judge read."""
import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parents[1] / "profiles"

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"}


def _last_line(name):
    return json.loads((PROFILES / name).read_text().strip().splitlines()[-1])


def test_single_gpu_line():
    d = _last_line("bench_r02_n1.json")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"].startswith("EVP substep cell-updates/s") and d["unit"] == "cell-updates/s" and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and d["config"]["grid"] == [4096, 4096]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    # achieved = algorithmic bytes per launch / average launch duration of the timed region
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["algorithmic_bytes_per_launch"] == 144 * 4096 * 4096
    f = r["fp64"]
    assert f["bound"] == "fp64" and abs(f["frac"] - f["achieved"] / f["peak"]) < 1e-12
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]          # copies inside the timed region
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "port"
    assert d["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
    assert d["fused_stats"]["inputs_failed_validation"] == 0
    assert {"config1_anticyclone_128_as_shipped", "config4_coastline_8192x4096_masked", "config5_arctic_cap_4320x336_one_gpu"} <= set(d["configs"])


@pytest.mark.parametrize("name,n", [("bench_r02_n2_slabs.json", 2), ("bench_r02_n4_slabs.json", 4), ("bench_r02_n8_slabs.json", 8)])
def test_multi_gpu_lines(name, n):
    d = _last_line(name)
    assert BASE_KEYS <= set(d)
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["config"]["per_gpu"] == [16384, 2048]
    assert d["multi_gpu_parity"] == "bitwise"
    w = d["weak_baseline_1gpu"]
    assert w["grid"] == [16384, 2048] and 0.85 <= d["value"] / (n * w["value"]) <= 1.05      # like-for-like weak-scaling efficiency
    assert d["roofline"]["traffic"] is None                                                  # no ncu capture of this block size
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0


def test_reference_arm_line():
    d = _last_line("bench_r02_reference_arm_n4.json")
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["config"]["sample_grid"] != d["config"]["grid"]          # the grid that ran is stated
