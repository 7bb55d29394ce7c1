"""The JSON line bench.py prints is a contract with the driver: check the committed end-of-round lines (profiles/r1i_final_*.json, written by
bench.py on the GPU box) carry every key it asks for, with sane values.  No GPU."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


@pytest.mark.parametrize("name", ["r1i_final_h1.json", "r1i_final_c2.json", "r1i_final_c5.json"])
def test_bench_line_has_the_contract_keys(name):
    l = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in l, k
    assert l["unit"] == "measurements/s" and l["dtype"] == "f64" and l["data"] == "synthetic" and l["higher_is_better"] is True
    assert l["scaling"] == "weak" and l["vs_baseline"] is None and l["steps"] >= 20 and l["warmup"] >= 3
    assert "workload" in l["config"] and "model" not in l["config"]
    e = l["e2e"]
    assert e["unit"] == l["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < l["value"]
    r = l["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0 < r["frac"] < 1
    c = l["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and isinstance(c["sample"], str)
    assert l["gpu_launches"] > 0
    assert set(l["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"} and not set(l["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value is whole-job throughput of the timed region
    rows = l["config"]["measurements_per_step_per_gpu"] * l["n_gpus"]
    assert abs(l["value"] - rows / (l["ms_per_step"] * 1e-3)) < 1e-6 * l["value"]


def test_reference_arm_line():
    l = _line("r1i_final_ref.json")
    assert l["impl"] == "reference" and l["e2e"]["h2d_bytes_per_step"] == 0 and l["e2e"]["d2h_bytes_per_step"] == 0
    assert l["e2e"]["value"] == l["value"] == l["cpu_baseline"]["value"] and l["cpu_baseline"]["kind"] == "port"


def test_optimised_cpu_variant_is_reported_on_the_headline():
    o = _line("r1i_final_h1.json")["cpu_baseline"]["optimised"]
    assert o["kind"] == "port-analytic" and o["value"] > _line("r1i_final_h1.json")["cpu_baseline"]["value"]
