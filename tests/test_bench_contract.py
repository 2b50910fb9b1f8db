"""The bench.py contract the driver relies on: the reference arm runs here (it is the CPU port of the reference path) and the recorded
B200 line under profiles/ carries every key of the contract with consistent values."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
             "e2e", "cpu_baseline"}


def _line(text):
    lines = [l for l in text.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly ONE JSON line"
    return json.loads(lines[0])


def test_reference_arm_runs_on_the_host_cores():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = _line(out.stdout)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "nde_column_steps_per_sec_fwd" and d["unit"] == "column-steps/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_recorded_b200_line_carries_the_contract():
    d = _line(open(os.path.join(ROOT, "profiles", "r01_bench_tc_forward_final.json")).read())
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d)
    assert d["metric"] == "nde_column_steps_per_sec_fwd" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert abs(d["value"] - d["config"]["column_steps_per_pass_rank0"] * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] in ("tensor", "hbm") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    # algorithmic flops of one launch / that kernel's own duration (CUDA events around the kernel inside bench.py)
    assert abs(r["achieved"] - d["config"]["column_steps_per_pass_rank0"] * r["flops_per_column_step"] / (r["kernel_ms"] * 1e-3) / 1e12) < 1e-6 * r["achieved"]
    assert r["kernel_ms"] <= d["ms_per_step"]
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 65536 * (32 + 3) * 4 and e["d2h_bytes_per_step"] == 65536 * 129 * 32 * 4 + 3 * 65536 * 4
    assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
