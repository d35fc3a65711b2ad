"""bench.py prints ONE JSON line carrying the keys the driver reads (task contract): checked on the reference arm
without a GPU and on the product arm on a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line: %r" % lines
    return json.loads(lines[0])


@pytest.mark.timeout(900)
def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == base["metric"] and d["unit"] == "projected actions/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_product_arm_line():
    d = _run("--steps", "5", "--warmup", "3")
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d) and "impl" not in d or d.get("impl") == "b200"
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] >= 3 and d["value"] > 1e8 and d["dtype"] == "f32"
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
    assert d["gpu_launches"] == 2 * 5                                  # one projection + one backward kernel per step
    assert r["frac"] < 1.0 and r["dense_model"]["frac_dense_model"] > r["frac"]      # `frac` counts executed flops
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == (33 + 14 + 43) * 4 * (1 << 20) and e["d2h_bytes_per_step"] == (43 + 14) * 4 * (1 << 20)
    assert e["full_rows"]["h2d_bytes_per_step"] == (57 + 14 + 43) * 4 * (1 << 20)
    assert e["value"] < d["value"]                                     # host buffers cross PCIe inside the timed region
    c = d["cpu_baseline"]
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "rpo", "env", "electrical_grid", "evopf.py"))
    assert c["kind"] == ("reference" if have_ref else "port") and c["value"] > 0 and c["cores"] >= 1
    o = d["other_configs"]                                             # BASELINE.json configs[0] / configs[1] beside the headline
    assert o["safe_cartpole"]["value"] > 1e9 and o["spring_pendulum"]["value"] > 1e9
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


@pytest.mark.timeout(900)
def test_reference_arm_train_workload_line():
    """configs[2] on the host cores: the unmodified reference's RPODDPG.run (needs the reference tree or its oracle/_ref copy)."""
    if not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "rpo", "env", "electrical_grid", "evopf.py")):
        pytest.skip("oracle/_ref not present (run oracle/make_ref_copy.sh)")
    d = _run("--impl", "reference", "--workload", "train", "--steps", "3", "--warmup", "1")
    assert d["impl"] == "reference" and d["unit"] == "iterations/s" and d["value"] > 0 and d["steps"] == 3
    assert d["cpu_baseline"]["kind"] == "reference" and "configs[2]" in d["config"]["workload"]


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_product_arm_train_workload_line():
    if not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "rpo", "env", "electrical_grid", "evopf.py")):
        pytest.skip("oracle/_ref not present (run oracle/make_ref_copy.sh)")
    d = _run("--workload", "train", "--steps", "40", "--warmup", "5")
    assert d["unit"] == "iterations/s" and d["value"] > 20 and d["config"]["device"] == "cuda" and d["config"]["replay_size"] >= 40
