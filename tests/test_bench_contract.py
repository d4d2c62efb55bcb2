"""bench.py prints ONE JSON line with the keys the driver's contract names."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def run(*args, timeout=300):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=timeout)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout                     # exactly one line on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1")
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "GB/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "no sampling" in d["cpu_baseline"]["sample"]
    assert d["steps"] == 1 and d["warmup"] == 3            # what ran is what is reported (warm-ups are floored at 3)
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_cuda_arm_line():
    d = run("--gpus", "1", "--steps", "5", "--warmup", "3", "--blocks", "4", "--secondary-reps", "2")
    assert BASE | {"roofline", "gpu_launches", "clocks"} <= set(d) and "impl" not in d
    assert d["metric"].startswith("JopDBlock mul!+adjoint") and d["unit"] == "GB/s" and d["dtype"] == "f32"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["achieved"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert "traffic" in r
    assert d["gpu_launches"] == 2 * 5                      # one fused GEMV launch per apply, two applies per step
    assert d["e2e"]["h2d_bytes_per_step"] == 4 * 4096 * 4 and d["e2e"]["d2h_bytes_per_step"] == 4 * 4096 * 4
    assert d["e2e"]["value"] > 0 and d["e2e"]["value"] != d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["parity_relerr_block1"] <= 1e-5
    assert d["e2e"]["sync_every_step"]["value"] > 0
    # the grid exchange (BASELINE configs[2]) and the scalars are timed and parity-checked in the same run
    grids = {tuple(c["grid"]) for c in d["secondary"]["config3"]}
    assert grids == {(2, 2), (1, 4)}
    for c in d["secondary"]["config3"]:
        assert c["relerr_max"] <= 1e-12 and set(c["relerr"]) == {"forward_block_row_1", "adjoint_block_column_1", "dot", "norm"}
        assert c["GBps_fwd"] > 0 and c["GBps_adj"] > 0 and c["algorithmic_bytes_per_apply"] == 34360786944
