"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU on config 1 (the reference's own
CPU-runnable case: its rasterizer needs CUDA, so the CPU restatement stands in for it here; the aggregator is the genuine
one when oracle/_ref is built) and prints the JSON line the driver expects."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", SMESH_REF_BUDGET_S="30")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "views/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["steps"] == 2
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["product_lib_mapped"] is False, "the reference arm must not load libsmesh_b200.so"


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1", "--steps", "1",
                          "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
