"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) runs the unmodified reference's CPU
build on a bounded sample and prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")

needs_ref = pytest.mark.skipif(not os.access(os.path.join(REF, "nbody_d3"), os.X_OK), reason="oracle/_ref not built")


def run_bench(*args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


@needs_ref
@pytest.mark.parametrize("algo,extra", [("all-pairs", []), ("octree", ["--precision", "double"]), ("bvh", [])])
def test_reference_arm_line(algo, extra):
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--algorithm", algo, "-n", "20000", *extra)
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["unit"] == ("Gpairs/s" if algo == "all-pairs" else "Mbody-steps/s")
    assert d["config"]["workload"].startswith(algo) and "n=20000" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "unmodified reference" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@needs_ref
def test_reference_arm_other_ranks_stay_silent():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""
