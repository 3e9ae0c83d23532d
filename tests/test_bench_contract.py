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


# ---- our arm: host-side logic that needs no GPU -----------------------------------------------------------------------
def test_our_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the nbx arm exits with a message instead of measuring anything."""
    sys.path.insert(0, ROOT)
    import _pkg
    if _pkg.load().nbx.device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_extra_configs_and_verify_cases_cover_baseline_json():
    sys.path.insert(0, ROOT)
    import bench
    cfgs = bench.EXTRA_CONFIGS
    assert cfgs["C3"] == dict(algorithm="all-pairs-collapsed", precision="double", dim=3, n=262144, theta=0.5)
    assert cfgs["C4"] == dict(algorithm="octree", precision="double", dim=3, n=10_000_000, theta=0.5)
    assert cfgs["bvh_f32_n10M"]["algorithm"] == "bvh" and cfgs["bvh_f32_n10M"]["n"] == 10_000_000
    algos = {c[0] for c in bench.VERIFY_CASES}
    assert algos == {"all-pairs", "all-pairs-collapsed", "octree", "bvh"}
    assert ("all-pairs", "float", 1_000_000, 3, 1, 0) in bench.VERIFY_CASES and ("bvh", "float", 10_000_000, 3, 1, 0) in bench.VERIFY_CASES
    import argparse
    ns = argparse.Namespace(algorithm="all-pairs", precision="float", dim=3, n=1_000_000)
    assert bench.is_headline(ns)
    ns.n = 999
    assert not bench.is_headline(ns)


def test_hbm_phase_bytes_are_the_documented_per_body_figures():
    """leapfrog 7 records per body; BVH sort phase = keys 24 + onesweep 8 + 8 x 24 + gather 4 + 8 records."""
    sys.path.insert(0, ROOT)
    import argparse

    import bench
    n = 1000
    b = bench.hbm_phase_bytes(argparse.Namespace(algorithm="bvh"), n, 4)
    assert b["accel"] == 7 * 16 * n and b["sort"] == (24 + 200 + 132) * n
    b8 = bench.hbm_phase_bytes(argparse.Namespace(algorithm="bvh"), 1 << 20, 4, world=8)
    assert b8["sort"] < bench.hbm_phase_bytes(argparse.Namespace(algorithm="bvh"), 1 << 20, 4)["sort"]
    o = bench.hbm_phase_bytes(argparse.Namespace(algorithm="octree"), n, 8)
    assert o["accel"] == 7 * 32 * n and o["sort"] == (40 + 200 + 16) * n
    assert set(bench.hbm_phase_bytes(argparse.Namespace(algorithm="all-pairs"), n, 4)) == {"accel"}
