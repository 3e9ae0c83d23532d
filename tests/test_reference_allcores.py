"""The ALL-CORES reference build (oracle/_ref/nbody_d{2,3}_omp: unmodified reference + oracle/pstl_backend_omp.h) must
compute what the serial-backend build of the same source computes; it is the CPU baseline bench.py times."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def have(dim):
    return all(os.access(os.path.join(REF, f"nbody_d{dim}{s}"), os.X_OK) for s in ("", "_omp"))


def state(exe, algo, precision, n=3000, steps=2, threads=4):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    out = subprocess.run([exe, "-n", str(n), "-s", str(steps), "--workload", "galaxy", "--algorithm", algo, "--precision",
                          precision, "--theta", "0.5", "--print-state"], capture_output=True, text=True, check=True, env=env)
    return [ln for ln in out.stdout.splitlines() if not ln.startswith("Total time")]


# deterministic paths only: all-pairs-collapsed accumulates with relaxed float atomics (order-dependent), and -Ofast
# vectorises the float octree differently on the parallel path (last-digit differences in the printed state)
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("algo,precision", [("all-pairs", "float"), ("all-pairs", "double"), ("octree", "double"),
                                            ("bvh", "float"), ("bvh", "double")])
def test_allcores_build_matches_serial_build(dim, algo, precision):
    if not have(dim):
        pytest.skip("oracle/_ref not built")
    a = state(os.path.join(REF, f"nbody_d{dim}_omp"), algo, precision)
    b = state(os.path.join(REF, f"nbody_d{dim}"), algo, precision)
    assert len(a) == len(b) > 3000
    assert a == b


def test_bench_reference_arm_uses_all_cores():
    import bench
    exe, threads = bench.ref_binary(3)
    if exe is None or not exe.endswith("_omp"):
        pytest.skip("all-cores reference build absent")
    assert threads == bench.host_threads() >= 1
