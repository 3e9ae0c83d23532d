"""The OpenMP PSTL backend behind the all-cores reference build (oracle/pstl_backend_omp.h + the shim copy of libstdc++'s pstl/
headers under oracle/_ref/pstl_shim) checked on its own: oracle/pstl_shim_test.cpp runs the parallel algorithms the
reference uses and compares them with the sequential ones."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "_ref", "pstl_shim")
GXX = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else shutil.which("g++")


@pytest.mark.skipif(not os.path.isdir(os.path.join(SHIM, "pstl")) or not GXX, reason="pstl shim not built (make -C oracle ref)")
def test_openmp_pstl_backend(tmp_path):
    exe = str(tmp_path / "pstl_shim_test")
    subprocess.run([GXX, "-std=c++20", "-O2", "-fopenmp", "-isystem", SHIM, os.path.join(ROOT, "oracle", "pstl_shim_test.cpp"),
                    "-o", exe], check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="4"), timeout=300)
    assert r.returncode == 0 and "PSTL_SHIM_TEST PASS threads=4" in r.stdout, r.stdout + r.stderr
