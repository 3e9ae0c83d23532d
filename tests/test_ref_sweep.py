"""the reference's benchmark sweep + scraper. tools/ref_sweep.py runs ci/benchmark's matrix
(ci/benchmark:64-98) through the nbx host driver and writes the same log; /root/reference/ci/data.py must parse it.
The reference tree only exists in the build container, so: the CPU test pipes a COMMITTED sweep log (recorded on a B200
by tools/ref_sweep.py, profiles/bench/r02_ref_sweep_b200.log) through the real ci/data.py and compares with the tool's
own restatement of the scraper; the GPU test runs a small sweep live and scrapes it with that restatement."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_sweep  # noqa: E402

DATA_PY = "/root/reference/ci/data.py"
LOG = os.path.join(ROOT, "profiles", "bench", "r02_ref_sweep_b200.log")
HEADER = "gpu,driver,cpu,#cores,seq,compiler,hostname,algorithm,dim,precision,nsteps,nbodies,total [s]"


def check_rows(rows, nrows):
    assert rows[0] == HEADER
    body = [r.split(",") for r in rows[1:]]
    assert len(body) == nrows
    for r in body:
        assert len(r) == 13 and r[7] in ("all-pairs", "all-pairs-collapsed", "octree", "bvh")
        assert r[5] in ("nbx", "gcc-omp-shim") and float(r[12]) >= 0 and int(r[11]) > 0 and r[8] == "3" and r[9] == "64"


@pytest.mark.skipif(not (os.path.exists(DATA_PY) and os.path.exists(LOG)), reason="needs /root/reference and the committed sweep log")
def test_committed_sweep_log_through_the_references_scraper():
    r = subprocess.run([sys.executable, DATA_PY, LOG], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = r.stdout.strip().splitlines()
    # the full ci/benchmark matrix: 4 algorithms at n = 100 000 + octree, bvh at n = 1 000 000 (nbx), + the reference arm
    nbx_rows = [x for x in rows[1:] if ",nbx," in x]
    assert len(nbx_rows) == 6
    assert sorted((x.split(",")[7], x.split(",")[11]) for x in nbx_rows) == sorted(
        [(a, "100000") for a in ref_sweep.ALGOS_ALL] + [(a, "1000000") for a in ref_sweep.ALGOS_LARGE])
    assert all(x.split(",")[10] == "190" for x in nbx_rows)  # -s 200 minus the 10 hidden warm-up steps
    assert "B200" in nbx_rows[0].split(",")[0]
    check_rows(rows, len(rows) - 1)
    ref_rows = [x for x in rows[1:] if ",gcc-omp-shim," in x]  # the unmodified reference's CPU build, bounded sizes
    assert len(ref_rows) in (0, 6)
    # the tool's restatement of the scraper agrees with the real one, line for line
    assert ref_sweep.parse_log(open(LOG).read().splitlines()) == rows


def test_scraper_restatement_on_a_synthetic_log():
    log = ["name, driver_version", "NVIDIA B200, 580.159", "Model name:   Some CPU", "Core(s) per socket:  16", "hostname:box",
           "compiler:nbx", "algorithm,dim,precision,nsteps,nbodies,total [s]", "octree,3,64,190,100000,0.21", "compiler:nbx",
           "algorithm,dim,precision,nsteps,nbodies,total [s]", "bvh,3,64,190,100000,0.50"]
    rows = ref_sweep.parse_log(log)
    assert rows == [HEADER, "NVIDIA B200,580.159,Some CPU,16,False,nbx,box,octree,3,64,190,100000,0.21",
                    "NVIDIA B200,580.159,Some CPU,16,False,nbx,box,bvh,3,64,190,100000,0.50"]
    if os.path.exists(DATA_PY):
        import tempfile
        with tempfile.NamedTemporaryFile("w", suffix=".log", delete=False) as f:
            f.write("\n".join(log) + "\n")
        r = subprocess.run([sys.executable, DATA_PY, f.name], capture_output=True, text=True)
        os.unlink(f.name)
        assert r.stdout.strip().splitlines() == rows


@pytest.mark.gpu
def test_live_small_sweep_scrapes():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_sweep.py"), "--steps", "12", "--small", "3000", "--large", "20000",
                        "--csv"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = r.stdout.strip().splitlines()
    check_rows(rows, 6)
    assert all(x.split(",")[10] == "2" for x in rows[1:])  # 12 - 10 warm-up steps
