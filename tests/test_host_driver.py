"""The C++ host driver (stdpar-nbody_b200/host/main.cpp -> bin/nbody_d{2,3}) keeps the reference's CLI surface.
CPU part: argument handling and bit-identical workload generation (vs the oracle and, where present, the compiled
reference). GPU part: the README's check (README.md:122-129) — same printed final state as the reference binary."""
import os
import subprocess

import numpy as np
import pytest

import _pkg
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "stdpar-nbody_b200", "bin")
nbx = _pkg.load().nbx


def exe(dim):
    path = os.path.join(BIN, f"nbody_d{dim}")
    if not os.access(path, os.X_OK):
        import __graft_entry__ as ge
        ge.build()
    return path


def run(dim, args, cwd=None, check=True):
    r = subprocess.run([exe(dim)] + args, capture_output=True, text=True, cwd=cwd)
    if check:
        assert r.returncode == 0, r.stdout + r.stderr
    return r


def dry_state(dim, dtype, workload, n, tmp_path):
    prec = "float" if dtype == np.float32 else "double"
    run(dim, ["-n", str(n), "--workload", workload, "--precision", prec, "--save", "pos", "--dry-run"], cwd=tmp_path)
    buf = open(tmp_path / "positions.bin", "rb").read()
    hdr = np.frombuffer(buf, np.uint32, 4)
    assert hdr[2] == np.dtype(dtype).itemsize and hdr[3] == dim
    rest = np.frombuffer(buf, dtype, offset=16)
    size = len(rest) // (2 * dim + 1)
    x = rest[: size * dim].reshape(size, dim)
    v = rest[size * dim: 2 * size * dim].reshape(size, dim)
    m = rest[2 * size * dim:]
    return m, x, v


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [10, 11, 1000])
def test_galaxy_is_bit_identical_to_oracle(oracle, tmp_path, dim, dtype, n):
    m, x, v = dry_state(dim, dtype, "galaxy", n, tmp_path)
    s = oracle.galaxy(n, dtype, dim)
    assert m.tobytes() == s["m"].tobytes() and x.tobytes() == s["x"].tobytes() and v.tobytes() == s["v"].tobytes()


@pytest.mark.skipif(not O.ref_available(3), reason="oracle/_ref not built")
@pytest.mark.parametrize("workload,dim", [("uniform", 2), ("uniform", 3), ("plummer", 3)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_other_workloads_bit_identical_to_reference(tmp_path, workload, dim, dtype):
    m, x, v = dry_state(dim, dtype, workload, 500, tmp_path)
    buf = O.refdump(workload, dtype, dim, 500)
    ref, _ = O._parse_state(buf, 500, dim, dtype, 8)
    assert m.tobytes() == ref["m"].tobytes() and x.tobytes() == ref["x"].tobytes() and v.tobytes() == ref["v"].tobytes()


def test_cli_errors_match_reference():
    r = run(3, ["--algorithm", "nope"], check=False)
    assert r.returncode != 0 and 'Unknown algorithm: "nope".' in r.stderr
    r = run(3, ["--bogus"], check=False)
    assert r.returncode != 0 and "Unknown argument: '--bogus'" in r.stdout
    r = run(3, ["--csv-detailed", "--csv-total"], check=False)
    assert r.returncode != 0 and "Cannot capture a CSV detailed and coarse trace" in r.stderr
    r = run(2, ["--workload", "plummer", "--dry-run"], check=False)
    assert r.returncode != 0 and "Cannot build Plummer model for D=2" in r.stderr
    assert "--algorithm" in run(3, ["--help"]).stdout


@pytest.mark.skipif(not O.ref_available(3), reason="oracle/_ref not built")
def test_starting_state_print_matches_reference_binary():
    mine = run(2, ["-n", "10", "--workload", "galaxy", "--print-state", "--dry-run"]).stdout
    ref = subprocess.run([os.path.join(O.REF_DIR, "nbody_d2"), "-n", "10", "--workload", "galaxy", "--algorithm", "all-pairs",
                          "--print-state"], capture_output=True, text=True).stdout
    start = ref[: ref.index("Starting simulation")]
    assert mine == start


# ---- GPU -------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.skipif(not O.ref_available(3), reason="oracle/_ref not built")
@pytest.mark.parametrize("algo", ["all-pairs", "all-pairs-collapsed", "octree", "bvh"])
@pytest.mark.parametrize("dim,prec", [(2, "float"), (3, "double")])
def test_final_state_print_matches_reference_binary(algo, dim, prec):
    """Same command line into both binaries: banner, starting state and final state (3 significant digits, components 0
    and 1: src/system.h:92-94) must be textually identical. -s 12 => 10 warm-up + 2 timed steps."""
    args = ["-n", "64", "-s", "12", "--workload", "galaxy", "--algorithm", algo, "--precision", prec, "--theta", "0.5",
            "--print-state"]
    ref_exe = os.path.join(O.REF_DIR, f"nbody_d{dim}")
    if algo == "bvh" and dim == 2 and prec == "float":
        # 2-D float Hilbert keys hit the float->u32 overflow that is UB in the reference: only the
        # AVX-512 -march=native build saturates like CUDA does; the generic build wraps and sorts 2 bodies elsewhere.
        ref_exe = os.path.join(O.REF_DIR, "nbody_d2_native")
        if not (O.avx512_host() and os.access(ref_exe, os.X_OK)):
            pytest.skip("needs the AVX-512 native reference build")
    mine = run(dim, args).stdout
    ref = subprocess.run([ref_exe] + args, capture_output=True, text=True).stdout

    def strip(out):
        return [ln for ln in out.splitlines() if not ln.startswith("Total time")]
    a, b = strip(mine), strip(ref)
    assert len(a) == len(b)
    bad = [(x, y) for x, y in zip(a, b) if x != y]
    # the print has 3 significant digits; tolerate a last-digit flip on at most 2 lines for float
    assert len(bad) <= (2 if prec == "float" else 0), bad[:4]


@pytest.mark.gpu
@pytest.mark.skipif(not O.ref_available(3), reason="oracle/_ref not built")
def test_csv_and_saved_positions(tmp_path):
    args = ["-n", "200", "-s", "3", "--workload", "galaxy", "--algorithm", "octree", "--precision", "double", "--csv-detailed",
            "--save", "all"]
    d1, d2 = tmp_path / "mine", tmp_path / "ref"
    d1.mkdir(); d2.mkdir()
    mine = run(3, args, cwd=d1).stdout.splitlines()
    ref = subprocess.run([os.path.join(O.REF_DIR, "nbody_d3")] + args, capture_output=True, text=True, cwd=d2).stdout.splitlines()
    assert mine[0] == ref[0]  # CSV header
    assert mine[1].split(",")[:5] == ref[1].split(",")[:5] and len(mine[1].split(",")) == len(ref[1].split(","))
    pa, pb = open(d1 / "positions.bin", "rb").read(), open(d2 / "positions.bin", "rb").read()
    assert len(pa) == len(pb) and pa[:16] == pb[:16]
    xa, xb = np.frombuffer(pa, np.float64, offset=16), np.frombuffer(pb, np.float64, offset=16)
    assert np.allclose(xa, xb, rtol=1e-10, atol=1e-12)
    ea, eb = open(d1 / "energy.bin", "rb").read(), open(d2 / "energy.bin", "rb").read()
    assert len(ea) == len(eb) and ea[:8] == eb[:8]
    assert np.allclose(np.frombuffer(ea, np.float64, offset=8), np.frombuffer(eb, np.float64, offset=8), rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.skipif(not O.ref_available(3), reason="oracle/_ref not built")
def test_load_workload_matches_reference_binary(tmp_path):
    """--workload load <file.bin> (src/saving.h:25-68; writer: scripts/thuering_nbody/conv_csv.py:62-80)."""
    rng = np.random.default_rng(4)
    n, dim = 50, 3
    body = np.concatenate([rng.random((n, 1)) * 5, rng.standard_normal((n, dim)) * 10, rng.standard_normal((n, dim)) * 0.01], axis=1)
    f = tmp_path / "in.bin"
    with open(f, "wb") as fh:
        fh.write(np.array([n, dim], np.uint32).tobytes())
        fh.write(np.array([0.5, 1e-3], np.float32).tobytes())
        fh.write(body.astype(np.float32).tobytes())
    args = ["-s", "12", "--workload", "load", str(f), "--algorithm", "all-pairs", "--precision", "double", "--print-state"]
    mine = run(3, args).stdout
    ref = subprocess.run([os.path.join(O.REF_DIR, "nbody_d3")] + args, capture_output=True, text=True).stdout
    strip = lambda out: [ln for ln in out.splitlines() if not ln.startswith("Total time")]  # noqa: E731
    assert strip(mine) == strip(ref)
