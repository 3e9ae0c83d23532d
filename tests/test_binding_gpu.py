"""The drop-in binding proven with the reference's own code: oracle/_ref/nbody_nbx_d{2,3} is the UNMODIFIED
/root/reference/src/main.cpp (parse_args, workload builders, run_simulation, System::print) with `run_nbx`
(oracle/nbx_backend.h = INTEGRATION.md §2) bound as its sim_func_t and linked to libnbx.so (`make -C oracle binding`).
Same command line into that binary and into the stock reference binary => the same text."""
import os
import subprocess

import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def nbx_exe(dim):
    return os.path.join(O.REF_DIR, f"nbody_nbx_d{dim}")


needs_binding = pytest.mark.skipif(not (os.access(nbx_exe(2), os.X_OK) and os.access(nbx_exe(3), os.X_OK) and O.ref_available(3)),
                                   reason="oracle/_ref binding not built (needs /root/reference at build time)")


def strip(out):
    return [ln for ln in out.splitlines() if not ln.startswith("Total time")]


@needs_binding
def test_integration_md_prints_the_compiled_binding():
    """INTEGRATION.md §2 shows exactly the header that is compiled (no prose drift)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    src = open(os.path.join(ROOT, "oracle", "nbx_backend.h")).read()
    body = src[src.index("#pragma once"):]
    assert body.strip() in doc


@needs_binding
def test_binding_links_the_product_library():
    out = subprocess.run(["ldd", nbx_exe(3)], capture_output=True, text=True).stdout
    assert "libnbx.so" in out


@pytest.mark.gpu
@needs_binding
@pytest.mark.parametrize("algo", ["all-pairs", "all-pairs-collapsed", "octree", "bvh"])
@pytest.mark.parametrize("dim,prec", [(2, "float"), (3, "double"), (3, "float")])
def test_binding_final_state_matches_reference_binary(algo, dim, prec):
    args = ["-n", "64", "-s", "12", "--workload", "galaxy", "--algorithm", algo, "--precision", prec, "--theta", "0.5",
            "--print-state"]
    ref_exe = os.path.join(O.REF_DIR, f"nbody_d{dim}")
    if algo == "bvh" and dim == 2 and prec == "float":  # only the AVX-512 native build saturates like CUDA
        ref_exe = os.path.join(O.REF_DIR, "nbody_d2_native")
        if not (O.avx512_host() and os.access(ref_exe, os.X_OK)):
            pytest.skip("needs the AVX-512 native reference build")
    mine = subprocess.run([nbx_exe(dim)] + args, capture_output=True, text=True)
    assert mine.returncode == 0, mine.stdout + mine.stderr
    ref = subprocess.run([ref_exe] + args, capture_output=True, text=True).stdout
    a, b = strip(mine.stdout), strip(ref)
    assert len(a) == len(b) and a[0] == b[0] == "Starting state:"
    bad = [(x, y) for x, y in zip(a, b) if x != y]
    # the print has 3 significant digits; tolerate a last-digit flip on at most 2 lines for float
    assert len(bad) <= (2 if prec == "float" else 0), bad[:4]


@pytest.mark.gpu
@needs_binding
def test_binding_uniform_and_csv_row():
    args = ["-n", "500", "-s", "11", "--workload", "uniform", "--algorithm", "octree", "--precision", "double", "--csv-total"]
    mine = subprocess.run([nbx_exe(3)] + args, capture_output=True, text=True)
    assert mine.returncode == 0, mine.stdout + mine.stderr
    ref = subprocess.run([os.path.join(O.REF_DIR, "nbody_d3")] + args, capture_output=True, text=True).stdout.splitlines()
    got = mine.stdout.splitlines()
    assert got[0] == ref[0] and got[1].split(",")[:5] == ref[1].split(",")[:5]


@needs_binding
def test_binding_fails_loudly_without_a_gpu():
    """No CPU fallback behind the binding either: without a device the run_nbx call throws the library's error."""
    import _pkg
    if _pkg.load().nbx.device_count() > 0:
        pytest.skip("a GPU is present")
    r = subprocess.run([nbx_exe(3), "-n", "16", "-s", "11", "--workload", "galaxy", "--algorithm", "all-pairs"], capture_output=True,
                       text=True)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
