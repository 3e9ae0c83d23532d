"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/nbx.h declares, and refuses to run
without a GPU (no CPU fallback). No compute is invoked here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nbx():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "stdpar-nbody_b200", "lib", "libnbx.so")):
        ge.build()
    return _pkg.load().nbx


def test_header_symbols_all_exported(nbx):
    hdr = open(os.path.join(ROOT, "include", "nbx.h")).read()
    declared = set(re.findall(r"^\s*(?:const char\*|int)\s+(nbx_\w+)\s*\(", hdr, re.M))
    assert declared, "no declarations parsed"
    assert declared == set(nbx.SYMBOLS)
    lib = nbx.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.nbx_version() == 100


def test_config_struct_matches_header(nbx):
    # 4+4+4+4+4+4 + 3*8 + 4+4+4+4 = 64 bytes, doubles 8-aligned
    assert C.sizeof(nbx.Config) == 64


def test_invalid_arguments_fail_loudly(nbx):
    with pytest.raises(nbx.NbxError) as ei:
        nbx.Engine(10, 4, np.float32, "all-pairs", 10, 1e-4)
    assert ei.value.code == -1 and "dim" in str(ei.value)
    with pytest.raises(nbx.NbxError):
        nbx.Engine(0, 3, np.float32, "all-pairs", 10, 1e-4)


@pytest.mark.skipif(_pkg.load().nbx.device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback(nbx):
    with pytest.raises(nbx.NbxError) as ei:
        nbx.Engine(10, 3, np.float32, "all-pairs", 10, 1e-4)
    assert ei.value.code == -3


def test_shard_bounds_cover_everything(nbx):
    for n in (1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            spans = [nbx.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
