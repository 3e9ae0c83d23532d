"""stdpar-nbody_b200 — B200-native drop-in for the force + leapfrog hot path of UoB-HPC/stdpar-nbody.

The product is `lib/libnbx.so` (hand-written sm_100a CUDA behind the C ABI of include/nbx.h, sources in csrc/) and the
C++ host driver in host/. This package only carries the ctypes binding used by the tests and bench.py.
Import it through `_pkg.load()` at the repo root (the directory name is not a Python identifier).
"""
from . import nbx  # noqa: F401
from .nbx import Engine, NbxError  # noqa: F401
