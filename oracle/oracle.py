"""ctypes front-end for the CPU oracle (oracle/nbody_oracle.c) and for the compiled reference
(oracle/_ref/refdump_d{2,3}).  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

ALGOS = {"all-pairs": 0, "all-pairs-collapsed": 1, "octree": 2, "bvh": 3}


def build(fast: bool = True) -> None:
    """Compile the C restatement (and, where /root/reference exists, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", HERE, "all", "CC=gcc", "CXX=g++"], check=True)


def _lib(fast: bool):
    name = "libnbody_oracle_fast.so" if fast else "libnbody_oracle.so"
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


def _suffix(dtype, dim):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return f"f{dim}", C.c_float
    if dtype == np.float64:
        return f"d{dim}", C.c_double
    raise ValueError(dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """Thin typed wrapper; every method mirrors one reference function (see the C file for file:line)."""

    def __init__(self, fast: bool = False):
        self.lib = _lib(fast)
        self.fast = fast

    def _fn(self, name, dtype, dim, restype=None):
        suf, _ = _suffix(dtype, dim)
        f = getattr(self.lib, f"{name}_{suf}")
        f.restype = restype
        return f

    # models.h:112-136
    def galaxy(self, n, dtype, dim):
        size = int(2 * (n / 2.0))
        # the model always places its two centre bodies (models.h:126,132): for n = 1 the reference writes the second one
        # past its 1-body System; give the restatement room for it and return the System-sized prefix
        cap = max(size, 2)
        m = np.zeros(cap, dtype)
        x = np.zeros((cap, dim), dtype)
        v = np.zeros((cap, dim), dtype)
        got = self._fn("nbo_galaxy", dtype, dim, C.c_uint32)(C.c_uint32(n), _p(m), _p(x), _p(v))
        assert got == size
        m, x, v = m[:size].copy(), x[:size].copy(), v[:size].copy()
        return dict(m=m, x=x, v=v, a=np.zeros_like(x), ao=np.zeros_like(x), dt=np.dtype(dtype).type(10.0), G=np.dtype(dtype).type(1e-4))

    # all_pairs.h:14-27
    def all_pairs_force(self, m, x, G, targets=None):
        n, dim = x.shape
        _, ct = _suffix(x.dtype, dim)
        if targets is None:
            a = np.empty_like(x)
            self._fn("nbo_all_pairs_force", x.dtype, dim)(C.c_uint32(n), ct(G), _p(m), _p(x), _p(a))
            return a
        targets = np.ascontiguousarray(targets, np.uint32)
        a = np.empty((len(targets), dim), x.dtype)
        self._fn("nbo_all_pairs_force_targets", x.dtype, dim)(
            C.c_uint32(n), ct(G), _p(m), _p(x), C.c_uint32(len(targets)), _p(targets), _p(a))
        return a

    def all_pairs_force_truth(self, m, x, G, targets=None):
        """The formula of all_pairs.h:14-27 in DOUBLE arithmetic with eps of x's own precision (see the C file)."""
        n, dim = x.shape
        eps = float(np.finfo(x.dtype).eps)
        m64, x64 = np.ascontiguousarray(m, np.float64), np.ascontiguousarray(x, np.float64)
        f = getattr(self.lib, f"nbo_all_pairs_force_eps_d{dim}")
        f.restype = None
        if targets is None:
            a = np.empty_like(x64)
            f(C.c_uint32(n), C.c_double(G), C.c_double(eps), _p(m64), _p(x64), C.c_uint32(0), None, _p(a))
            return a
        targets = np.ascontiguousarray(targets, np.uint32)
        a = np.empty((len(targets), dim), np.float64)
        f(C.c_uint32(n), C.c_double(G), C.c_double(eps), _p(m64), _p(x64), C.c_uint32(len(targets)), _p(targets), _p(a))
        return a

    # all_pairs.h:29-50 (updates a in place, returns it)
    def collapsed_force(self, m, x, a, ao, G):
        n, dim = x.shape
        _, ct = _suffix(x.dtype, dim)
        a = np.ascontiguousarray(a).copy()
        self._fn("nbo_collapsed_force", x.dtype, dim)(C.c_uint32(n), ct(G), _p(m), _p(x), _p(a), _p(ao))
        return a

    # system.h:52-60 (in place on copies; returns x, v, ao)
    def accelerate(self, x, v, a, ao, dt):
        n, dim = x.shape
        _, ct = _suffix(x.dtype, dim)
        x, v, ao = x.copy(), v.copy(), ao.copy()
        self._fn("nbo_accelerate", x.dtype, dim)(C.c_uint32(n), ct(dt), _p(x), _p(v), _p(a), _p(ao))
        return x, v, ao

    def energies(self, m, x, v, G):
        n, dim = x.shape
        _, ct = _suffix(x.dtype, dim)
        k, g = ct(0), ct(0)
        self._fn("nbo_energies", x.dtype, dim)(C.c_uint32(n), ct(G), _p(m), _p(x), _p(v), C.byref(k), C.byref(g))
        return k.value, g.value

    # bvh.h:17-22
    def bbox(self, x):
        n, dim = x.shape
        lo, hi = np.empty(dim, x.dtype), np.empty(dim, x.dtype)
        self._fn("nbo_bbox", x.dtype, dim)(C.c_uint32(n), _p(x), _p(lo), _p(hi))
        return lo, hi

    # bvh.h:33-45
    def keys(self, x, lo, hi):
        n, dim = x.shape
        k = np.empty(n, np.uint64)
        self._fn("nbo_keys", x.dtype, dim)(C.c_uint32(n), _p(x), _p(lo), _p(hi), _p(k))
        return k

    def hilbert(self, cells):
        cells = np.ascontiguousarray(cells, np.uint32)
        dim = cells.shape[1]
        f = self.lib.nbo_hilbert2 if dim == 2 else self.lib.nbo_hilbert3
        f.restype = C.c_uint64
        return np.array([f(*[C.c_uint32(int(c)) for c in row]) for row in cells], np.uint64)

    # bvh.h:62-69 (stable)
    def sort_perm(self, keys):
        keys = np.ascontiguousarray(keys, np.uint64)
        perm = np.empty(len(keys), np.uint32)
        self.lib.nbo_sort_perm(C.c_uint32(len(keys)), _p(keys), _p(perm))
        return perm

    def bvh_levels(self, n):
        f = self.lib.nbo_bvh_levels_f3
        f.restype = C.c_uint32
        return int(f(C.c_uint32(n)))

    # bvh.h:175-244
    def bvh_build(self, m, x):
        n, dim = x.shape
        nn = (1 << self.bvh_levels(n)) - 1
        node_m = np.zeros((nn, dim + 1), x.dtype)
        bw = np.zeros(nn, x.dtype)
        b = np.zeros((nn, 2, dim), x.dtype)
        self._fn("nbo_bvh_build", x.dtype, dim)(C.c_uint32(n), _p(m), _p(x), _p(node_m), _p(bw), _p(b))
        return node_m, bw, b

    # bvh.h:251-324
    def bvh_force(self, m, x, node_m, bw, G, theta, targets=None):
        n, dim = x.shape
        _, ct = _suffix(x.dtype, dim)
        f = self._fn("nbo_bvh_force", x.dtype, dim, C.c_uint64)
        if targets is None:
            a = np.empty_like(x)
            visits = f(C.c_uint32(n), ct(G), ct(theta), _p(m), _p(x), _p(node_m), _p(bw), C.c_uint32(0), None, _p(a))
        else:
            targets = np.ascontiguousarray(targets, np.uint32)
            a = np.empty((len(targets), dim), x.dtype)
            visits = f(C.c_uint32(n), ct(G), ct(theta), _p(m), _p(x), _p(node_m), _p(bw),
                       C.c_uint32(len(targets)), _p(targets), _p(a))
        return a, int(visits)

    # octree.h:77-224
    def octree_build(self, m, x):
        n, dim = x.shape
        cc = 1 << dim
        cap = max(cc * n, 1000)
        _, ct = _suffix(x.dtype, dim)
        fc = np.empty(cap, np.uint32)
        par = np.empty(1 + cap // cc, np.uint32)
        node_m = np.empty((cap, dim + 1), x.dtype)
        side = ct(0)
        root = np.empty(dim, x.dtype)
        used = self._fn("nbo_octree_build", x.dtype, dim, C.c_uint32)(
            C.c_uint32(n), _p(m), _p(x), C.c_uint32(cap), _p(fc), _p(par), _p(node_m), C.byref(side), _p(root))
        if used == 0:
            raise RuntimeError("octree capacity exceeded")
        return dict(used=int(used), side=x.dtype.type(side.value), root=root, first_child=fc[:used],
                    parent=par[:1 + used // cc], node_m=node_m[:used])

    def octree_canonical(self, tree, dim):
        fc = np.ascontiguousarray(tree["first_child"])
        node_m = np.ascontiguousarray(tree["node_m"])
        used = len(fc)
        depth = np.empty(used, np.uint32)
        path = np.empty(used, np.uint64)
        kind = np.empty(used, np.uint32)
        mo = np.empty((used, dim + 1), node_m.dtype)
        cnt = self._fn("nbo_octree_canonical", node_m.dtype, dim, C.c_uint32)(
            C.c_uint32(used), _p(fc), _p(depth), _p(path), _p(kind), _p(mo), _p(node_m))
        return depth[:cnt], path[:cnt], kind[:cnt], mo[:cnt]

    # octree.h:227-263
    def octree_force(self, x, tree, G, theta, targets=None):
        n, dim = x.shape
        _, ct = _suffix(x.dtype, dim)
        f = self._fn("nbo_octree_force", x.dtype, dim, C.c_uint64)
        fc = np.ascontiguousarray(tree["first_child"])
        par = np.ascontiguousarray(tree["parent"])
        nm = np.ascontiguousarray(tree["node_m"])
        if targets is None:
            a = np.empty_like(x)
            visits = f(C.c_uint32(n), ct(G), ct(theta), _p(x), ct(tree["side"]), _p(fc), _p(par), _p(nm),
                       C.c_uint32(0), None, _p(a))
        else:
            targets = np.ascontiguousarray(targets, np.uint32)
            a = np.empty((len(targets), dim), x.dtype)
            visits = f(C.c_uint32(n), ct(G), ct(theta), _p(x), ct(tree["side"]), _p(fc), _p(par), _p(nm),
                       C.c_uint32(len(targets)), _p(targets), _p(a))
        return a, int(visits)

    def octree_visits_nonempty(self, x, tree, theta):
        """(body, node) tests of the reference walk that hit a NON-EMPTY node (empty children contribute +0)."""
        n, dim = x.shape
        _, ct = _suffix(x.dtype, dim)
        f = self._fn("nbo_octree_visits_nonempty", x.dtype, dim, C.c_uint64)
        return int(f(C.c_uint32(n), ct(theta), _p(x), ct(tree["side"]), _p(np.ascontiguousarray(tree["first_child"])),
                     _p(np.ascontiguousarray(tree["parent"])), _p(np.ascontiguousarray(tree["node_m"]))))

    def permute(self, perm, s):
        n, dim = s["x"].shape
        out = {k: np.ascontiguousarray(s[k]).copy() for k in ("m", "x", "v", "a", "ao")}
        perm = np.ascontiguousarray(perm, np.uint32)
        self._fn("nbo_permute", s["x"].dtype, dim)(C.c_uint32(n), _p(perm), _p(out["m"]), _p(out["x"]), _p(out["v"]),
                                                  _p(out["a"]), _p(out["ao"]))
        return out

    # k x (force + accelerate_step)
    def run(self, algo, s, steps, theta=0.5):
        n, dim = s["x"].shape
        _, ct = _suffix(s["x"].dtype, dim)
        out = {k: np.ascontiguousarray(s[k]).copy() for k in ("m", "x", "v", "a", "ao")}
        rc = self._fn("nbo_run", s["x"].dtype, dim, C.c_int)(
            C.c_int(ALGOS[algo]), C.c_uint32(n), ct(s["dt"]), ct(s["G"]), ct(theta), C.c_uint32(steps),
            _p(out["m"]), _p(out["x"]), _p(out["v"]), _p(out["a"]), _p(out["ao"]))
        if rc != 0:
            raise RuntimeError("oracle run failed (octree capacity)")
        out["dt"], out["G"] = s["dt"], s["G"]
        return out


# ----------------------------------------------------------------------------------------------------------
# The compiled reference (oracle/_ref/refdump_d{2,3}); present in the build container and, as prebuilt
# binaries inside the gpurun snapshot, on the GPU box.  One subprocess per call (see refdump.cpp header).

def ref_available(dim=3) -> bool:
    return os.access(os.path.join(REF_DIR, f"refdump_d{dim}"), os.X_OK)


def _state_bytes(s):
    hdr = np.array([float(s["dt"]), float(s["G"])], np.float64).tobytes()
    return hdr + b"".join(np.ascontiguousarray(s[k]).tobytes() for k in ("m", "x", "v", "a", "ao"))


def _parse_state(buf, n, dim, dtype, off=0):
    dtype = np.dtype(dtype)
    dt, G = np.frombuffer(buf, np.float64, 2, off)
    off += 16
    out = {"dt": dtype.type(dt), "G": dtype.type(G)}
    for k, cnt in (("m", n), ("x", n * dim), ("v", n * dim), ("a", n * dim), ("ao", n * dim)):
        arr = np.frombuffer(buf, dtype, cnt, off).copy()
        off += cnt * dtype.itemsize
        out[k] = arr if k == "m" else arr.reshape(n, dim)
    return out, off


def avx512_host() -> bool:
    try:
        return "avx512f" in open("/proc/cpuinfo").read()
    except OSError:
        return False


def ref_native_available() -> bool:
    """refdump_native_d2 is built with -march=native in the build container (an AVX-512 host)."""
    return os.access(os.path.join(REF_DIR, "refdump_native_d2"), os.X_OK) and avx512_host()


def refdump(op, dtype, dim, n, state=None, theta=0.5, steps=1, raw_in=None, native=False) -> bytes:
    exe = os.path.join(REF_DIR, f"refdump_native_d{dim}" if native else f"refdump_d{dim}")
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        if state is not None:
            open(fin, "wb").write(_state_bytes(state))
        elif raw_in is not None:
            open(fin, "wb").write(raw_in)
        else:
            fin = "-"
        prec = "f" if np.dtype(dtype) == np.float32 else "d"
        subprocess.run([exe, op, prec, str(n), repr(float(theta)), str(steps), fin, fout], check=True)
        return open(fout, "rb").read()


def ref_galaxy(n, dtype, dim):
    buf = refdump("galaxy", dtype, dim, n)
    size = int(np.frombuffer(buf, np.uint32, 1, 0)[0])
    s, _ = _parse_state(buf, size, dim, dtype, 8)
    return s


def ref_state_op(op, s, theta=0.5, steps=1, native=False):
    n, dim = s["x"].shape
    buf = refdump(op, s["x"].dtype, dim, n, s, theta, steps, native=native)
    out, off = _parse_state(buf, n, dim, s["x"].dtype)
    return out, buf, off
