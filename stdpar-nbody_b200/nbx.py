"""ctypes binding of libnbx.so (include/nbx.h) — the reference-side stub a Python caller would use, and what the
parity tests and bench.py drive. It holds NO compute: every method is one C-ABI call. If the CUDA library is
missing it raises; there is no CPU path.

The method names mirror the reference's per-phase functions (src/all_pairs.h, src/bvh.h, src/octree.h,
src/system.h) so the parity tests read like calls into the reference.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NBX_LIB_PATH") or os.path.join(HERE, "lib", "libnbx.so")

ALL_PAIRS, ALL_PAIRS_COLLAPSED, OCTREE, BVH = 0, 1, 2, 3
ALGORITHMS = {"all-pairs": ALL_PAIRS, "all-pairs-collapsed": ALL_PAIRS_COLLAPSED, "octree": OCTREE, "bvh": BVH}
F32, F64 = 4, 8
FLAG_COLLAPSED_FIX_Z = 0x1
FLAG_NO_FUSED_INTEGRATE = 0x2
FLAG_ALLPAIRS_ORDERED = 0x4
FLAG_ALLPAIRS_SYMMETRIC = 0x8
UNIQUE_ID_BYTES = 128
PEER_HANDLE_BYTES = 128
PHASES = ("force", "accel", "bbox", "sort", "build", "multipoles", "traverse", "comm")

# every symbol include/nbx.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "nbx_last_error", "nbx_version", "nbx_device_count", "nbx_create", "nbx_destroy", "nbx_upload", "nbx_download", "nbx_upload_shard", "nbx_download_shard",
    "nbx_step", "nbx_step_timed", "nbx_sync", "nbx_all_pairs_force", "nbx_all_pairs_collapsed_force",
    "nbx_accelerate_step", "nbx_calc_energies", "nbx_bvh_bounding_box", "nbx_bvh_hilbert_sort", "nbx_bvh_build_tree",
    "nbx_bvh_compute_force", "nbx_bvh_get_keys", "nbx_bvh_get_nodes", "nbx_octree_build", "nbx_octree_compute_force",
    "nbx_octree_get_root", "nbx_octree_get_canonical", "nbx_stream_positions_begin", "nbx_stream_positions_end",
    "nbx_comm_unique_id", "nbx_comm_init_rank", "nbx_peer_export", "nbx_peer_import",
    "nbx_measure_fma_peak", "nbx_traversal_stats", "nbx_walk_width", "nbx_get_counters", "nbx_set_phase_timing", "nbx_get_phase_ms",
]


class NbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nbx error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("dim", C.c_int32), ("precision", C.c_int32), ("algorithm", C.c_int32),
                ("n", C.c_uint32), ("device", C.c_int32), ("dt", C.c_double), ("G", C.c_double), ("theta", C.c_double),
                ("rank", C.c_int32), ("world_size", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


_lib = None


def lib():
    """Load libnbx.so (fails loudly when it has not been built: the product has no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {HERE}` (or __graft_entry__.build())")
        _lib = C.CDLL(LIB_PATH)
        _lib.nbx_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            getattr(_lib, name)  # AttributeError if the library does not export what the header declares
    return _lib


def _check(rc):
    if rc != 0:
        raise NbxError(rc, lib().nbx_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count() -> int:
    return int(lib().nbx_device_count())


def measure_fma_peak(precision=F32, device=0) -> float:
    out = C.c_double(0)
    _check(lib().nbx_measure_fma_peak(C.c_int(device), C.c_int(precision), C.byref(out)))
    return out.value


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    _check(lib().nbx_comm_unique_id(buf))
    return buf.raw


class Engine:
    """One engine = one System<T,N> resident on one GPU (plus its tree)."""

    def __init__(self, n, dim, dtype, algorithm, dt, G, theta=0.5, device=0, rank=0, world_size=1, flags=0):
        self.dtype = np.dtype(dtype)
        self.dim, self.n = int(dim), int(n)
        algo = ALGORITHMS[algorithm] if isinstance(algorithm, str) else int(algorithm)
        cfg = Config(C.sizeof(Config), self.dim, self.dtype.itemsize, algo, self.n, device, float(dt), float(G),
                     float(theta), rank, world_size, flags, 0)
        self._h = C.c_void_p()
        _check(lib().nbx_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().nbx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- state_t hand-off (src/system.h:41-50) -----------------------------------------------------------------
    def _arr(self, a, vec):
        if a is None:
            return None
        a = np.ascontiguousarray(a, self.dtype)
        assert a.shape == ((self.n, self.dim) if vec else (self.n,)), a.shape
        return a

    def upload(self, m=None, x=None, v=None, a=None, ao=None):
        m, x, v, a, ao = self._arr(m, 0), self._arr(x, 1), self._arr(v, 1), self._arr(a, 1), self._arr(ao, 1)
        _check(lib().nbx_upload(self._h, _p(m), _p(x), _p(v), _p(a), _p(ao)))

    def upload_shard(self, m=None, x=None, v=None, a=None, ao=None):
        """Collective: every rank copies only its shard of the (full) host arrays; the shards are all-gathered on the device."""
        m, x, v, a, ao = self._arr(m, 0), self._arr(x, 1), self._arr(v, 1), self._arr(a, 1), self._arr(ao, 1)
        _check(lib().nbx_upload_shard(self._h, _p(m), _p(x), _p(v), _p(a), _p(ao)))

    def download_shard_into(self, m=None, x=None, v=None, a=None, ao=None):
        """Writes this rank's shard of the bodies into the given full-size host arrays (other entries are left alone)."""
        _check(lib().nbx_download_shard(self._h, _p(m), _p(x), _p(v), _p(a), _p(ao)))

    def upload_state(self, s):
        self.upload(s["m"], s["x"], s["v"], s["a"], s["ao"])

    def download(self, which=("m", "x", "v", "a", "ao")):
        out = {k: np.empty((self.n,) if k == "m" else (self.n, self.dim), self.dtype) for k in which}
        _check(lib().nbx_download(self._h, *[_p(out.get(k)) for k in ("m", "x", "v", "a", "ao")]))
        return out

    # ---- drivers ------------------------------------------------------------------------------------------------
    def step(self, steps=1):
        _check(lib().nbx_step(self._h, C.c_uint32(steps)))

    def step_timed(self, steps=1) -> float:
        ms = C.c_float(0)
        _check(lib().nbx_step_timed(self._h, C.c_uint32(steps), C.byref(ms)))
        return ms.value

    def sync(self):
        _check(lib().nbx_sync(self._h))

    # ---- phases -------------------------------------------------------------------------------------------------
    def all_pairs_force(self):
        _check(lib().nbx_all_pairs_force(self._h))

    def all_pairs_collapsed_force(self):
        _check(lib().nbx_all_pairs_collapsed_force(self._h))

    def accelerate_step(self):
        _check(lib().nbx_accelerate_step(self._h))

    def calc_energies(self):
        k, g = C.c_double(0), C.c_double(0)
        _check(lib().nbx_calc_energies(self._h, C.byref(k), C.byref(g)))
        return k.value, g.value

    def bounding_box(self):
        lo, hi = np.empty(self.dim, self.dtype), np.empty(self.dim, self.dtype)
        _check(lib().nbx_bvh_bounding_box(self._h, _p(lo), _p(hi)))
        return lo, hi

    def hilbert_sort(self):
        _check(lib().nbx_bvh_hilbert_sort(self._h))

    def bvh_keys(self):
        keys, perm = np.empty(self.n, np.uint64), np.empty(self.n, np.uint32)
        _check(lib().nbx_bvh_get_keys(self._h, _p(keys), _p(perm)))
        return keys, perm

    def build_tree(self):
        _check(lib().nbx_bvh_build_tree(self._h))

    def bvh_nodes(self):
        nn = C.c_uint64(0)
        _check(lib().nbx_bvh_get_nodes(self._h, C.byref(nn), None, None, None))
        nn = nn.value
        node_m = np.empty((nn, self.dim + 1), self.dtype)
        bw = np.empty(nn, self.dtype)
        b = np.empty((nn, 2, self.dim), self.dtype)
        cnt = C.c_uint64(0)
        _check(lib().nbx_bvh_get_nodes(self._h, C.byref(cnt), _p(node_m), _p(bw), _p(b)))
        return node_m, bw, b

    def bvh_compute_force(self):
        _check(lib().nbx_bvh_compute_force(self._h))

    def octree_build(self):
        _check(lib().nbx_octree_build(self._h))

    def octree_compute_force(self):
        _check(lib().nbx_octree_compute_force(self._h))

    def octree_root(self):
        side = np.empty(1, self.dtype)
        root = np.empty(self.dim, self.dtype)
        used = C.c_uint64(0)
        _check(lib().nbx_octree_get_root(self._h, _p(side), _p(root), C.byref(used)))
        return side[0], root, used.value

    def octree_canonical(self):
        cnt = C.c_uint64(0)
        _check(lib().nbx_octree_get_canonical(self._h, C.byref(cnt), None, None, None, None))
        k = cnt.value
        depth, path, kind = np.empty(k, np.uint32), np.empty(k, np.uint64), np.empty(k, np.uint32)
        mono = np.empty((k, self.dim + 1), self.dtype)
        _check(lib().nbx_octree_get_canonical(self._h, C.byref(cnt), _p(depth), _p(path), _p(kind), _p(mono)))
        return depth, path, kind, mono

    # ---- Saver streaming (src/saving.h:110-114) -------------------------------------------------------------------
    def stream_positions_begin(self):
        _check(lib().nbx_stream_positions_begin(self._h))

    def stream_positions_end(self):
        x = np.empty((self.n, self.dim), self.dtype)
        _check(lib().nbx_stream_positions_end(self._h, _p(x)))
        return x

    # ---- multi-GPU ----------------------------------------------------------------------------------------------
    def comm_init_rank(self, unique_id: bytes):
        assert len(unique_id) == UNIQUE_ID_BYTES
        _check(lib().nbx_comm_init_rank(self._h, C.c_char_p(unique_id)))

    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        _check(lib().nbx_peer_export(self._h, buf))
        return buf.raw

    def peer_import(self, handles_in_rank_order):
        """None: drop the peer buffers (back to the NCCL all-gather)."""
        if handles_in_rank_order is None:
            _check(lib().nbx_peer_import(self._h, None))
            return
        blob = b"".join(handles_in_rank_order)
        _check(lib().nbx_peer_import(self._h, C.c_char_p(blob)))

    # ---- measurement --------------------------------------------------------------------------------------------
    def counters(self):
        k, h, d = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().nbx_get_counters(self._h, C.byref(k), C.byref(h), C.byref(d)))
        return dict(kernel_launches=k.value, h2d_bytes=h.value, d2h_bytes=d.value)

    def traversal_stats(self):
        v, a, w = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().nbx_traversal_stats(self._h, C.byref(v), C.byref(a), C.byref(w)))
        width = C.c_uint32(0)
        _check(lib().nbx_walk_width(self._h, C.byref(width)))
        return dict(node_visits=v.value, interactions=a.value, warp_steps=w.value, width=width.value)

    def set_phase_timing(self, enable=True):
        _check(lib().nbx_set_phase_timing(self._h, C.c_int(1 if enable else 0)))

    def phase_ms(self):
        arr = (C.c_float * len(PHASES))()
        cnt = C.c_int(0)
        _check(lib().nbx_get_phase_ms(self._h, arr, C.c_int(len(PHASES)), C.byref(cnt)))
        return {PHASES[i]: float(arr[i]) for i in range(cnt.value)}


def shard_bounds(n, rank, world_size):
    """Targets owned by `rank`: [rank*ceil(n/W), (rank+1)*ceil(n/W)) clipped to n (same rule as nbx_create)."""
    chunk = (n + world_size - 1) // world_size
    return min(chunk * rank, n), min(chunk * (rank + 1), n)
