"""Loader for the committed golden fixtures (tests/golden/*.npz, generated from the reference by
tests/golden/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("f32", 2), ("f32", 3), ("f64", 2), ("f64", 3)]
IDS = [f"{t}-{d}d" for t, d in CASES]
DT = {"f32": np.float32, "f64": np.float64}
N = 96
STEPS = 3
THETA = 0.5


def load(tag, dim):
    return dict(np.load(os.path.join(GOLDEN, f"galaxy_{tag}_d{dim}_n{N}.npz")))


def init_state(g):
    x = g["init_x"]
    return dict(m=g["init_m"], x=x, v=g["init_v"], a=np.zeros_like(x), ao=np.zeros_like(x),
                dt=x.dtype.type(g["dt"]), G=x.dtype.type(g["G"]))


def same(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def rel_err(a, ref):
    """per-body relative error |a-ref| / |ref| (vector norm), float64"""
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    num = np.linalg.norm(a - ref, axis=-1)
    den = np.linalg.norm(ref, axis=-1)
    return num / np.where(den > 0, den, 1.0)


def rms(e):
    return float(np.sqrt(np.mean(np.square(e))))
