"""Experiment helper: time the BVH sort phase (keys + 8-pass radix sort + gather) at a few sizes."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg
nbx = _pkg.load().nbx
for n in (1_000_000, 10_000_000, 50_000_000):
    rng = np.random.default_rng(0)
    x = (rng.random((n, 3), dtype=np.float32) * 200 - 100)
    m = np.full(n, 1.0 / n, np.float32)
    z = np.zeros_like(x)
    with nbx.Engine(n, 3, np.float32, "bvh", 0.1, 1.0) as e:
        e.upload(m, x, z, z, z)
        e.set_phase_timing(True)
        best = 1e9
        for _ in range(4):
            e.bounding_box(); e.hilbert_sort(); e.sync()
            best = min(best, e.phase_ms()["sort"])
        keys, perm = e.bvh_keys()
        ok = bool((keys[perm][1:] >= keys[perm][:-1]).all())
    print(f"n={n}: sort phase {best:.3f} ms  ({n / best / 1e6:.2f} Gkeys/s, {n * 256 / best / 1e6:.0f} GB/s algorithmic) sorted={ok}", flush=True)
