"""Experiment helper: times nbx_calc_energies (the O(n^2) potential-energy sweep behind --save energy)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402

nbx = _pkg.load().nbx
orc = O.Oracle(fast=True)
for dt, n in ((np.float32, 262144), (np.float64, 262144), (np.float32, 1000000)):
    s = orc.galaxy(n, dt, 3)
    with nbx.Engine(n, 3, dt, "all-pairs", s["dt"], s["G"]) as e:
        e.upload_state(s)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            k, g = e.calc_energies()
            ts.append(1e3 * (time.perf_counter() - t0))
    dtm = min(ts) * 1e-3
    print(f"energies {np.dtype(dt).name} n={n}: best {1e3 * dtm:9.2f} ms  {n * (n - 1) / dtm / 1e9:8.1f} G ordered pairs/s  "
          f"all calls {[round(t, 1) for t in ts]} ms  k={k:.6e} g={g:.6e}", flush=True)
