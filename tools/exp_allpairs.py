"""Experiment helper (not part of the product): time + accuracy of the all-pairs kernel variants selected by env vars."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402
from golden_util import rel_err, rms  # noqa: E402

nbx = _pkg.load().nbx
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dt = np.float32 if (len(sys.argv) <= 3 or sys.argv[3] == "float") else np.float64
orc = O.Oracle(fast=True)
s = orc.galaxy(n, dt, dim)
rng = np.random.default_rng(0)
targets = np.sort(rng.choice(n, 256, replace=False)).astype(np.uint32)
truth = orc.all_pairs_force_truth(s["m"], s["x"], s["G"], targets=targets)
flags = int(os.environ.get("EXP_FLAGS", "0"))
with nbx.Engine(n, dim, dt, "all-pairs", s["dt"], s["G"], flags=flags) as e:
    e.upload_state(s)
    e.all_pairs_force()
    a = e.download(("a",))["a"]
    err = rel_err(a[targets], truth)
    e.set_phase_timing(True)
    best = 1e9
    for _ in range(5):
        e.step_timed(1)
        best = min(best, e.phase_ms()["force"])
pairs = n * (n - 1)
print(f"env={ {k: v for k, v in os.environ.items() if k.startswith('NBX_')} } n={n} dim={dim} {np.dtype(dt).name}: "
      f"{best:.3f} ms  {pairs / best / 1e6:.1f} Gpairs/s  err rms={rms(err):.2e} max={err.max():.2e} finite={np.isfinite(a).all()}")
