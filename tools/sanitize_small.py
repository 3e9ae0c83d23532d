"""Small end-to-end pass over every engine (all algorithms x float/double x 2-D/3-D, odd and power-of-two sizes): a quick
"does every path run and stay finite" check on the GPU box, and the workload to put under
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py
where that tool is available (it is closed on this round's pool; the plain run passed)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402

nbx = _pkg.load().nbx
orc = O.Oracle(fast=True)
for algo in ("all-pairs", "all-pairs-collapsed", "octree", "bvh"):
    for dt in (np.float32, np.float64):
        for dim in (2, 3):
            for n in (2, 33, 1024, 3001, 20000):
                if algo.startswith("all-pairs") and n == 20000 and dt == np.float64 and dim == 2:
                    continue
                s = orc.galaxy(n, dt, dim)
                with nbx.Engine(len(s["m"]), dim, dt, algo, s["dt"], s["G"], theta=0.5) as e:
                    e.upload_state(s)
                    e.step(2)
                    if algo in ("octree", "bvh"):
                        e.traversal_stats()
                    out = e.download()
                assert np.isfinite(out["x"]).all(), (algo, dt, dim, n)
    print(algo, "ok", flush=True)
print("SANITIZE_SMALL PASS")
