"""Experiment helper: the ordered all-pairs kernel with FP32x2 arithmetic over pairs of targets (NBX_AP_PACKED=1) against the
scalar one — results must be bit-identical (same operations per target), then the C1 timing under the target blockings."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402

nbx = _pkg.load().nbx
orc = O.Oracle(fast=True)
ok = True
for n, dim in ((10000, 2), (5000, 3), (12345, 3)):
    s = orc.galaxy(n, np.float32, dim)
    res = {}
    for ti in ("2", "4"):
        for pk in ("0", "1"):
            os.environ["NBX_AP_TI"], os.environ["NBX_AP_PACKED"] = ti, pk
            with nbx.Engine(len(s["m"]), dim, np.float32, "all-pairs", s["dt"], s["G"], flags=nbx.FLAG_ALLPAIRS_ORDERED) as e:
                e.upload_state(s)
                e.step(2)
                res[(ti, pk)] = e.download()
        same = all(res[(ti, "0")][k].tobytes() == res[(ti, "1")][k].tobytes() for k in ("x", "v", "a"))
        ok &= same
        print(f"n={n} dim={dim} TI={ti}: packed == scalar bitwise: {same}", flush=True)
print("EXP_AP_PACKED", "PASS" if ok else "FAIL", flush=True)
for ti, pk in (("", "0"), ("2", "0"), ("2", "1"), ("4", "0"), ("4", "1")):
    env = dict(os.environ, NBX_AP_PACKED=pk)
    env.pop("NBX_AP_TI", None)
    if ti:
        env["NBX_AP_TI"] = ti
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_matrix.py"), "c1"], capture_output=True, text=True, env=env)
    print(f"TI={ti or 'auto'} packed={pk}: {r.stdout.strip()[:140]}", flush=True)
