"""Experiment helper: BVH walk time for 1 / 2 / 4 bodies per lane (NBX_BVH_NB) at several sizes and both precisions."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sizes = [int(a) for a in sys.argv[1:]] or [100_000, 1_000_000]
for n in sizes:
    for prec in ("float", "double"):
        row = []
        for nb in (1, 2, 4):
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--algorithm", "bvh", "-n", str(n), "--precision", prec,
                                "--steps", "3", "--warmup", "2", "--no-cpu-baseline", "--no-e2e", "--no-configs"], capture_output=True,
                               text=True, env=dict(os.environ, NBX_BVH_NB=str(nb)))
            d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
            row.append(f"NB={nb}: step {d['ms_per_step']:.3f} walk {d['config']['phase_ms']['traverse']:.3f}")
        print(f"n={n} {prec}: " + " | ".join(row), flush=True)
