"""Experiment helper: runs bench.py over a list of configs and prints one compact line each (raw JSON -> gpurun_out/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = {
    "c1": ["--algorithm", "all-pairs", "-n", "10000", "--dim", "2", "--precision", "float"],
    "c2": ["--algorithm", "all-pairs", "-n", "1000000", "--dim", "3", "--precision", "float"],
    "c3": ["--algorithm", "all-pairs-collapsed", "-n", "262144", "--dim", "3", "--precision", "double"],
    "c3f": ["--algorithm", "all-pairs-collapsed", "-n", "262144", "--dim", "3", "--precision", "float"],
    "apd": ["--algorithm", "all-pairs", "-n", "262144", "--dim", "3", "--precision", "double"],
    "c4": ["--algorithm", "octree", "-n", "10000000", "--dim", "3", "--precision", "double"],
    "c4f": ["--algorithm", "octree", "-n", "10000000", "--dim", "3", "--precision", "float"],
    "bvh10": ["--algorithm", "bvh", "-n", "10000000", "--dim", "3", "--precision", "float"],
    "bvh10d": ["--algorithm", "bvh", "-n", "10000000", "--dim", "3", "--precision", "double"],
    "c5": ["--algorithm", "bvh", "-n", "100000000", "--dim", "3", "--precision", "float"],
    "s_ap": ["--algorithm", "all-pairs", "-n", "100000", "--dim", "3", "--precision", "double"],
    "s_apc": ["--algorithm", "all-pairs-collapsed", "-n", "100000", "--dim", "3", "--precision", "double"],
    "s_oct": ["--algorithm", "octree", "-n", "100000", "--dim", "3", "--precision", "double"],
    "s_bvh": ["--algorithm", "bvh", "-n", "100000", "--dim", "3", "--precision", "double"],
    "s_oct1m": ["--algorithm", "octree", "-n", "1000000", "--dim", "3", "--precision", "double"],
    "s_bvh1m": ["--algorithm", "bvh", "-n", "1000000", "--dim", "3", "--precision", "double"],
    "oct1": ["--algorithm", "octree", "-n", "1000000", "--dim", "3", "--precision", "float"],
    "bvh1": ["--algorithm", "bvh", "-n", "1000000", "--dim", "3", "--precision", "float"],
}
names = sys.argv[1:] or list(CONFIGS)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for nm in names:
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu-baseline", "--no-e2e"] + CONFIGS[nm]
    r = subprocess.run(cmd, capture_output=True, text=True)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if not lines:
        print(nm, "FAILED", r.stderr[-600:].replace("\n", " | "))
        continue
    d = json.loads(lines[-1])
    open(os.path.join(ROOT, "gpurun_out", f"matrix_{nm}.json"), "w").write(lines[-1])
    ph = {k: round(v, 3) for k, v in d["config"]["phase_ms"].items() if v}
    rf = d.get("roofline") or {}
    print(f"{nm:7s} {d['config']['workload']:55s} {d['value']:10.2f} {d['unit']:14s} {d['ms_per_step']:10.3f} ms/step  frac={rf.get('frac')}  {ph}", flush=True)
