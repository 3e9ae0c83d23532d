"""Experiment helper: the octree walk at small n with partially filled warps (NBX_OCT_LANES = 32 | 16 | 8 bodies per warp), one
process, the same state for every mode; the accelerations must stay bit-identical. Prints ms per step (no L2 flush) and the
walk phase. (The look-ahead sector touches measured with it — NBX_OCT_AHEAD — did not pay and are gone from the kernels.)"""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

MODES = [(l, "0") for l in ("32", "16", "8")]


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [10_000, 30_000, 100_000, 300_000]
    ctx = bench.Ctx()
    for n in sizes:
        for prec, dt in (("float", np.float32), ("double", np.float64)):
            cfg = argparse.Namespace(algorithm="octree", precision=prec, dim=3, n=n, theta=0.5)
            s = bench.make_state(n, dt, 3)
            row, dig = [], set()
            for lanes, ahead in MODES:
                os.environ["NBX_OCT_LANES"], os.environ["NBX_OCT_AHEAD"] = lanes, ahead
                eng = ctx.new_engine(s, cfg)
                eng.step(3)
                eng.sync()
                ms = min(eng.step_timed(5) / 5 for _ in range(3))
                eng.set_phase_timing(True)
                eng.step_timed(1)
                walk = min(eng.step_timed(1) and eng.phase_ms()["traverse"] for _ in range(3))
                dig.add(hashlib.sha1(eng.download(("a",))["a"].tobytes()).hexdigest())
                eng.close()
                row.append(f"{lanes}/a{ahead}: {ms:.3f} ({walk:.3f})")
            print(f"n={n} {prec} lanes/ahead: step (walk) ms: " + " | ".join(row) + f" | a identical: {len(dig) == 1}", flush=True)


if __name__ == "__main__":
    main()
