"""Experiment helper: the BVH walk at small n with partially filled warps (NBX_BVH_LANES = 32 | 16 | 8 bodies per warp) and
sector touches (NBX_BVH_TOUCH = 0 none | 1 right sibling | 2 left child | 3 both), one process, the same state for every
mode; the accelerations must stay bit-identical (neither changes a body's arithmetic). Prints ms per step (no L2 flush: the
small-n regime of the reference's sweep) and the walk phase."""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


MODES = [(l, t) for l in ("32", "16", "8") for t in ("0", "1")]


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [30_000, 100_000, 300_000, 1_000_000]
    ctx = bench.Ctx()
    os.environ["NBX_BVH_NB"] = "1"
    for n in sizes:
        for prec, dt in (("float", np.float32), ("double", np.float64)):
            cfg = argparse.Namespace(algorithm="bvh", precision=prec, dim=3, n=n, theta=0.5)
            s = bench.make_state(n, dt, 3)
            row, dig = [], set()
            for lanes, touch in MODES:
                os.environ["NBX_BVH_LANES"], os.environ["NBX_BVH_TOUCH"] = lanes, touch
                eng = ctx.new_engine(s, cfg)
                eng.step(3)
                eng.sync()
                ms = min(eng.step_timed(5) / 5 for _ in range(3))
                eng.set_phase_timing(True)
                eng.step_timed(1)
                walk = min(eng.step_timed(1) and eng.phase_ms()["traverse"] for _ in range(3))
                dig.add(hashlib.sha1(eng.download(("a",))["a"].tobytes()).hexdigest())
                eng.close()
                row.append(f"{lanes}/t{touch}: {ms:.3f} ({walk:.3f})")
            print(f"n={n} {prec} lanes/touch: step (walk) ms: " + " | ".join(row) + f" | a identical: {len(dig) == 1}", flush=True)


if __name__ == "__main__":
    main()
