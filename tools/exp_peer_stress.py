"""Stress check of the peer-memory exchange of the BVH accelerations (NBX_PEER=1): many short steps at small n, where the
ranks drift apart most, compared bit for bit with a single-GPU run. Under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/exp_peer_stress.py"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    os.environ["NBX_PEER"] = "1"
    ctx = bench.Ctx()
    bad = 0
    for n, dt, dim, steps in ((50021, np.float32, 3, 40), (7001, np.float64, 2, 60), (300007, np.float32, 3, 12), (20011, np.float64, 3, 40)):
        cfg = argparse.Namespace(algorithm="bvh", precision="float" if dt == np.float32 else "double", dim=dim, n=n, theta=0.5)
        s = bench.make_state(n, dt, dim)
        with ctx.new_engine(s, cfg, multi=True) as e:
            e.step(steps)
            multi = e.download()
        with ctx.new_engine(s, cfg, multi=False) as e:
            e.step(steps)
            single = e.download()
        same = all(multi[k].tobytes() == single[k].tobytes() for k in ("m", "x", "v", "a", "ao"))
        bad += 0 if same else 1
        print(f"[rank {ctx.rank}/{ctx.world}] peer buffers in use: {getattr(ctx, 'peer', False)}; bvh n={n} {np.dtype(dt).name} {dim}-D "
              f"{steps} steps: {'bit-exact' if same else 'DIFFERS'}", flush=True)
    t = ctx.torch.tensor([bad], device="cuda")
    ctx.dist.all_reduce(t)
    if ctx.rank == 0:
        print("PEER_STRESS", "PASS" if int(t.item()) == 0 else "FAIL", flush=True)
    ctx.dist.barrier()
    ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
