"""Experiment: C5-shaped BVH step with the accelerations exchanged by the walk kernel itself (P2P stores into the peers'
arrays, NBX_PEER=1) vs one NCCL all-gather after the walk (NBX_PEER=0), on the same state, under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/exp_c5_peer.py [n]"""
import argparse
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    ctx = bench.Ctx()
    cfg = argparse.Namespace(algorithm="bvh", precision="float", dim=3, n=n, theta=0.5)
    s = bench.make_state(n, np.float32, 3)
    digests = {}
    for peer in ("1", "0", "1", "0"):
        os.environ["NBX_PEER"] = peer
        eng = ctx.new_engine(s, cfg)
        for _ in range(2):
            eng.step(1)
        eng.sync()
        ms = []
        for _ in range(3):
            ctx.flush_l2()
            ctx.barrier()
            ms.append(eng.step_timed(1))
        ctx.barrier()
        total = ctx.max_over_ranks(sum(ms)) / 3
        eng.set_phase_timing(True)
        eng.step_timed(1)
        ph = eng.phase_ms()
        a = eng.download(("a",))["a"]
        digests.setdefault(peer, hashlib.sha1(a.tobytes()).hexdigest())
        eng.close()
        if ctx.rank == 0:
            print(f"n={n} gpus={ctx.world} NBX_PEER={peer} (peer buffers in use: {getattr(ctx, 'peer', False) if peer == '1' else False}): "
                  f"{total:.2f} ms/step  phases (rank 0) { {k: round(v, 2) for k, v in ph.items() if v} }", flush=True)
    if ctx.rank == 0:
        print("accelerations identical with and without peer buffers:", digests["1"] == digests["0"], flush=True)
    if ctx.dist:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
