"""Multi-GPU consistency check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu_check.py
Every rank runs its shard of the targets through libnbx; afterwards each rank compares its (replicated) state with a
single-GPU run of the same problem done on its own device. Trees must agree bit-for-bit (the per-body arithmetic does
not depend on the sharding); all-pairs agrees to rounding (the j-split of the partial sums depends on the shard size)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402

nbx = _pkg.load().nbx


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    orc = O.Oracle(fast=True)
    ok = True
    for algo, dt, n, dim in (("all-pairs", np.float32, 9001, 3), ("all-pairs", np.float32, 70001, 3), ("all-pairs-collapsed", np.float64, 3001, 3),
                             ("bvh", np.float32, 50021, 3), ("octree", np.float64, 40009, 3), ("bvh", np.float64, 7001, 2),
                             ("octree", np.float32, 30011, 3)):
        # the last case runs the octree walk with its lanes in Hilbert order (the default only from n = 4 M)
        os.environ["NBX_OCT_HILBERT"] = "1" if (algo, dt) == ("octree", np.float32) else "0"
        s = orc.galaxy(n, dt, dim)
        n = len(s["m"])
        with nbx.Engine(n, dim, dt, algo, s["dt"], s["G"], device=local, rank=rank, world_size=world) as e:
            ids = [nbx.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            e.comm_init_rank(ids[0])
            e.upload_state(s)
            e.step(3)
            multi = e.download()
        with nbx.Engine(n, dim, dt, algo, s["dt"], s["G"], device=local) as e:
            e.upload_state(s)
            e.step(3)
            single = e.download()
        lo, hi = nbx.shard_bounds(n, rank, world)
        keys = ("x", "v", "a", "ao") if algo in ("bvh", "octree") else ("x",)
        for k in keys:
            a, b = multi[k].astype(np.float64), single[k].astype(np.float64)
            if algo in ("bvh", "octree"):
                good = multi[k].tobytes() == single[k].tobytes()
                detail = "bit-exact" if good else f"max abs diff {np.abs(a - b).max():.3e}"
            else:
                err = np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)
                good = err.max() < (1e-5 if dt == np.float32 else 1e-12)
                detail = f"max rel diff {err.max():.2e}"
            ok &= bool(good)
            if rank == 0 or not good:
                print(f"[rank {rank}/{world}] {algo} {np.dtype(dt).name} {dim}-D n={n} {k}: {'OK' if good else 'FAIL'} ({detail})",
                      flush=True)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(t.item()) else "FAIL", flush=True)
    sys.exit(0 if int(t.item()) else 1)


if __name__ == "__main__":
    main()
