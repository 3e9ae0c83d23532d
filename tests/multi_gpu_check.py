"""Multi-GPU consistency check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu_check.py [--full]
Every rank runs its shard of the force work through libnbx; afterwards each rank compares its (replicated) state — x, v, a
and ao of ALL bodies — with a single-GPU run of the same problem done on its own device. Trees must agree bit-for-bit
(keys, permutation and the per-body arithmetic do not depend on the sharding); all-pairs agrees to rounding (the partial
sums are grouped differently). The cases and the comparison are bench.py's `verify` leg (bench.verify_multi), so the
driver's N > 1 bench lines carry exactly this check; `--full` adds the C2- and C5-shaped cases (n = 1 M all-pairs,
n = 10 M bvh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ctx = bench.Ctx()
    cases = bench.VERIFY_CASES if "--full" in sys.argv else bench.VERIFY_CASES[:7]
    v = bench.verify_multi(ctx, cases, log=lambda m: print(m, flush=True))
    if ctx.dist:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    if ctx.rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if v["ok"] else "FAIL", flush=True)
    sys.exit(0 if v["ok"] else 1)


if __name__ == "__main__":
    main()
