"""N>1 path: world_size-2 gloo tests on CPU for the host-side sharding logic, plus (on a box with >= 2 GPUs) the real
NCCL consistency check of tests/multi_gpu_check.py."""
import os
import subprocess
import sys

import pytest

import _pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
nbx = _pkg.load().nbx

GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["NBX_ROOT"])
import _pkg
nbx = _pkg.load().nbx
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 1000003
lo, hi = nbx.shard_bounds(n, rank, world)
chunk = (n + world - 1) // world
# what the engine does on the GPU: every rank owns one chunk of a padded array and all-gathers it in place
mine = torch.full((chunk,), -1, dtype=torch.int64)
mine[: hi - lo] = torch.arange(lo, hi)
full = torch.empty(chunk * world, dtype=torch.int64)
dist.all_gather_into_tensor(full, mine)
got = full[full >= 0]
assert got.numel() == n and bool((got == torch.arange(n)).all()), "shards must tile [0, n) in rank order"
# rank 0 creates the NCCL unique id in the real path; here: the same broadcast plumbing with a stand-in payload
ids = [bytes(range(128)) if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
assert ids[0] == bytes(range(128))
dist.barrier()
dist.destroy_process_group()
print("GLOO_OK", rank)
'''


def torchrun(args, env=None, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29561"] + args
    e = dict(os.environ, NBX_ROOT=ROOT, **(env or {}))
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_gloo_world2_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    r = torchrun([str(script)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("GLOO_OK") == 2


def test_reference_arm_under_torchrun_prints_once():
    """bench.py --impl reference under torchrun: rank 0 alone runs and prints, the other rank exits 0 without work."""
    if not os.access(os.path.join(ROOT, "oracle", "_ref", "nbody_d3"), os.X_OK):
        pytest.skip("oracle/_ref not built")
    r = torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                  "-n", "3000"])
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1 and '"impl": "reference"' in lines[0]


@pytest.mark.gpu
@pytest.mark.skipif(nbx.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_multi_gpu_matches_single_gpu():
    r = torchrun([os.path.join(ROOT, "tests", "multi_gpu_check.py")], timeout=900)
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-4000:] + r.stderr[-4000:]
