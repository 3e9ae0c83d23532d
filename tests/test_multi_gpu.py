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


# The data flow of the multi-GPU sharded sort (csrc/nbx_sort.cu sort_pairs_sharded), restated with torch + gloo: evenly spaced
# sample -> sorted -> world-1 splitters at the quantiles; owner = number of splitters <= key; stable partition by owner;
# every rank sorts (stably) only its own key range; segments exchanged with one broadcast per owner. The result must be the
# global STABLE sort permutation whatever the key distribution (ties never straddle two owners).
SHARDED_SORT_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)  # same keys on every rank, like the replicated Hilbert keys
for n, dist_name in ((300001, "uniform"), (262144, "ties"), (500000, "clustered")):
    if dist_name == "uniform":
        keys = rng.integers(0, 2**63, n, dtype=np.uint64)
    elif dist_name == "ties":
        keys = rng.integers(0, 50, n, dtype=np.uint64) << np.uint64(40)   # only 50 distinct keys: many owners may be empty
    else:
        keys = (rng.normal(2**40, 2**20, n).clip(0, 2**62)).astype(np.uint64)
    S = 65536
    sample = np.sort(keys[(np.arange(S, dtype=np.uint64) * np.uint64(n) // np.uint64(S)).astype(np.int64)], kind="stable")
    split = sample[[q * (S // world) for q in range(1, world)]]
    owner = (split[None, :] <= keys[:, None]).sum(1)
    part = np.argsort(owner, kind="stable")                                # the partition sweep: stable, by owner
    counts = np.bincount(owner, minlength=world)
    offs = np.concatenate([[0], np.cumsum(counts)])
    mine = part[offs[rank]:offs[rank + 1]]
    perm = np.full(n, -1, np.int64)
    perm[offs[rank]:offs[rank + 1]] = mine[np.argsort(keys[mine], kind="stable")]   # this rank sorts its key range only
    t = torch.from_numpy(perm)
    for q in range(world):                                                  # grouped broadcast, one per owner
        if counts[q]:
            seg = t[offs[q]:offs[q + 1]].clone()
            dist.broadcast(seg, src=q)
            t[offs[q]:offs[q + 1]] = seg
    want = np.argsort(keys, kind="stable")
    assert np.array_equal(t.numpy(), want), (dist_name, rank)
dist.barrier()
dist.destroy_process_group()
print("SHARDED_SORT_OK", rank)
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


def test_gloo_world2_sharded_sort_equals_global_stable_sort(tmp_path):
    script = tmp_path / "sharded_sort_worker.py"
    script.write_text(SHARDED_SORT_WORKER)
    r = torchrun([str(script)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("SHARDED_SORT_OK") == 2


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


@pytest.mark.gpu
@pytest.mark.skipif(nbx.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("algo,prec", [("all-pairs", "float"), ("bvh", "float"), ("octree", "double"), ("all-pairs-collapsed", "double")])
def test_cpp_driver_gpus2_prints_the_single_gpu_state(algo, prec):
    """The C++ host driver's `--gpus 2` (one engine + host thread per GPU inside one process): the printed final state —
    positions, velocities AND accelerations of rank 0's replicated copy — equals the single-GPU run's (trees bit-exact;
    all-pairs may differ in the last printed digit on a couple of lines)."""
    exe = os.path.join(ROOT, "stdpar-nbody_b200", "bin", "nbody_d3")
    n = "300000" if algo == "bvh" else "2000"  # bvh: large enough for the sharded sort
    args = ["-n", n, "-s", "12", "--workload", "galaxy", "--algorithm", algo, "--precision", prec, "--theta", "0.5", "--print-state"]
    one = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
    two = subprocess.run([exe] + args + ["--gpus", "2"], capture_output=True, text=True, timeout=600)
    assert one.returncode == 0 and two.returncode == 0, one.stderr + two.stderr
    def strip(out):  # NCCL announces its version on stdout when a communicator is first created
        return [ln for ln in out.splitlines() if not ln.startswith(("Total time", "NCCL version"))]
    a, b = strip(one.stdout), strip(two.stdout)
    assert len(a) == len(b) and len(a) > int(n)
    bad = [(x, y) for x, y in zip(a, b) if x != y]
    assert len(bad) <= (0 if algo in ("bvh", "octree") else 2), bad[:3]
