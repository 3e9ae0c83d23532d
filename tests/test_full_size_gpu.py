"""Size-independent properties at the remaining BASELINE.json full sizes (C2 / C4 / the n = 4 M and 10 M BVH cases live
in test_allpairs_gpu.py, test_octree_gpu.py and test_bvh_gpu.py):
  C3  all-pairs-collapsed 3-D double galaxy n = 262 144  (pair-parallel; components 0,1 only, a = (a - ao) + c*sum)
  C5  bvh 3-D float galaxy n = 100 M                     (one GPU holds the whole replicated problem: ~21 GB)"""
import numpy as np
import pytest

import _pkg
from golden_util import rel_err, rms

pytestmark = pytest.mark.gpu
nbx = _pkg.load().nbx


def test_c3_collapsed_262144(oracle_fast):
    n = 262_144
    s = oracle_fast.galaxy(n, np.float64, 3)
    with nbx.Engine(n, 3, np.float64, "all-pairs-collapsed", s["dt"], s["G"]) as e:
        e.upload_state(s)
        e.all_pairs_collapsed_force()          # a = (a - ao) + c * sum over j, with a = ao = 0 at entry
        a = e.download(("a",))["a"]
    assert np.isfinite(a).all()
    assert not a[:, 2].any()                   # the reference never accumulates component 2
    rng = np.random.default_rng(11)
    targets = np.sort(rng.choice(n, 64, replace=False)).astype(np.uint32)
    ref = oracle_fast.all_pairs_force(s["m"], s["x"], s["G"], targets=targets)
    err = rel_err(a[targets][:, :2], ref[:, :2])
    assert err.max() <= 1e-12, err.max()
    f = a[:, :2] * s["m"][:, None]             # Newton's third law: every pair term is exactly antisymmetric
    assert np.abs(f.sum(0)).max() <= 1e-12 * np.abs(f).sum(0).max()


def test_c5_bvh_100M(oracle_fast):
    n = 100_000_000
    s = oracle_fast.galaxy(n, np.float32, 3)
    with nbx.Engine(n, 3, np.float32, "bvh", s["dt"], s["G"], theta=0.5) as e:
        e.upload_state(s)
        lo, hi = e.bounding_box()
        e.hilbert_sort()
        keys, perm = e.bvh_keys()
        e.build_tree()
        e.bvh_compute_force()
        a = e.download(("a",))["a"]
        st = e.traversal_stats()
    olo, ohi = oracle_fast.bbox(s["x"])
    assert lo.tobytes() == olo.tobytes() and hi.tobytes() == ohi.tobytes()
    # permutation + sortedness (stable ties), and the keys of a 1 M-body sample bit-exact against the pinned arithmetic
    seen = np.zeros(n, np.bool_)
    seen[perm] = True
    assert seen.all()
    del seen
    ks = keys[perm]
    assert (ks[1:] >= ks[:-1]).all()
    eq = ks[1:] == ks[:-1]
    assert (perm[1:][eq] > perm[:-1][eq]).all()
    del ks, eq
    rng = np.random.default_rng(12)
    sample = np.sort(rng.choice(n, 1_000_000, replace=False))
    from oracle import oracle as O
    pinned = O.Oracle(fast=False)  # -O2, no contraction: the build whose float keys are the reference's
    assert keys[sample].tobytes() == pinned.keys(np.ascontiguousarray(s["x"][sample]), olo, ohi).tobytes()
    # accelerations of 256 sampled bodies against the direct sum in double: Barnes-Hut error at theta = 0.5, not rounding
    assert np.isfinite(a).all()
    slots = np.sort(rng.choice(n, 256, replace=False))          # sorted slots; the state stays Hilbert-permuted
    bodies = perm[slots].astype(np.uint32)
    truth = oracle_fast.all_pairs_force_truth(s["m"], s["x"], s["G"], targets=bodies)
    err = rel_err(a[slots], truth)
    assert np.median(err) <= 2e-3 and err.max() <= 5e-2, (np.median(err), err.max())
    assert 5000 <= st["node_visits"] / n <= 20000                # ~12 k node tests per body estimated
