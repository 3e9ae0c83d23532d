"""The CPU oracle against the committed golden vectors (generated from the reference itself). Runs anywhere."""
import numpy as np
import pytest

from golden_util import CASES, IDS, STEPS, THETA, init_state, load, same


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_galaxy_and_all_pairs(oracle, tag, dim):
    g = load(tag, dim)
    s0 = init_state(g)
    s = oracle.galaxy(len(s0["m"]), s0["x"].dtype, dim)
    for k in "mxv":
        assert same(s[k], s0[k]), k
    assert same(oracle.all_pairs_force(s0["m"], s0["x"], s0["G"]), g["a_all_pairs"])


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_bvh_pipeline(oracle, tag, dim):
    g = load(tag, dim)
    s0 = init_state(g)
    lo, hi = oracle.bbox(s0["x"])
    assert same(np.stack([lo, hi]), g["bbox"])
    keys = oracle.keys(s0["x"], lo, hi)
    assert same(keys, g["keys"])
    so = oracle.permute(oracle.sort_perm(keys), s0)
    assert same(so["x"], g["sorted_x"]) and same(so["m"], g["sorted_m"]) and same(so["v"], g["sorted_v"])
    nm, bw, b = oracle.bvh_build(so["m"], so["x"])
    assert same(nm, g["bvh_m"]) and same(bw, g["bvh_bw"]) and same(b, g["bvh_b"])
    for theta in (0.0, THETA):
        a, _ = oracle.bvh_force(so["m"], so["x"], nm, bw, s0["G"], theta)
        assert same(a, g[f"a_bvh_theta{theta}"])


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_octree_pipeline(oracle, tag, dim):
    g = load(tag, dim)
    s0 = init_state(g)
    t = oracle.octree_build(s0["m"], s0["x"])
    assert t["used"] == int(g["octree_used"]) and t["side"] == g["octree_side"] and same(t["root"], g["octree_root"])
    depth, path, kind, mo = oracle.octree_canonical(t, dim)
    assert same(depth, g["octree_depth"]) and same(path, g["octree_path"]) and same(kind, g["octree_kind"])
    assert same(mo, g["octree_m"])
    for theta in (0.0, THETA):
        a, _ = oracle.octree_force(s0["x"], t, s0["G"], theta)
        assert same(a, g[f"a_octree_theta{theta}"])


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
@pytest.mark.parametrize("algo,key", [("all-pairs", "all_pairs"), ("all-pairs-collapsed", "collapsed"),
                                      ("octree", "octree"), ("bvh", "bvh")])
def test_multi_step(oracle, tag, dim, algo, key):
    g = load(tag, dim)
    out = oracle.run(algo, init_state(g), STEPS, THETA)
    for k in ("x", "v", "a"):
        assert same(out[k], g[f"run_{key}_{k}"]), k


def test_hilbert_known_answers(oracle):
    import os
    from golden_util import GOLDEN
    ka = np.load(os.path.join(GOLDEN, "hilbert_known_answers.npz"))
    assert same(oracle.hilbert(ka["c2"]), ka["k2"])
    assert same(oracle.hilbert(ka["c3"]), ka["k3"])
    assert int(ka["k3"][6]) == 0x774DC24749504E4D and int(ka["k2"][4]) == 0x42484ACA4842406A


def test_theta0_trees_match_all_pairs(oracle):
    """README.md:122-129: theta=0 Barnes-Hut must reproduce all-pairs (here: to rounding, double precision)."""
    from golden_util import rel_err
    g = load("f64", 3)
    assert rel_err(g["a_octree_theta0.0"], g["a_all_pairs"]).max() < 1e-11
    # bvh output is Hilbert-permuted: compare through the sort permutation
    s0 = init_state(g)
    lo, hi = oracle.bbox(s0["x"])
    perm = oracle.sort_perm(oracle.keys(s0["x"], lo, hi))
    assert rel_err(g["a_bvh_theta0.0"], g["a_all_pairs"][perm]).max() < 1e-12


def test_fast_oracle_close_to_pinned(oracle, oracle_fast):
    """The -Ofast/OpenMP build (CPU baseline 'port') agrees with the pinned build to rounding."""
    from golden_util import rel_err
    s = oracle.galaxy(2000, np.float64, 3)
    a0 = oracle.all_pairs_force(s["m"], s["x"], s["G"])
    a1 = oracle_fast.all_pairs_force(s["m"], s["x"], s["G"])
    assert rel_err(a1, a0).max() < 1e-12
