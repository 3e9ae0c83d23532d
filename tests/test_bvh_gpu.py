"""GPU parity of the BVH pipeline (bbox -> Hilbert keys -> radix sort -> permute -> build -> traversal) through the
C ABI. Integer/index artefacts and tree nodes are BIT-EXACT against the oracle; accelerations are within tolerance
(the traversal takes the same accept/open decisions as the reference, only the accumulation differs by rounding):
  float : per-body rel. error vs the pinned float oracle rms <= 2e-5, max <= 5e-4;  double: max <= 1e-11.
"""
import numpy as np
import pytest

import _pkg
from golden_util import CASES, DT, IDS, STEPS, THETA, init_state, load, rel_err, rms, same

pytestmark = pytest.mark.gpu
nbx = _pkg.load().nbx

TOL = {np.dtype(np.float32): (2e-5, 5e-4), np.dtype(np.float64): (1e-12, 1e-11)}


def engine(s, theta=THETA, algo="bvh"):
    n, dim = s["x"].shape
    e = nbx.Engine(n, dim, s["x"].dtype, algo, s["dt"], s["G"], theta=theta)
    e.upload_state(s)
    return e


def check_pipeline(oracle, s, thetas=(0.0, THETA)):
    """Runs every phase and checks each artefact against the oracle fed the same bytes."""
    dt = s["x"].dtype
    lo, hi = oracle.bbox(s["x"])
    keys = oracle.keys(s["x"], lo, hi)
    perm = oracle.sort_perm(keys)
    so = oracle.permute(perm, s)
    nm, bw, b = oracle.bvh_build(so["m"], so["x"])
    with engine(s) as e:
        glo, ghi = e.bounding_box()
        assert same(glo, lo) and same(ghi, hi)
        e.hilbert_sort()
        gkeys, gperm = e.bvh_keys()
        assert same(gkeys, keys), "Hilbert keys must be bit-exact"
        assert same(gperm, perm), "sort permutation must be exact (stable)"
        st = e.download()
        for k in ("m", "x", "v", "a", "ao"):
            assert same(st[k], so[k]), k
        e.build_tree()
        gm, gbw, gb = e.bvh_nodes()
        assert same(gm, nm) and same(gbw, bw) and same(gb, b), "BVH nodes must be bit-exact"
    for theta in thetas:
        ref, _ = oracle.bvh_force(so["m"], so["x"], nm, bw, s["G"], theta)
        with engine(s, theta) as e:
            e.bounding_box(); e.hilbert_sort(); e.build_tree(); e.bvh_compute_force()
            a = e.download(("a",))["a"]
        err = rel_err(a, ref)
        tr, tm = TOL[dt]
        assert rms(err) <= tr and err.max() <= tm, (theta, rms(err), err.max())
    return perm


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_pipeline_vs_golden(oracle, tag, dim):
    g = load(tag, dim)
    s = init_state(g)
    with engine(s) as e:
        lo, hi = e.bounding_box()
        assert same(np.stack([lo, hi]), g["bbox"])
        e.hilbert_sort()
        keys, perm = e.bvh_keys()
        assert same(keys, g["keys"])
        st = e.download()
        assert same(st["x"], g["sorted_x"]) and same(st["m"], g["sorted_m"]) and same(st["v"], g["sorted_v"])
        e.build_tree()
        nm, bw, b = e.bvh_nodes()
        assert same(nm, g["bvh_m"]) and same(bw, g["bvh_bw"]) and same(b, g["bvh_b"])
    tr, tm = TOL[s["x"].dtype]
    for theta in (0.0, THETA):
        with engine(s, theta) as e:
            e.bounding_box(); e.hilbert_sort(); e.build_tree(); e.bvh_compute_force()
            a = e.download(("a",))["a"]
        err = rel_err(a, g[f"a_bvh_theta{theta}"])
        assert rms(err) <= tr and err.max() <= tm, (theta, rms(err), err.max())


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
@pytest.mark.parametrize("n", [2, 3, 1000, 4097])
def test_pipeline_vs_oracle(oracle, tag, dim, n):
    check_pipeline(oracle, oracle.galaxy(n, DT[tag], dim))


def test_pipeline_vs_oracle_65536(oracle):
    # the pinned (-O2, no contraction) oracle: the -Ofast build changes float keys
    check_pipeline(oracle, oracle.galaxy(65536, np.float32, 3), thetas=(THETA,))


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_steps_vs_golden(tag, dim):
    g = load(tag, dim)
    s = init_state(g)
    with engine(s) as e:
        e.step(STEPS)
        out = e.download()
    tol = 2e-5 if tag == "f32" else 1e-11
    # the state is left in the Hilbert order of the last step, exactly like the reference
    assert rms(rel_err(out["x"], g["run_bvh_x"])) <= tol
    assert rms(rel_err(out["v"], g["run_bvh_v"])) <= 20 * tol


def test_theta0_equals_all_pairs(oracle):
    """README.md:122-129."""
    s = oracle.galaxy(3000, np.float64, 3)
    with engine(s, theta=0.0) as e:
        e.bounding_box(); e.hilbert_sort(); e.build_tree(); e.bvh_compute_force()
        _, perm = e.bvh_keys()
        a = e.download(("a",))["a"]
    with engine(s, algo="all-pairs") as e:
        e.all_pairs_force()
        ap = e.download(("a",))["a"]
    assert rel_err(a, ap[perm]).max() < 1e-11


def test_ties_are_stable(oracle):
    """Duplicate positions give duplicate keys: the radix sort keeps index order (the reference's std::sort leaves
    ties undefined)."""
    s = oracle.galaxy(2048, np.float32, 3)
    s["x"][100:200] = s["x"][7]
    with engine(s) as e:
        e.bounding_box(); e.hilbert_sort()
        keys, perm = e.bvh_keys()
    assert same(perm, oracle.sort_perm(keys))
    tied = perm[np.isin(perm, np.r_[7, 100:200])]
    assert (np.diff(tied.astype(np.int64)) > 0).all()


def test_sort_properties_10M():
    """Size-independent properties at a BASELINE-scale n: perm is a permutation, keys come out sorted, ties in order."""
    n = 10_000_000
    rng = np.random.default_rng(5)
    x = (rng.random((n, 3), dtype=np.float32) * 200 - 100).astype(np.float32)
    s = dict(m=np.full(n, 1.0 / n, np.float32), x=x, v=np.zeros_like(x), a=np.zeros_like(x), ao=np.zeros_like(x),
             dt=np.float32(0.1), G=np.float32(1))
    with engine(s) as e:
        e.bounding_box(); e.hilbert_sort()
        keys, perm = e.bvh_keys()
        xs = e.download(("x",))["x"]
    assert np.array_equal(np.bincount(perm, minlength=n), np.ones(n, np.int64))
    ks = keys[perm]
    assert (ks[1:] >= ks[:-1]).all()
    eq = ks[1:] == ks[:-1]
    assert (perm[1:][eq] > perm[:-1][eq]).all()
    assert same(xs, x[perm])


def test_traversal_stats_match_oracle_visits(oracle_fast):
    """The warp-cooperative walk performs, per body, exactly the reference's sequence of node tests: the (body, node)
    test count equals the oracle's loop-iteration count."""
    s = oracle_fast.galaxy(20000, np.float64, 3)
    lo, hi = oracle_fast.bbox(s["x"])
    so = oracle_fast.permute(oracle_fast.sort_perm(oracle_fast.keys(s["x"], lo, hi)), s)
    nm, bw, _ = oracle_fast.bvh_build(so["m"], so["x"])
    _, visits = oracle_fast.bvh_force(so["m"], so["x"], nm, bw, s["G"], THETA)
    with engine(s) as e:
        e.bounding_box(); e.hilbert_sort(); e.build_tree()
        st = e.traversal_stats()
    assert st["node_visits"] == visits
    assert st["warp_steps"] * st["width"] >= st["node_visits"] and st["interactions"] <= st["node_visits"]


def test_bvh_full_pipeline_4M(oracle):
    """A BASELINE-scale run (n = 4M, 3-D float): keys and permutation bit-exact against the pinned oracle, root node =
    total mass / centre of mass, and the traversal of 1000 sampled (sorted) bodies against the oracle's walk."""
    n = 4_000_000
    s = oracle.galaxy(n, np.float32, 3)
    with engine(s) as e:
        lo, hi = e.bounding_box()
        e.hilbert_sort()
        keys, perm = e.bvh_keys()
        e.build_tree()
        e.bvh_compute_force()
        out = e.download()
        st = e.traversal_stats()
    olo, ohi = oracle.bbox(s["x"])
    assert same(lo, olo) and same(hi, ohi)
    okeys = oracle.keys(s["x"], olo, ohi)
    assert same(keys, okeys), "Hilbert keys must be bit-exact at scale"
    assert same(perm, oracle.sort_perm(okeys))
    assert same(out["x"], s["x"][perm]) and same(out["m"], s["m"][perm])
    nm, bw, _ = oracle.bvh_build(out["m"], out["x"])
    rng = np.random.default_rng(3)
    targets = np.sort(rng.choice(n, 1000, replace=False)).astype(np.uint32)
    ref, visits = oracle.bvh_force(out["m"], out["x"], nm, bw, s["G"], THETA, targets=targets)
    err = rel_err(out["a"][targets], ref)
    assert rms(err) <= 2e-5 and err.max() <= 5e-4, (rms(err), err.max())
    assert abs(st["node_visits"] / n - visits / len(targets)) < 0.1 * visits / len(targets)
    assert np.isfinite(out["a"]).all()
