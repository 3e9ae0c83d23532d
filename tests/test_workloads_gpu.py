"""Other input distributions: a uniform cube and a centrally concentrated (Plummer-like) sphere
stress tree depth and the close-pair softening differently from the galaxy discs. Every algorithm against the oracle."""
import numpy as np
import pytest

import _pkg
from golden_util import rel_err, rms, same

pytestmark = pytest.mark.gpu
nbx = _pkg.load().nbx


def uniform_cube(n, dim, dt, seed=1):
    rng = np.random.default_rng(seed)
    x = (rng.random((n, dim)) * 2 - 1).astype(dt)
    v = (rng.random((n, dim)) * 2 - 1).astype(dt)
    return dict(m=np.full(n, 1.0 / n, dt), x=x, v=v, a=np.zeros_like(x), ao=np.zeros_like(x), dt=dt(0.1), G=dt(1.0))


def plummer_like(n, dt, seed=2):
    rng = np.random.default_rng(seed)
    r = 1.0 / np.sqrt(rng.random(n) ** (-2.0 / 3.0) - 1.0)
    u = rng.standard_normal((n, 3))
    x = (u / np.linalg.norm(u, axis=1, keepdims=True) * r[:, None]).astype(dt)
    v = (rng.standard_normal((n, 3)) * 0.1).astype(dt)
    return dict(m=np.full(n, 1.0 / n, dt), x=x, v=v, a=np.zeros_like(x), ao=np.zeros_like(x), dt=dt(1.0), G=dt(6.674e-11))


STATES = {
    "uniform-3d-f32": lambda: uniform_cube(6000, 3, np.float32),
    "uniform-2d-f64": lambda: uniform_cube(6000, 2, np.float64),
    "plummer-3d-f64": lambda: plummer_like(6000, np.float64),
    "plummer-3d-f32": lambda: plummer_like(6000, np.float32),
}


def tolerances(dt):
    return (5e-5, 2e-3) if np.dtype(dt) == np.float32 else (1e-12, 1e-10)


@pytest.mark.parametrize("name", list(STATES))
@pytest.mark.parametrize("algo", ["all-pairs", "all-pairs-sym", "octree", "bvh"])
def test_force_vs_oracle(oracle_fast, oracle, name, algo):
    s = STATES[name]()
    n, dim = s["x"].shape
    dt = s["x"].dtype
    flags = nbx.FLAG_ALLPAIRS_SYMMETRIC if algo == "all-pairs-sym" else 0
    real_algo = "all-pairs" if algo.startswith("all-pairs") else algo
    with nbx.Engine(n, dim, dt, real_algo, s["dt"], s["G"], theta=0.5, flags=flags) as e:
        e.upload_state(s)
        if real_algo == "all-pairs":
            e.all_pairs_force()
            a = e.download(("a",))["a"]
            ref = oracle_fast.all_pairs_force_truth(s["m"], s["x"], s["G"])
        elif algo == "octree":
            e.octree_build(); e.octree_compute_force()
            a = e.download(("a",))["a"]
            t = oracle.octree_build(s["m"], s["x"])
            gd, gp, gk, gm = e.octree_canonical()
            depth, path, kind, mo = oracle.octree_canonical(t, dim)
            assert same(gd, depth) and same(gk, kind) and same(gm, mo)
            ref, _ = oracle_fast.octree_force(s["x"], t, s["G"], 0.5)
        else:
            e.bounding_box(); e.hilbert_sort(); e.build_tree(); e.bvh_compute_force()
            out = e.download()
            a = out["a"]
            keys, perm = e.bvh_keys()
            lo, hi = oracle.bbox(s["x"])
            assert same(keys, oracle.keys(s["x"], lo, hi))
            nm, bw, _ = oracle.bvh_build(out["m"], out["x"])
            ref, _ = oracle_fast.bvh_force(out["m"], out["x"], nm, bw, s["G"], 0.5)
    err = rel_err(a, ref)
    tr, tm = tolerances(dt)
    assert rms(err) <= tr and err.max() <= tm, (rms(err), err.max())


@pytest.mark.parametrize("name", ["uniform-3d-f32", "plummer-3d-f64"])
@pytest.mark.parametrize("algo", ["all-pairs", "octree", "bvh"])
def test_five_steps_vs_oracle(oracle_fast, name, algo):
    s = STATES[name]()
    n, dim = s["x"].shape
    with nbx.Engine(n, dim, s["x"].dtype, algo, s["dt"], s["G"], theta=0.5) as e:
        e.upload_state(s)
        e.step(5)
        out = e.download()
    ref = oracle_fast.run(algo, s, 5, 0.5)
    tol = 2e-4 if s["x"].dtype == np.float32 else 1e-9
    scale = np.abs(ref["x"]).max()
    assert np.abs(out["x"].astype(np.float64) - ref["x"]).max() <= tol * scale


# ---- degenerate geometries: the integer artefacts must stay bit-exact ---------------------------------------------------------
def _state(x, dt):
    x = np.ascontiguousarray(x, dt)
    n = len(x)
    rng = np.random.default_rng(11)
    return dict(m=(rng.random(n) + 0.5).astype(dt), x=x, v=np.zeros_like(x), a=np.zeros_like(x), ao=np.zeros_like(x),
                dt=dt(0.01), G=dt(1.0))


def _geometries(dt, dim):
    rng = np.random.default_rng(5)
    n = 3000
    line = np.zeros((n, dim)); line[:, 0] = np.linspace(-7, 13, n)                  # zero extent in the other axes
    plane = rng.random((n, dim)) * 50; plane[:, -1] = 3.25                           # one flat axis
    far = rng.random((n, dim)) * 1e-3 + 1e4                                          # tiny cloud far from the origin
    big = (rng.random((n, dim)) - 0.5) * 2e6                                         # huge coordinates
    clusters = np.concatenate([rng.standard_normal((n // 2, dim)) * 1e-2 - 40, rng.standard_normal((n - n // 2, dim)) * 5 + 60])
    return {"line": line, "plane": plane, "far-cloud": far, "huge": big, "two-clusters": clusters}


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("dim", [2, 3])
def test_degenerate_geometries_bvh_and_octree(oracle, oracle_fast, dt, dim):
    for name, x in _geometries(dt, dim).items():
        s = _state(x, dt)
        if len(np.unique(s["x"], axis=0)) != len(s["x"]):
            continue  # rounding to dt made coincident bodies: the octree would (correctly) report them
        n = len(s["m"])
        # BVH: keys, permutation, nodes
        lo, hi = oracle.bbox(s["x"])
        keys = oracle.keys(s["x"], lo, hi)
        perm = oracle.sort_perm(keys)
        so = oracle.permute(perm, s)
        nm, bw, b = oracle.bvh_build(so["m"], so["x"])
        with nbx.Engine(n, dim, dt, "bvh", s["dt"], s["G"], theta=0.5) as e:
            e.upload_state(s)
            glo, ghi = e.bounding_box()
            e.hilbert_sort()
            gkeys, gperm = e.bvh_keys()
            e.build_tree()
            gm, gbw, gb = e.bvh_nodes()
            e.bvh_compute_force()
            a = e.download(("a",))["a"]
        assert same(glo, lo) and same(ghi, hi), name
        assert same(gkeys, keys) and same(gperm, perm), name
        assert same(gm, nm) and same(gbw, bw) and same(gb, b), name
        ref, _ = oracle_fast.bvh_force(so["m"], so["x"], nm, bw, s["G"], 0.5)
        assert rms(rel_err(a, ref)) <= (5e-5 if dt == np.float32 else 1e-12), name
        # octree: canonical topology + monopoles
        try:
            t = oracle.octree_build(s["m"], s["x"])
        except RuntimeError:
            continue  # deeper than the reference's own node capacity allows
        depth, path, kind, mo = oracle.octree_canonical(t, dim)
        with nbx.Engine(n, dim, dt, "octree", s["dt"], s["G"], theta=0.5) as e:
            e.upload_state(s)
            try:
                e.octree_build()
            except nbx.NbxError as ex:
                assert ex.code == -4, name   # only "more levels than two key words hold" may be refused
                continue
            gd, gp, gk, gmo = e.octree_canonical()
            e.octree_compute_force()
            a = e.download(("a",))["a"]
        assert same(gd, depth) and same(gk, kind) and same(gmo, mo), name
        ref, _ = oracle_fast.octree_force(s["x"], t, s["G"], 0.5)
        assert rms(rel_err(a, ref)) <= (5e-5 if dt == np.float32 else 1e-12), name


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("algo", ["bvh", "octree"])
def test_small_n_walk_variants_keep_every_bit(oracle, monkeypatch, algo, dt):
    """Small problems run the tree walks with partially filled warps (8 / 16 bodies per warp) and, for the BVH, a sector
    touch of the right sibling. Neither changes a body's sequence of tests or its arithmetic: the whole
    state after three steps is the same bit for bit for every combination (NBX_BVH_LANES / NBX_OCT_LANES / NBX_BVH_TOUCH)."""
    n, dim = 6000, 3
    s = oracle.galaxy(n, dt, dim)
    got = []
    for lanes in ("32", "16", "8"):
        for touch in ("0", "1"):
            monkeypatch.setenv("NBX_BVH_LANES", lanes)
            monkeypatch.setenv("NBX_OCT_LANES", lanes)
            monkeypatch.setenv("NBX_BVH_TOUCH", touch)
            with nbx.Engine(n, dim, dt, algo, s["dt"], s["G"], theta=0.5) as e:
                e.upload_state(s)
                e.step(3)
                got.append(e.download())
    for other in got[1:]:
        for k in ("m", "x", "v", "a", "ao"):
            assert other[k].tobytes() == got[0][k].tobytes(), k
