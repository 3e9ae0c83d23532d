"""GPU parity of the octree pipeline (bounds -> path keys -> sort -> cells -> monopoles -> traversal) through the C ABI.

The tree is compared in the numbering-independent canonical form: DFS list of non-empty nodes
(depth, path code, leaf/internal, monopole) — BIT-EXACT against the oracle (which executes the reference's insert
sequentially) including the monopoles. Accelerations: float rms <= 2e-5 / max <= 5e-4 of the pinned float oracle
(the float walk uses sqrt/rcp approximations and `side < theta*dx` for `side/dx < theta`), double max <= 1e-11.
"""
import numpy as np
import pytest

import _pkg
from golden_util import CASES, DT, IDS, STEPS, THETA, init_state, load, rel_err, rms, same

pytestmark = pytest.mark.gpu
nbx = _pkg.load().nbx

TOL = {np.dtype(np.float32): (2e-5, 5e-4), np.dtype(np.float64): (1e-12, 1e-11)}


def engine(s, theta=THETA, algo="octree"):
    n, dim = s["x"].shape
    e = nbx.Engine(n, dim, s["x"].dtype, algo, s["dt"], s["G"], theta=theta)
    e.upload_state(s)
    return e


def check_tree(oracle, s, thetas=(0.0, THETA)):
    dim = s["x"].shape[1]
    t = oracle.octree_build(s["m"], s["x"])
    depth, path, kind, mo = oracle.octree_canonical(t, dim)
    with engine(s) as e:
        e.octree_build()
        side, root, used = e.octree_root()
        assert side == t["side"] and same(root, t["root"]) and used == t["used"]
        gd, gp, gk, gm = e.octree_canonical()
        assert same(gd, depth) and same(gp, path) and same(gk, kind), "octree topology must be identical"
        assert same(gm, mo), "monopoles must be bit-exact"
    for theta in thetas:
        ref, _ = oracle.octree_force(s["x"], t, s["G"], theta)
        with engine(s, theta) as e:
            e.octree_build(); e.octree_compute_force()
            a = e.download(("a",))["a"]
        err = rel_err(a, ref)
        tr, tm = TOL[s["x"].dtype]
        assert rms(err) <= tr and err.max() <= tm, (theta, rms(err), err.max())


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_tree_vs_golden(tag, dim):
    g = load(tag, dim)
    s = init_state(g)
    with engine(s) as e:
        e.octree_build()
        side, root, used = e.octree_root()
        assert used == int(g["octree_used"]) and side == g["octree_side"] and same(root, g["octree_root"])
        gd, gp, gk, gm = e.octree_canonical()
        assert same(gd, g["octree_depth"]) and same(gp, g["octree_path"]) and same(gk, g["octree_kind"])
        assert same(gm, g["octree_m"])
    tr, tm = TOL[s["x"].dtype]
    for theta in (0.0, THETA):
        with engine(s, theta) as e:
            e.octree_build(); e.octree_compute_force()
            a = e.download(("a",))["a"]
        err = rel_err(a, g[f"a_octree_theta{theta}"])
        assert rms(err) <= tr and err.max() <= tm, (theta, rms(err), err.max())


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
@pytest.mark.parametrize("n", [1, 2, 3, 1000, 4097])
def test_tree_vs_oracle(oracle, tag, dim, n):
    s = oracle.galaxy(max(n, 2), DT[tag], dim)
    if n == 1:
        s = {k: (v[:1].copy() if isinstance(v, np.ndarray) else v) for k, v in s.items()}
    check_tree(oracle, s)


def test_tree_vs_oracle_65536(oracle):
    check_tree(oracle, oracle.galaxy(65536, np.float32, 3), thetas=(THETA,))


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_steps_vs_golden(tag, dim):
    g = load(tag, dim)
    s = init_state(g)
    with engine(s) as e:
        e.step(STEPS)
        out = e.download()
    tol = 2e-5 if tag == "f32" else 1e-11
    assert rms(rel_err(out["x"], g["run_octree_x"])) <= tol
    assert rms(rel_err(out["v"], g["run_octree_v"])) <= 20 * tol
    assert same(out["m"], s["m"])  # the octree never permutes the state


def test_theta0_equals_all_pairs(oracle):
    """README.md:122-129 (different softening form: (sqrt(d2)+eps)^3 vs d2^1.5+eps)."""
    s = oracle.galaxy(3000, np.float64, 3)
    with engine(s, theta=0.0) as e:
        e.octree_build(); e.octree_compute_force()
        a = e.download(("a",))["a"]
    with engine(s, algo="all-pairs") as e:
        e.all_pairs_force()
        ap = e.download(("a",))["a"]
    assert rel_err(a, ap).max() < 1e-10


@pytest.mark.parametrize("tag,dim,gap", [("f64", 3, 1e-7), ("f64", 2, 1e-9), ("f32", 3, 2e-5)])
def test_deep_tree_two_word_keys(oracle, tag, dim, gap):
    """Two bodies closer than the 21-level (3-D) / 32-level (2-D) cell size: the build falls back to two-word keys and
    still reproduces the reference's (deeper) tree: same depths/kinds/monopoles in DFS order, same forces."""
    dt = DT[tag]
    s = oracle.galaxy(600, dt, dim)
    s["x"][41] = s["x"][40]
    s["x"][41, 0] += dt(gap)
    t = oracle.octree_build(s["m"], s["x"])
    depth, path, kind, mo = oracle.octree_canonical(t, dim)
    assert depth.max() > (21 if dim == 3 else 32)
    with engine(s) as e:
        e.octree_build()
        side, root, used = e.octree_root()
        assert used == t["used"]
        gd, gp, gk, gm = e.octree_canonical()
        assert same(gd, depth) and same(gk, kind) and same(gm, mo)
        e.octree_compute_force()
        a = e.download(("a",))["a"]
    ref, _ = oracle.octree_force(s["x"], t, s["G"], THETA)
    tr, tm = TOL[np.dtype(dt)]
    err = rel_err(a, ref)
    assert rms(err) <= tr and err.max() <= 10 * tm, (rms(err), err.max())


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_hilbert_lane_order_is_bit_identical(oracle, monkeypatch, tag, dim):
    """For n >= 4 M the walk assigns its lanes along a Hilbert curve instead of the path (Z) order. Only the grouping of
    bodies into warps changes: every body performs the same operations on the same records, so the result is bit-identical
    (forced on and off here at a small n)."""
    s = oracle.galaxy(6001, DT[tag], dim)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("NBX_OCT_HILBERT", flag)
        with engine(s) as e:
            e.octree_build()
            e.octree_compute_force()
            out[flag] = e.download(("a",))["a"]
            st = e.traversal_stats()
        out["visits" + flag] = st["node_visits"]
    assert same(out["0"], out["1"])
    assert out["visits0"] == out["visits1"]


def test_coincident_bodies_are_reported(oracle):
    """The reference splits forever on coincident bodies (no capacity check, octree.h:146-169); nbx reports it."""
    s = oracle.galaxy(64, np.float32, 3)
    s["x"][10] = s["x"][11]
    with engine(s) as e:
        with pytest.raises(nbx.NbxError) as ei:
            e.octree_build()
        assert ei.value.code == -4


def test_tree_properties_10M(oracle_fast):
    """BASELINE config 4 size (n = 10M, 3-D double): structural invariants + sampled targets against the oracle."""
    n = 10_000_000
    s = oracle_fast.galaxy(n, np.float64, 3)
    with engine(s) as e:
        e.octree_build()
        side, root, used = e.octree_root()
        depth, path, kind, mono = e.octree_canonical()
        e.octree_compute_force()
        a = e.download(("a",))["a"]
    assert kind.sum() == n                                  # every body is exactly one leaf
    assert (used - 1) % 8 == 0 and (used - 1) // 8 == (kind == 0).sum()
    assert depth[0] == 0 and kind[0] == 0
    assert abs(mono[0, 3] - s["m"].sum()) <= 1e-9 * s["m"].sum()   # root mass = total mass
    com = (s["m"][:, None] * s["x"]).sum(0) / s["m"].sum()
    assert np.abs(mono[0, :3] - com).max() < 1e-6
    assert np.isfinite(a).all()
    rng = np.random.default_rng(1)
    targets = np.sort(rng.choice(n, 2000, replace=False)).astype(np.uint32)
    t = oracle_fast.octree_build(s["m"], s["x"])
    assert t["used"] == used
    ref, visits = oracle_fast.octree_force(s["x"], t, s["G"], THETA, targets=targets)
    err = rel_err(a[targets], ref)
    assert err.max() <= 1e-10, err.max()
