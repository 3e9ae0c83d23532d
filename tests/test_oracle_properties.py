"""Size-independent properties of the oracle itself (cheap CPU checks that complement the golden vectors)."""
import numpy as np

from golden_util import rel_err


def test_hilbert_keys_are_injective_and_order_preserving_per_cell(oracle):
    rng = np.random.default_rng(0)
    c3 = np.unique(rng.integers(0, 1 << 21, (4000, 3), dtype=np.uint32), axis=0)
    c2 = np.unique(rng.integers(0, 1 << 32, (4000, 2), dtype=np.uint64).astype(np.uint32), axis=0)
    k3, k2 = oracle.hilbert(c3), oracle.hilbert(c2)
    assert len(np.unique(k3)) == len(c3) and len(np.unique(k2)) == len(c2)
    assert int(k3.max()) < (1 << 63)          # 3 x 21 bits


def test_hilbert_2d_visits_neighbouring_cells(oracle):
    """A true Hilbert curve: consecutive keys over a full 2^4 x 2^4 top-level grid are edge neighbours."""
    side = 16
    cells = np.array([[x << 28, y << 28] for x in range(side) for y in range(side)], np.uint32)
    keys = oracle.hilbert(cells)
    order = np.argsort(keys)
    xy = (cells[order] >> 28).astype(np.int64)
    step = np.abs(np.diff(xy, axis=0)).sum(axis=1)
    assert (step == 1).all()


def test_sort_perm_is_a_stable_permutation(oracle):
    rng = np.random.default_rng(1)
    keys = rng.integers(0, 50, 5000, dtype=np.uint64)  # many ties
    perm = oracle.sort_perm(keys)
    assert np.array_equal(np.sort(perm), np.arange(5000, dtype=np.uint32))
    assert np.array_equal(perm, np.argsort(keys, kind="stable").astype(np.uint32))


def test_newtons_third_law_and_leapfrog_time_reversal(oracle):
    s = oracle.galaxy(400, np.float64, 3)
    a = oracle.all_pairs_force(s["m"], s["x"], s["G"])
    f = a * s["m"][:, None]
    assert np.abs(f.sum(0)).max() <= 1e-12 * np.abs(f).sum(0).max()
    # theta = 0 trees reproduce all-pairs (README.md:122-129)
    t = oracle.octree_build(s["m"], s["x"])
    ao, _ = oracle.octree_force(s["x"], t, s["G"], 0.0)
    assert rel_err(ao, a).max() < 1e-10
