"""Pins the plain-C oracle (oracle/nbody_oracle.c) bit-for-bit against the UNMODIFIED reference compiled into
oracle/_ref/refdump_d{2,3} (same -O2 -ffp-contract=off flags). Skipped where the _ref binaries are absent."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.skipif(not (O.ref_available(2) and O.ref_available(3)), reason="oracle/_ref not built")

CASES = [(np.float32, 2), (np.float32, 3), (np.float64, 2), (np.float64, 3)]
IDS = ["f32-2d", "f32-3d", "f64-2d", "f64-3d"]


def same(a, b):
    return np.ascontiguousarray(a).tobytes() == np.ascontiguousarray(b).tobytes()


def use_native(dt, dim):
    # 2-D float keys hit the float->u32 overflow: only the AVX-512 native build saturates like
    # the oracle (and CUDA) do.
    return dt == np.float32 and dim == 2


@pytest.mark.parametrize("dt,dim", CASES, ids=IDS)
@pytest.mark.parametrize("n", [10, 11, 1000])
def test_galaxy_bit_exact(oracle, dt, dim, n):
    r, s = O.ref_galaxy(n, dt, dim), oracle.galaxy(n, dt, dim)
    assert r["x"].shape == s["x"].shape
    for k in "mxv":
        assert same(r[k], s[k]), k
    if n % 2:  # last body all-zero
        assert s["m"][-1] == 0 and not s["x"][-1].any()


@pytest.mark.parametrize("dt,dim", CASES, ids=IDS)
def test_force_and_integrate_bit_exact(oracle, dt, dim):
    s = O.ref_galaxy(300, dt, dim)
    r, _, _ = O.ref_state_op("force_all_pairs", s)
    assert same(r["a"], oracle.all_pairs_force(s["m"], s["x"], s["G"]))
    s2 = dict(s, a=r["a"], ao=(r["a"] * dt(0.5)).astype(dt))
    rc, _, _ = O.ref_state_op("force_collapsed", s2)
    assert same(rc["a"], oracle.collapsed_force(s2["m"], s2["x"], s2["a"], s2["ao"], s2["G"]))
    ra, _, _ = O.ref_state_op("accelerate", s2)
    x, v, ao = oracle.accelerate(s2["x"], s2["v"], s2["a"], s2["ao"], s2["dt"])
    assert same(ra["x"], x) and same(ra["v"], v) and same(ra["ao"], ao)


@pytest.mark.parametrize("dt,dim", CASES, ids=IDS)
def test_bvh_artefacts_bit_exact(oracle, dt, dim):
    native = use_native(dt, dim)
    if native and not O.ref_native_available():
        pytest.skip("needs the AVX-512 native refdump (2-D float key overflow is UB in the reference)")
    n = 500
    s = O.ref_galaxy(n, dt, dim)
    bb = np.frombuffer(O.refdump("bbox", dt, dim, n, s, native=native), dt).reshape(2, dim)
    lo, hi = oracle.bbox(s["x"])
    assert same(bb[0], lo) and same(bb[1], hi)
    kr = np.frombuffer(O.refdump("keys", dt, dim, n, s, native=native), np.uint64)
    ko = oracle.keys(s["x"], lo, hi)
    assert same(kr, ko)
    assert len(np.unique(ko)) == n  # no ties => unstable std::sort is deterministic here
    rs, _, _ = O.ref_state_op("sort", s, native=native)
    so = oracle.permute(oracle.sort_perm(ko), s)
    for k in ("m", "x", "v", "a", "ao"):
        assert same(rs[k], so[k]), k
    _, buf, off = O.ref_state_op("bvh_build", s, native=native)
    nn = int(np.frombuffer(buf, np.uint64, 1, off)[0]); off += 8
    nm = np.frombuffer(buf, dt, nn * (dim + 1), off).reshape(nn, dim + 1); off += nm.nbytes
    bw = np.frombuffer(buf, dt, nn, off); off += bw.nbytes
    b = np.frombuffer(buf, dt, nn * 2 * dim, off).reshape(nn, 2, dim)
    onm, obw, ob = oracle.bvh_build(so["m"], so["x"])
    assert same(nm, onm) and same(bw, obw) and same(b, ob)
    for theta in (0.0, 0.5):
        rf, _, _ = O.ref_state_op("bvh_force", s, theta=theta, native=native)
        a, _ = oracle.bvh_force(so["m"], so["x"], onm, obw, s["G"], theta)
        assert same(rf["a"], a), theta


def test_f32_2d_key_overflow_is_the_only_difference(oracle):
    """Generic x86-64 builds of the reference wrap the out-of-range cell of the extreme bodies to 0; the oracle
    saturates. Every other key must agree."""
    dt, dim, n = np.float32, 2, 500
    s = O.ref_galaxy(n, dt, dim)
    lo, hi = oracle.bbox(s["x"])
    kr = np.frombuffer(O.refdump("keys", dt, dim, n, s), np.uint64)
    ko = oracle.keys(s["x"], lo, hi)
    cell = (hi - lo) / dt(0xFFFFFFFF)
    q = (s["x"] - lo) / cell
    overflow = (q >= dt(4294967296.0)).any(axis=1)
    assert overflow.sum() >= 1
    assert same(kr[~overflow], ko[~overflow])


@pytest.mark.parametrize("dt,dim", CASES, ids=IDS)
def test_octree_bit_exact(oracle, dt, dim):
    n = 500
    s = O.ref_galaxy(n, dt, dim)
    _, buf, off = O.ref_state_op("octree_build", s)
    used = int(np.frombuffer(buf, np.uint64, 1, off)[0]); off += 8
    side = np.frombuffer(buf, dt, 1, off)[0]; off += np.dtype(dt).itemsize
    root = np.frombuffer(buf, dt, dim, off); off += root.nbytes
    fc = np.frombuffer(buf, np.uint32, used, off); off += fc.nbytes
    par = np.frombuffer(buf, np.uint32, 1 + used // (1 << dim), off); off += par.nbytes
    nm = np.frombuffer(buf, dt, used * (dim + 1), off).reshape(used, dim + 1)
    t = oracle.octree_build(s["m"], s["x"])
    # the reference executed sequentially (serial PSTL backend) numbers nodes exactly like the oracle's schedule
    assert used == t["used"] and side == t["side"] and same(root, t["root"])
    assert same(fc, t["first_child"]) and same(par, t["parent"])
    mask = fc != 0xFFFFFFFF
    assert same(nm[mask], t["node_m"][mask])
    for theta in (0.0, 0.5):
        rf, _, _ = O.ref_state_op("octree_force", s, theta=theta)
        a, _ = oracle.octree_force(s["x"], t, s["G"], theta)
        assert same(rf["a"], a), theta


@pytest.mark.parametrize("dt,dim", CASES, ids=IDS)
@pytest.mark.parametrize("algo,op", [("all-pairs", "run_all_pairs"), ("all-pairs-collapsed", "run_collapsed"),
                                     ("octree", "run_octree"), ("bvh", "run_bvh")])
def test_multi_step_bit_exact(oracle, dt, dim, algo, op):
    native = algo == "bvh" and use_native(dt, dim)
    if native and not O.ref_native_available():
        pytest.skip("needs the AVX-512 native refdump")
    s = O.ref_galaxy(200, dt, dim)
    rr, _, _ = O.ref_state_op(op, s, theta=0.5, steps=3, native=native)
    oo = oracle.run(algo, s, 3, 0.5)
    for k in ("m", "x", "v", "a", "ao"):
        assert same(rr[k], oo[k]), k


def test_hilbert_known_answers_from_reference(oracle):
    """SURVEY §8(c)(1) known-answer table, re-derived here from the reference's own hilbert<N>()."""
    rng = np.random.default_rng(7)
    c3 = np.concatenate([np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [0x1FFFFF, 0, 0],
                                   [0x1FFFFF, 0x1FFFFF, 0x1FFFFF], [0x123456, 0x0ABCDE, 0x1F0F0F]], np.uint32),
                         rng.integers(0, 1 << 21, (200, 3), dtype=np.uint32)])
    k3 = np.frombuffer(O.refdump("hilbert_cells", np.float32, 3, len(c3), raw_in=c3.tobytes()), np.uint64)
    assert [hex(int(k)) for k in k3[:7]] == ["0x6", "0x2", "0x1", "0x5", "0x6db6db6db6db6db6", "0x5b6db6db6db6db6d",
                                             "0x774dc24749504e4d"]
    assert same(k3, oracle.hilbert(c3))
    c2 = np.concatenate([np.array([[1, 0], [1, 1], [0, 1], [0xFFFFFFFF, 0], [0x12345678, 0x9ABCDEF0]], np.uint32),
                         rng.integers(0, 1 << 32, (200, 2), dtype=np.uint64).astype(np.uint32)])
    k2 = np.frombuffer(O.refdump("hilbert_cells", np.float32, 2, len(c2), raw_in=c2.tobytes()), np.uint64)
    assert [hex(int(k)) for k in k2[:5]] == ["0x1", "0x2", "0x3", "0xffffffffffffffff", "0x42484aca4842406a"]
    assert same(k2, oracle.hilbert(c2))
