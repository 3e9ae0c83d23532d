"""GPU parity of the all-pairs / all-pairs-collapsed / leapfrog kernels (through the C ABI) against the CPU oracle
and the committed golden vectors.

Tolerances (SURVEY §8(c) contract; floor = what two builds of the reference differ by):
  * one force evaluation `a`: float  per-body relative error vs the DOUBLE oracle: rms <= 5e-5, and no worse than
    2x the reference's own float-vs-double error;  double: max <= 1e-12.
  * leapfrog given the same `a`: bit-exact.
  * k-step trajectories: float rms-rel(x) <= 1e-5 at n=96 (3 steps), double <= 1e-11.
"""
import numpy as np
import pytest

import _pkg
from golden_util import CASES, DT, IDS, STEPS, init_state, load, rel_err, rms, same

pytestmark = pytest.mark.gpu
nbx = _pkg.load().nbx

TOL_A = {np.dtype(np.float32): (5e-5, 5e-4), np.dtype(np.float64): (1e-13, 1e-12)}  # (rms, max)


def run_force(s, algorithm="all-pairs", flags=0):
    n, dim = s["x"].shape
    with nbx.Engine(n, dim, s["x"].dtype, algorithm, s["dt"], s["G"], flags=flags) as e:
        e.upload_state(s)
        if algorithm == "all-pairs":
            e.all_pairs_force()
        else:
            e.all_pairs_collapsed_force()
        return e.download()


def as64(s):
    return {k: (np.asarray(v, np.float64) if isinstance(v, np.ndarray) else np.float64(v)) for k, v in s.items()}


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_force_vs_golden(tag, dim):
    g = load(tag, dim)
    s = init_state(g)
    out = run_force(s)
    err = rel_err(out["a"], g["a_all_pairs"])
    tol_rms, tol_max = TOL_A[s["x"].dtype]
    assert rms(err) <= tol_rms and err.max() <= tol_max, (rms(err), err.max())
    assert same(out["x"], s["x"]) and same(out["m"], s["m"]) and same(out["v"], s["v"])  # round trip untouched


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
@pytest.mark.parametrize("n", [2, 10, 1001, 4096])
def test_force_vs_oracle(oracle, oracle_fast, tag, dim, n):
    dt = DT[tag]
    s = oracle.galaxy(n, dt, dim)
    out = run_force(s)
    truth = oracle_fast.all_pairs_force_truth(s["m"], s["x"], s["G"])  # double arithmetic, eps of this precision
    err = rel_err(out["a"], truth)
    tol_rms, tol_max = TOL_A[np.dtype(dt)]
    assert rms(err) <= tol_rms and err.max() <= tol_max, (rms(err), err.max())
    if dt == np.float32:
        ref_err = rel_err(oracle.all_pairs_force(s["m"], s["x"], s["G"]), truth)  # the reference's own float error
        assert rms(err) <= 2 * rms(ref_err) + 1e-7
    if n % 2:
        assert not out["a"][-1].any() or s["m"][-1] == 0  # zero-mass body at the origin still gets a finite force
    assert np.isfinite(out["a"]).all()


def test_coincident_and_zero_mass_bodies(oracle):
    """Edge cases: duplicate positions (d2 = 0 between distinct bodies) and zero masses give finite results equal to
    the reference formula m*0/eps = 0."""
    s = oracle.galaxy(64, np.float32, 3)
    s["x"][10] = s["x"][11]
    s["m"][5] = 0
    out = run_force(s)
    ref = oracle.all_pairs_force(s["m"], s["x"], s["G"])
    assert np.isfinite(out["a"]).all()
    assert rel_err(out["a"], ref).max() < 1e-4


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_accelerate_step_bit_exact(oracle, tag, dim):
    g = load(tag, dim)
    s = init_state(g)
    s["a"] = g["a_all_pairs"]
    s["ao"] = (g["a_all_pairs"] * s["x"].dtype.type(0.25)).astype(s["x"].dtype)
    x, v, ao = oracle.accelerate(s["x"], s["v"], s["a"], s["ao"], s["dt"])
    with nbx.Engine(len(s["m"]), dim, s["x"].dtype, "all-pairs", s["dt"], s["G"]) as e:
        e.upload_state(s)
        e.accelerate_step()
        out = e.download()
    assert same(out["x"], x) and same(out["v"], v) and same(out["ao"], ao) and same(out["a"], s["a"])


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_steps_vs_golden_and_fused_equals_unfused(tag, dim):
    g = load(tag, dim)
    s = init_state(g)
    outs = []
    for flags in (0, nbx.FLAG_NO_FUSED_INTEGRATE):
        with nbx.Engine(len(s["m"]), dim, s["x"].dtype, "all-pairs", s["dt"], s["G"], flags=flags) as e:
            e.upload_state(s)
            e.step(STEPS)
            outs.append(e.download())
    for k in ("x", "v", "a", "ao"):
        assert same(outs[0][k], outs[1][k]), k  # the fused epilogue is the same arithmetic
    tol = 1e-5 if tag == "f32" else 1e-11
    assert rms(rel_err(outs[0]["x"], g["run_all_pairs_x"])) <= tol
    assert rms(rel_err(outs[0]["v"], g["run_all_pairs_v"])) <= 10 * tol
    assert same(outs[0]["a"], outs[0]["ao"])


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_collapsed_vs_oracle(oracle, tag, dim):
    """all-pairs-collapsed only accumulates components 0 and 1 and resets through `a -= ao`."""
    dt = DT[tag]
    s = oracle.galaxy(777, dt, dim)
    rng = np.random.default_rng(3)
    s["a"] = (rng.standard_normal(s["x"].shape) * 1e-5).astype(dt)
    s["ao"] = s["a"].copy()
    s["ao"][:, 0] *= dt(0.5)  # make the "reset" leave a residue that must be carried
    ref = oracle.collapsed_force(s["m"], s["x"], s["a"], s["ao"], s["G"])
    out = run_force(s, "all-pairs-collapsed")
    tol_rms, tol_max = TOL_A[np.dtype(dt)]
    err = rel_err(out["a"][:, :2], np.asarray(ref, np.float64)[:, :2])
    assert rms(err) <= max(tol_rms, 1e-12) and err.max() <= max(tol_max, 2e-11), (rms(err), err.max())
    if dim == 3:
        assert same(out["a"][:, 2], s["a"][:, 2])  # z untouched, bug-compatible
        fixed = run_force(s, "all-pairs-collapsed", flags=nbx.FLAG_COLLAPSED_FIX_Z)
        full = oracle.all_pairs_force(s["m"], s["x"], s["G"])
        want_z = (s["a"][:, 2] - s["ao"][:, 2]) + full[:, 2]
        assert np.allclose(fixed["a"][:, 2], want_z, rtol=1e-4 if dt == np.float32 else 1e-11, atol=1e-12)


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_collapsed_steps_vs_golden(tag, dim):
    g = load(tag, dim)
    s = init_state(g)
    with nbx.Engine(len(s["m"]), dim, s["x"].dtype, "all-pairs-collapsed", s["dt"], s["G"]) as e:
        e.upload_state(s)
        e.step(STEPS)
        out = e.download()
    tol = 1e-5 if tag == "f32" else 1e-11
    assert rms(rel_err(out["x"], g["run_collapsed_x"])) <= tol
    if dim == 3:
        assert not out["a"][:, 2].any()


def test_energies(oracle):
    s = oracle.galaxy(500, np.float64, 3)
    k, gpot = oracle.energies(s["m"], s["x"], s["v"], s["G"])
    with nbx.Engine(500, 3, np.float64, "all-pairs", s["dt"], s["G"]) as e:
        e.upload_state(s)
        k2, g2 = e.calc_energies()
    assert abs(k2 - k) <= 1e-12 * abs(k) and abs(g2 - gpot) <= 1e-12 * abs(gpot)


@pytest.mark.parametrize("dt,dim,n,tol", [(np.float64, 2, 1999, 1e-12), (np.float64, 3, 5000, 1e-12),
                                          (np.float32, 3, 3001, 2e-6), (np.float32, 2, 777, 2e-6)])
def test_energies_sizes(oracle, dt, dim, n, tol):
    """calc_energies evaluates the strictly upper triangle and doubles it; compared with the formula in double
    (the reference accumulates in T: its own float result is only good to ~1e-4 at these sizes)."""
    s = oracle.galaxy(n, dt, dim)
    m, x, v = s["m"].astype(np.float64), s["x"].astype(np.float64), s["v"].astype(np.float64)
    eps = float(np.finfo(dt).eps)
    k = 0.5 * (m * (v * v).sum(1)).sum()
    d = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)) + eps
    w = (m[:, None] * m[None, :]) / d
    np.fill_diagonal(w, 0.0)
    g = -0.5 * float(s["G"]) * w.sum()
    with nbx.Engine(len(m), dim, dt, "all-pairs", s["dt"], s["G"]) as e:
        e.upload_state(s)
        k2, g2 = e.calc_energies()
    assert abs(k2 - k) <= max(tol, 1e-7 if dt == np.float32 else 0) * abs(k)
    assert abs(g2 - g) <= tol * abs(g), (g2, g)


def test_full_size_properties_1M(oracle_fast):
    """BASELINE config 2 (all-pairs 3-D float, n = 1M): sampled targets against the oracle, plus the
    size-independent property sum_i m_i a_i = 0 (Newton's third law; every pair term is exactly antisymmetric)."""
    n = 1_000_000
    s = oracle_fast.galaxy(n, np.float32, 3)
    out = run_force(s)
    assert np.isfinite(out["a"]).all()
    rng = np.random.default_rng(0)
    targets = np.sort(rng.choice(n, 48, replace=False)).astype(np.uint32)
    truth = oracle_fast.all_pairs_force_truth(s["m"], s["x"], s["G"], targets=targets)
    err = rel_err(out["a"][targets], truth)
    assert rms(err) <= 5e-5 and err.max() <= 5e-4, (rms(err), err.max())
    ref32 = oracle_fast.all_pairs_force(s["m"], s["x"], s["G"], targets=targets)
    assert rms(err) <= 2 * rms(rel_err(ref32, truth)) + 1e-7
    f = out["a"].astype(np.float64) * s["m"].astype(np.float64)[:, None]
    assert np.abs(f.sum(0)).max() <= 1e-4 * np.abs(f).sum(0).max()


# ---- symmetric (Newton's third law) all-pairs: default for n >= 16384, forced here at small n ---------------------------------
@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
@pytest.mark.parametrize("n", [2, 1001, 4096, 5000])
def test_symmetric_force_vs_oracle(oracle, oracle_fast, tag, dim, n):
    dt = DT[tag]
    s = oracle.galaxy(n, dt, dim)
    sym = run_force(s, flags=nbx.FLAG_ALLPAIRS_SYMMETRIC)
    ordered = run_force(s, flags=nbx.FLAG_ALLPAIRS_ORDERED)
    truth = oracle_fast.all_pairs_force_truth(s["m"], s["x"], s["G"])
    err = rel_err(sym["a"], truth)
    tol_rms, tol_max = TOL_A[np.dtype(dt)]
    assert rms(err) <= tol_rms and err.max() <= tol_max, (rms(err), err.max())
    # same pair terms, different summation order only
    assert rel_err(sym["a"], ordered["a"]).max() <= (1e-3 if dt == np.float32 else 1e-12)
    assert np.isfinite(sym["a"]).all()


@pytest.mark.parametrize("dim", [2, 3])
def test_packed_fp32x2_kernels_vs_scalar(oracle_fast, monkeypatch, dim):
    """The float symmetric kernel runs its pair arithmetic on FP32x2 instructions (two j bodies per instruction), with 8
    targets per thread where the block size allows (here: n in (2^18, 2^19] => B = 2048). Same pair terms as the scalar
    kernel, the action sums split into an even-j and an odd-j partial: identical up to summation order."""
    n = 300_000
    s = oracle_fast.galaxy(n, np.float32, dim)
    out = {}
    for mode in ("0", "1", "2"):
        monkeypatch.setenv("NBX_SYM_PACKED", mode)
        out[mode] = run_force(s)["a"]
    rng = np.random.default_rng(5)
    targets = np.sort(rng.choice(n, 64, replace=False)).astype(np.uint32)
    truth = oracle_fast.all_pairs_force_truth(s["m"], s["x"], s["G"], targets=targets)
    for mode in ("0", "1", "2"):
        err = rel_err(out[mode][targets], truth)
        assert rms(err) <= 5e-5 and err.max() <= 5e-4, (mode, rms(err), err.max())
        f = out[mode].astype(np.float64) * s["m"].astype(np.float64)[:, None]
        assert np.abs(f.sum(0)).max() <= 1e-4 * np.abs(f).sum(0).max()   # Newton's third law
    assert rms(rel_err(out["1"], out["0"])) <= 2e-6 and rms(rel_err(out["2"], out["0"])) <= 2e-6
    assert not np.array_equal(out["1"], out["0"])   # really a different kernel: the summation order differs


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_symmetric_steps_fused_equals_unfused_and_reproducible(oracle, tag, dim):
    dt = DT[tag]
    s = oracle.galaxy(3000, dt, dim)
    outs = []
    for flags in (nbx.FLAG_ALLPAIRS_SYMMETRIC, nbx.FLAG_ALLPAIRS_SYMMETRIC | nbx.FLAG_NO_FUSED_INTEGRATE,
                  nbx.FLAG_ALLPAIRS_SYMMETRIC):
        with nbx.Engine(3000, dim, dt, "all-pairs", s["dt"], s["G"], flags=flags) as e:
            e.upload_state(s)
            e.step(STEPS)
            outs.append(e.download())
    for k in ("x", "v", "a", "ao"):
        assert same(outs[0][k], outs[1][k]), k   # fused epilogue == separate kernels
        assert same(outs[0][k], outs[2][k]), k   # no atomics: bit-reproducible run to run
    ref = oracle.run("all-pairs", s, STEPS)
    tol = 1e-5 if tag == "f32" else 1e-11
    assert rms(rel_err(outs[0]["x"], ref["x"])) <= tol


@pytest.mark.parametrize("flags", [0, 0x4], ids=["symmetric-default", "ordered"])
def test_both_kernels_at_131072(oracle_fast, flags):
    n = 131072
    s = oracle_fast.galaxy(n, np.float32, 3)
    out = run_force(s, flags=flags)
    rng = np.random.default_rng(2)
    targets = np.sort(rng.choice(n, 64, replace=False)).astype(np.uint32)
    truth = oracle_fast.all_pairs_force_truth(s["m"], s["x"], s["G"], targets=targets)
    err = rel_err(out["a"][targets], truth)
    assert rms(err) <= 5e-5 and err.max() <= 5e-4, (rms(err), err.max())
    f = out["a"].astype(np.float64) * s["m"].astype(np.float64)[:, None]
    assert np.abs(f.sum(0)).max() <= 1e-4 * np.abs(f).sum(0).max()


@pytest.mark.parametrize("tag,dim", CASES, ids=IDS)
def test_collapsed_symmetric_path(oracle_fast, tag, dim):
    """n >= 16384: all-pairs-collapsed runs on the symmetric block-pair units; same collapsed semantics (components 0,1
    only, `a -= ao` reset carried), checked against the oracle's collapsed loop on sampled rows."""
    dt = DT[tag]
    n = 20000
    s = oracle_fast.galaxy(n, dt, dim)
    rng = np.random.default_rng(9)
    s["a"] = (rng.standard_normal(s["x"].shape) * 1e-5).astype(dt)
    s["ao"] = s["a"].copy()
    s["ao"][:, 0] *= dt(0.5)
    out = run_force(s, "all-pairs-collapsed")
    ordered = run_force(s, "all-pairs-collapsed", flags=nbx.FLAG_ALLPAIRS_ORDERED)   # the atomics kernel
    full = oracle_fast.all_pairs_force_truth(s["m"], s["x"], s["G"])
    want = (s["a"].astype(np.float64) - s["ao"].astype(np.float64))[:, :2] + full[:, :2]
    tol_rms, tol_max = TOL_A[np.dtype(dt)]
    for got in (out, ordered):
        err = rel_err(got["a"][:, :2], want)
        assert rms(err) <= max(tol_rms, 1e-12) and err.max() <= max(tol_max, 2e-11), (rms(err), err.max())
    if dim == 3:
        assert same(out["a"][:, 2], s["a"][:, 2])


def test_symmetric_buffer_too_large_falls_back_to_the_ordered_kernel(oracle_fast, monkeypatch):
    """When the partial-sum buffer of the symmetric kernel does not fit, a single-GPU engine serves the step with the
    ordered sweep instead (advisor finding: no bare cudaMalloc failure): bit-identical to an engine created with
    NBX_FLAG_ALLPAIRS_ORDERED."""
    monkeypatch.setenv("NBX_SYM_MAX_MB", "1")
    n, dim = 30000, 3
    for algo in ("all-pairs", "all-pairs-collapsed"):
        s = oracle_fast.galaxy(n, np.float32, dim)
        outs = []
        for flags in (0, nbx.FLAG_ALLPAIRS_ORDERED):
            with nbx.Engine(n, dim, np.float32, algo, s["dt"], s["G"], flags=flags) as e:
                e.upload_state(s)
                e.step(3)
                outs.append(e.download())
        for k in ("x", "v", "a", "ao"):
            assert outs[0][k].tobytes() == outs[1][k].tobytes(), (algo, k)
