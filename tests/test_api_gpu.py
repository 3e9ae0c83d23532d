"""C-ABI behaviour on a GPU: call-order errors, partial upload/download, re-upload, extreme theta, tiny n."""
import numpy as np
import pytest

import _pkg
from golden_util import rel_err, same

pytestmark = pytest.mark.gpu
nbx = _pkg.load().nbx


def test_call_order_errors(oracle):
    s = oracle.galaxy(100, np.float32, 3)
    with nbx.Engine(100, 3, np.float32, "bvh", s["dt"], s["G"]) as e:
        e.upload_state(s)
        with pytest.raises(nbx.NbxError) as ei:
            e.hilbert_sort()           # before bounding_box
        assert ei.value.code == -6
        with pytest.raises(nbx.NbxError):
            e.bvh_compute_force()      # before build_tree
        with pytest.raises(nbx.NbxError):
            e.octree_build()           # wrong engine kind
    with nbx.Engine(100, 3, np.float32, "octree", s["dt"], s["G"]) as e:
        e.upload_state(s)
        with pytest.raises(nbx.NbxError):
            e.octree_compute_force()   # before build
        with pytest.raises(nbx.NbxError):
            e.traversal_stats()
    with nbx.Engine(100, 3, np.float32, "all-pairs", s["dt"], s["G"]) as e:
        with pytest.raises(nbx.NbxError):
            e.traversal_stats()


def test_partial_upload_download_and_reupload(oracle):
    s = oracle.galaxy(500, np.float64, 2)
    with nbx.Engine(500, 2, np.float64, "all-pairs", s["dt"], s["G"]) as e:
        e.upload(m=s["m"], x=s["x"])                      # v, a, ao stay zero
        out = e.download()
        assert same(out["x"], s["x"]) and same(out["m"], s["m"]) and not out["v"].any()
        e.upload(v=s["v"])                                # later, only the velocities
        assert same(e.download(("v",))["v"], s["v"])
        e.step(2)
        first = e.download()
        e.upload_state(s)                                 # same engine, fresh state: identical trajectory
        e.step(2)
        second = e.download()
        for k in ("x", "v", "a", "ao"):
            assert same(first[k], second[k]), k
        c = e.counters()
        assert c["kernel_launches"] > 0 and c["h2d_bytes"] > 0 and c["d2h_bytes"] > 0


@pytest.mark.parametrize("algo", ["octree", "bvh"])
def test_extreme_theta(oracle, algo):
    """theta = 0 opens everything (== all-pairs); a huge theta accepts the root at once (one monopole per body)."""
    s = oracle.galaxy(400, np.float64, 3)
    with nbx.Engine(400, 3, np.float64, algo, s["dt"], s["G"], theta=1e6) as e:
        e.upload_state(s)
        e.step(1)
        st = e.traversal_stats()
        a_big = e.download(("a",))["a"]
    ref = oracle.run(algo, s, 1, 1e6)
    assert rel_err(a_big, ref["a"]).max() < 1e-11
    # octree: the root is accepted immediately; bvh: every body stops at the root or very near it
    assert st["node_visits"] <= 400 * (1 if algo == "octree" else 3)


@pytest.mark.parametrize("algo", ["all-pairs", "all-pairs-collapsed", "octree", "bvh"])
@pytest.mark.parametrize("n", [1, 2, 3])
def test_tiny_systems(oracle, algo, n):
    if algo == "bvh" and n < 2:
        with pytest.raises(nbx.NbxError):
            nbx.Engine(1, 3, np.float32, "bvh", 1.0, 1.0)
        return
    s = oracle.galaxy(4, np.float64, 3)
    s = {k: (v[:n].copy() if isinstance(v, np.ndarray) else v) for k, v in s.items()}
    with nbx.Engine(n, 3, np.float64, algo, s["dt"], s["G"]) as e:
        e.upload_state(s)
        e.step(2)
        out = e.download()
    ref = oracle.run(algo, s, 2)
    assert np.isfinite(out["x"]).all()
    assert np.allclose(out["x"], ref["x"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("algo,dt", [("all-pairs", np.float32), ("octree", np.float64), ("bvh", np.float32)])
def test_streamed_position_frames_equal_synchronous_downloads(oracle, algo, dt):
    """nbx_stream_positions_begin/end (Saver streaming): frames copied while the next steps run are bit-identical to the
    frames a synchronous nbx_download returns after the same steps."""
    s = oracle.galaxy(20000, dt, 3)
    with nbx.Engine(20000, 3, dt, algo, s["dt"], s["G"]) as e:
        e.upload_state(s)
        want = []
        for _ in range(5):
            e.step(1)
            want.append(e.download(("x",))["x"])
    with nbx.Engine(20000, 3, dt, algo, s["dt"], s["G"]) as e:
        e.upload_state(s)
        got = []
        for k in range(5):
            e.step(1)
            e.stream_positions_begin()          # frame k in flight while step k+1 runs
            if k >= 1:
                got.append(e.stream_positions_end())
        got.append(e.stream_positions_end())
        with pytest.raises(nbx.NbxError):
            e.stream_positions_end()            # nothing left in flight
    assert len(got) == 5
    for a, b in zip(got, want):
        assert same(a, b)


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
def test_bvh_graph_replay_equals_plain_launches(oracle, monkeypatch, dt):
    """The single-GPU BVH step is replayed from a CUDA graph (one per buffer parity). Same bits as launching kernel by
    kernel (NBX_GRAPH=0), also when per-phase calls in between change the buffer assignment (the graph is re-captured),
    when phase timing is switched on and off, and when the state is re-uploaded between steps."""
    n, dim = 5000, 3
    s = oracle.galaxy(n, dt, dim)

    def run():
        with nbx.Engine(n, dim, dt, "bvh", s["dt"], s["G"], theta=0.5) as e:
            e.upload_state(s)
            e.step(3)                       # capture parity 0, capture parity 1, replay parity 0
            e.bounding_box(); e.hilbert_sort()  # per-phase call: flips the buffers outside a step
            e.step(2)                       # parity now paired with other v/a/ao buffers -> re-capture
            e.set_phase_timing(True)
            e.step_timed(1)                 # plain launches with events in between
            e.set_phase_timing(False)
            mid = e.download()
            e.upload(x=mid["x"])            # re-upload between replays
            e.step(4)
            return e.download(), e.counters()["kernel_launches"]

    monkeypatch.setenv("NBX_GRAPH", "1")
    a, la = run()
    monkeypatch.setenv("NBX_GRAPH", "0")
    b, lb = run()
    for k in ("m", "x", "v", "a", "ao"):
        assert a[k].tobytes() == b[k].tobytes(), k
    assert la == lb  # replays account for the kernels they launch
