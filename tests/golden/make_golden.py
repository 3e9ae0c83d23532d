"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/refdump_d{2,3}, built by oracle/Makefile
from /root/reference/src with -O2 -ffp-contract=off). Run in the build container:  python tests/golden/make_golden.py
The 2-D float BVH artefacts come from the AVX-512 native build (float->u32 overflow saturates)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
N = 96
STEPS = 3
THETA = 0.5


def bvh_nodes(buf, off, dt, dim):
    nn = int(np.frombuffer(buf, np.uint64, 1, off)[0]); off += 8
    nm = np.frombuffer(buf, dt, nn * (dim + 1), off).reshape(nn, dim + 1); off += nm.nbytes
    bw = np.frombuffer(buf, dt, nn, off); off += bw.nbytes
    b = np.frombuffer(buf, dt, nn * 2 * dim, off).reshape(nn, 2, dim)
    return nm, bw, b


def octree_arrays(buf, off, dt, dim):
    used = int(np.frombuffer(buf, np.uint64, 1, off)[0]); off += 8
    side = np.frombuffer(buf, dt, 1, off)[0]; off += np.dtype(dt).itemsize
    root = np.frombuffer(buf, dt, dim, off); off += root.nbytes
    fc = np.frombuffer(buf, np.uint32, used, off); off += fc.nbytes
    par = np.frombuffer(buf, np.uint32, 1 + used // (1 << dim), off); off += par.nbytes
    nm = np.frombuffer(buf, dt, used * (dim + 1), off).reshape(used, dim + 1)
    return dict(used=used, side=side, root=root, first_child=fc, parent=par, node_m=nm)


def main():
    orc = O.Oracle(fast=False)  # only used for the numbering-independent canonicalisation of the reference's octree
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        for dim in (2, 3):
            native = dt == np.float32 and dim == 2
            if native and not O.ref_native_available():
                raise SystemExit("need an AVX-512 host for the 2-D float fixtures")
            s = O.ref_galaxy(N, dt, dim)
            out = {f"init_{k}": s[k] for k in ("m", "x", "v")}
            out["dt"], out["G"] = s["dt"], s["G"]
            r, _, _ = O.ref_state_op("force_all_pairs", s)
            out["a_all_pairs"] = r["a"]
            bb = np.frombuffer(O.refdump("bbox", dt, dim, N, s, native=native), dt).reshape(2, dim)
            out["bbox"] = bb
            out["keys"] = np.frombuffer(O.refdump("keys", dt, dim, N, s, native=native), np.uint64)
            rs, buf, off = O.ref_state_op("bvh_build", s, native=native)
            out["sorted_m"], out["sorted_x"], out["sorted_v"] = rs["m"], rs["x"], rs["v"]
            out["bvh_m"], out["bvh_bw"], out["bvh_b"] = bvh_nodes(buf, off, dt, dim)
            for theta in (0.0, THETA):
                rf, _, _ = O.ref_state_op("bvh_force", s, theta=theta, native=native)
                out[f"a_bvh_theta{theta}"] = rf["a"]
                rf, _, _ = O.ref_state_op("octree_force", s, theta=theta)
                out[f"a_octree_theta{theta}"] = rf["a"]
            _, buf, off = O.ref_state_op("octree_build", s)
            t = octree_arrays(buf, off, dt, dim)
            depth, path, kind, mo = orc.octree_canonical(t, dim)
            out["octree_used"], out["octree_side"], out["octree_root"] = t["used"], t["side"], t["root"]
            out["octree_depth"], out["octree_path"], out["octree_kind"], out["octree_m"] = depth, path, kind, mo
            for algo, op in (("all_pairs", "run_all_pairs"), ("collapsed", "run_collapsed"), ("octree", "run_octree"),
                             ("bvh", "run_bvh")):
                rr, _, _ = O.ref_state_op(op, s, theta=THETA, steps=STEPS, native=native and algo == "bvh")
                for k in ("x", "v", "a"):
                    out[f"run_{algo}_{k}"] = rr[k]
            np.savez_compressed(os.path.join(HERE, f"galaxy_{tag}_d{dim}_n{N}.npz"), **out)
    rng = np.random.default_rng(11)
    c3 = np.concatenate([np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [0x1FFFFF, 0, 0],
                                   [0x1FFFFF] * 3, [0x123456, 0x0ABCDE, 0x1F0F0F]], np.uint32),
                         rng.integers(0, 1 << 21, (249, 3), dtype=np.uint32)])
    c2 = np.concatenate([np.array([[1, 0], [1, 1], [0, 1], [0xFFFFFFFF, 0], [0x12345678, 0x9ABCDEF0]], np.uint32),
                         rng.integers(0, 1 << 32, (251, 2), dtype=np.uint64).astype(np.uint32)])
    k3 = np.frombuffer(O.refdump("hilbert_cells", np.float32, 3, len(c3), raw_in=c3.tobytes()), np.uint64)
    k2 = np.frombuffer(O.refdump("hilbert_cells", np.float32, 2, len(c2), raw_in=c2.tobytes()), np.uint64)
    np.savez_compressed(os.path.join(HERE, "hilbert_known_answers.npz"), c2=c2, k2=k2, c3=c3, k3=k3)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
