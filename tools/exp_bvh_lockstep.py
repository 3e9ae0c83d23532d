"""Experiment helper (CPU only, uses the oracle): emulates the warp-cooperative BVH walk of nbx_bvh.cu for a sample of
warps and classifies every warp step — all 32 lanes on the same node ("lock-step") or not, unanimous accept / reject /
mixed, and whether a conservative test against the warp's bounding box would already decide the node for all lanes.
It answers whether a Bonsai-style walk (group-level tests + interaction lists evaluated densely) could replace the
per-lane tests while keeping the reference's exact per-body interaction sets: at n = 1 M only 40 % of the steps are in
lock-step and 32 % are decided by the box test, in short runs (mean 7 steps) — an upper bound of ~1.25x for much more
machinery, so it was not pursued.
    python tools/exp_bvh_lockstep.py [n]
legend of the printed sequence: a/r = lock-step, decided by the box test; A/R = lock-step unanimous but not box-clear;
M = lock-step with mixed outcomes; B = body level; x = lanes on different nodes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
o = O.Oracle(fast=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
theta = 0.5
s = o.galaxy(n, np.float32, 3)
lo, hi = o.bbox(s["x"])
so = o.permute(o.sort_perm(o.keys(s["x"], lo, hi)), s)
nm, bw, _ = o.bvh_build(so["m"], so["x"])
x = so["x"].astype(np.float32)
levels = o.bvh_levels(n)
th2 = np.float32(theta) * np.float32(theta)
rng = np.random.default_rng(0)
nw = 24
warps = rng.choice(n // 32 - 1, nw, replace=False)
tot = dict(steps=0, lock=0, lock_acc=0, lock_rej=0, lock_mixed=0, body=0, pattern=0, visits=0, clear_lock=0)
t0 = time.time()
for w in warps:
    ids = np.arange(w * 32, w * 32 + 32)
    xs = x[ids]
    bmin, bmax = xs.min(0), xs.max(0)
    cov = np.zeros(32, np.int64); lev = np.zeros(32, np.int64)
    hist = []  # (lock, outcome) per step: outcome 'A','R','M','B'
    while True:
        active_mask = cov < n
        if not active_mask.any(): break
        key = np.where(active_mask, cov * 64 + lev, 1 << 62)
        kmin = key.min()
        act = key == kmin
        c, L = int(kmin // 64), int(kmin % 64)
        tot["steps"] += 1; tot["visits"] += int(act.sum())
        lock = bool(act.all())
        if L == levels:
            cov[act] += 2; lev[act] -= 1
            hist.append((lock, 'B')); tot["body"] += 1
            continue
        sh = levels - L
        k = (1 << L) - 1 + (c >> sh)
        com = nm[k, :3].astype(np.float32); w2 = np.float32(bw[k]) * np.float32(bw[k])
        d = com[None, :] - xs
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        ok = w2 < th2 * d2
        take = act & ok
        rej = act & ~ok
        # box clear test
        dd = np.maximum(np.maximum(bmin - com, com - bmax), 0); dmin2 = float((dd * dd).sum())
        far = np.maximum(np.abs(com - bmin), np.abs(com - bmax)); dmax2 = float((far * far).sum())
        clear = (w2 < th2 * dmin2 * (1 - 1e-5)) or (w2 >= th2 * dmax2 * (1 + 1e-5))
        if lock:
            tot["lock"] += 1
            if ok.all(): tot["lock_acc"] += 1; out = 'A'
            elif (~ok).all(): tot["lock_rej"] += 1; out = 'R'
            else: tot["lock_mixed"] += 1; out = 'M'
            if clear and out != 'M': tot["clear_lock"] += 1; out = out.lower()  # lower-case: also clear by the box test
        else:
            out = 'x'
        hist.append((lock, out))
        right = (k & 1) == 0 and k != 0
        cov[take] += 1 << sh
        if k == 0: pass
        lev[take] -= 1 if right else 0
        lev[rej] += 1
    # count steps inside r a a blocks (clear, lock-step)
    seq = ''.join(o_ for _, o_ in hist)
    i = 0
    while i + 2 < len(seq):
        if seq[i:i+3] == 'raa': tot["pattern"] += 3; i += 3
        else: i += 1
print(f"n={n} levels={levels} warps={nw} time {time.time()-t0:.0f}s")
st = tot["steps"]
for k, v in tot.items(): print(f"  {k:12s} {v:9d}  {100.0*v/st:6.1f}% of steps")
print("  lane util", tot["visits"] / (32 * st))
print(seq[:600])
import re
runs = re.findall(r'[ar]+', seq)
print("clear lock runs:", len(runs), "mean len", np.mean([len(r) for r in runs]), "steps in runs>=8:", sum(len(r) for r in runs if len(r)>=8), "of", len(seq))
from collections import Counter
c = Counter()
for r in runs:
    for i in range(len(r)-2): c[r[i:i+3]] += 1
print(c.most_common(8))
