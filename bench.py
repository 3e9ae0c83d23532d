#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json): all-pairs 3-D float galaxy, n = 1M bodies,
G pair-interactions/s, on 1/2/4/8 B200 (state replicated; block-pair units dealt round-robin over the ranks, one NCCL
all-reduce of the accelerations per step). At N = 1 the line also carries `configs`: the other BASELINE.json
configurations that fit one GPU (C3 collapsed double n = 262144, C4 octree double n = 10 M, bvh float n = 10 M), each
with its own roofline and CPU baseline; at N > 1 it carries `verify`: the sharded run checked against a single-GPU run.

  python bench.py --gpus N --steps K --warmup W            our arm (libnbx.so through the C ABI)
  python bench.py --impl reference ...                      the UNMODIFIED reference's CPU build (oracle/_ref) timed
                                                            on the host cores, bounded sample of the same workload
One JSON line on stdout (rank 0). A "step" = one force evaluation over all n(n-1) ordered pairs + leapfrog update.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = {3: 20.0, 2: 14.0}  # counted from all_pairs.h:23 + vec.h:232-252
METRIC = "G pair-interactions/s (all-pairs 3-D float galaxy)"
UNIT = "Gpairs/s"
REF_SAMPLE_N = 20000  # CPU-feasible sample of the same workload (the metric is a rate, not a wall time)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nbx", choices=["nbx", "reference"])
    ap.add_argument("-n", type=int, default=1_000_000)
    ap.add_argument("--algorithm", default="all-pairs", choices=["all-pairs", "all-pairs-collapsed", "octree", "bvh"])
    ap.add_argument("--precision", default="float", choices=["float", "double"])
    ap.add_argument("--dim", type=int, default=3, choices=[2, 3])
    ap.add_argument("--theta", type=float, default=0.5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the extra BASELINE configs appended at N = 1")
    ap.add_argument("--no-verify", action="store_true", help="skip the N > 1 result check against a single-GPU run")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        """Median SM clock and active throttle reasons of the samples taken in [t0, t1] (the timed region); the sampler is
        started before the warm-up so that it is already streaming when the region begins."""
        rows = [r for ts, r in self.rows if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.15)]
        if not rows:
            rows = [r for _, r in self.rows]
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if c > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def units_per_step(args, n):
    if args.algorithm.startswith("all-pairs"):
        return n * (n - 1) / 1e9  # G ordered pair interactions
    return n / 1e6  # M body-steps


def metric_unit(args):
    if args.algorithm.startswith("all-pairs"):
        return (f"G pair-interactions/s ({args.algorithm} {args.dim}-D {args.precision} galaxy)", "Gpairs/s")
    return (f"Mbody-steps/s ({args.algorithm} {args.dim}-D {args.precision} galaxy theta={args.theta})", "Mbody-steps/s")


def workload_name(args, n):
    return f"{args.algorithm} {args.dim}-D {args.precision} galaxy n={n}" + (
        "" if args.algorithm.startswith("all-pairs") else f" theta={args.theta}")


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation, the UNMODIFIED /root/reference/src/main.cpp built by oracle/Makefile into
# oracle/_ref/: nbody_d{2,3}_omp = its CPU flags + an OpenMP PSTL backend (oracle/pstl_backend_omp.h) so that its
# std::execution::par algorithms use ALL host cores (the image has no TBB, libstdc++'s only parallel backend);
# nbody_d{2,3}[_native] = the same source with the serial PSTL backend (1 core). Falls back to the oracle port.
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def ref_binary(dim):
    """-> (path, threads it runs on) or (None, 0)"""
    for name, threads in ((f"nbody_d{dim}_omp", host_threads()), (f"nbody_d{dim}_native", 1), (f"nbody_d{dim}", 1)):
        exe = os.path.join(ROOT, "oracle", "_ref", name)
        if os.access(exe, os.X_OK):
            try:  # the -march=native build may not run on this host
                r = subprocess.run([exe, "-n", "16", "-s", "1", "--workload", "galaxy", "--algorithm", "all-pairs",
                                    "--csv-detailed"], capture_output=True, text=True, timeout=60)
                if r.returncode == 0:
                    return exe, threads
            except (OSError, subprocess.TimeoutExpired):
                pass
    return None, 0


def ref_descr(exe, threads):
    if threads > 1:
        return (f"{os.path.basename(exe)}: unmodified reference, g++ -Ofast -fopenmp, OpenMP PSTL backend "
                f"(oracle/pstl_backend_omp.h; no TBB in the image), {threads} threads")
    return f"{os.path.basename(exe)}: unmodified reference, g++ -Ofast, serial PSTL backend (no TBB in the image), 1 thread"


def sample_n(args, threads=1):
    """Bounded sample of the workload: a few seconds of CPU work per step (the metric is a rate, not a wall time)."""
    par = threads > 1
    if args.algorithm == "all-pairs":
        return min(args.n, 100_000 if par else REF_SAMPLE_N)
    if args.algorithm == "all-pairs-collapsed":
        return min(args.n, 20_000 if par else 6000)
    return min(args.n, 1_000_000 if par else 200_000)


def run_reference_once(args, exe, n, steps, threads=1):
    """Runs of the reference binary with --csv-detailed (exactly the requested steps, no hidden warm-up: SURVEY §9 Q1).
    Returns seconds for `steps` steps as the reference itself measured them and wall seconds. Its timer prints 10 ms
    resolution, so a run shorter than 0.2 s is repeated with enough steps in ONE process to last about half a second and
    scaled back."""
    env = dict(os.environ, OMP_NUM_THREADS=str(max(threads, 1)), OMP_PROC_BIND="false")

    def once(k):
        cmd = [exe, "-n", str(n), "-s", str(k), "--workload", "galaxy", "--algorithm", args.algorithm,
               "--precision", args.precision, "--theta", str(args.theta), "--csv-detailed"]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True, check=True, env=env)
        wall = time.perf_counter() - t0
        row = [ln for ln in r.stdout.splitlines() if ln.startswith(args.algorithm + ",")][-1].split(",")
        return float(row[5]), wall

    total, wall = once(steps)
    if total < 0.2:
        k = steps * max(2, int(np.ceil(0.5 / max(total, wall / 20, 1e-3))))
        total_k, wall = once(k)
        total = (total_k if total_k >= 0.2 else wall) * steps / k
    return total, wall


def cpu_port_rate(args, n, steps):
    from oracle import oracle as O
    orc = O.Oracle(fast=True)
    dt = np.float32 if args.precision == "float" else np.float64
    s = orc.galaxy(n, dt, args.dim)
    t0 = time.perf_counter()
    orc.run(args.algorithm, s, steps, args.theta)
    return units_per_step(args, n) * steps / (time.perf_counter() - t0), os.cpu_count()


def cpu_baseline(args, steps=1):
    exe, threads = ref_binary(args.dim)
    n = sample_n(args, threads)
    _, unit = metric_unit(args)
    if exe:
        secs, _ = run_reference_once(args, exe, n, steps, threads)
        return {"value": units_per_step(args, n) * steps / secs, "unit": unit, "cores": threads, "kind": "reference",
                "sample": f"{workload_name(args, n)}, {steps} step(s), --csv-detailed; {ref_descr(exe, threads)}"}
    rate, cores = cpu_port_rate(args, n, steps)
    return {"value": rate, "unit": unit, "cores": cores, "kind": "port",
            "sample": f"oracle port (-Ofast, OpenMP) {workload_name(args, n)}, {steps} step(s)"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    metric, unit = metric_unit(args)
    exe, threads = ref_binary(args.dim)
    n = sample_n(args, threads)
    per_step = []
    for it in range(args.warmup + args.steps):
        if exe:
            secs, _ = run_reference_once(args, exe, n, 1, threads)
        else:
            rate, _ = cpu_port_rate(args, n, 1)
            secs = units_per_step(args, n) / rate
        if it >= args.warmup:
            per_step.append(secs)
    total = sum(per_step)
    value = units_per_step(args, n) * args.steps / total
    kind = "reference" if exe else "port"
    cores = threads if exe else os.cpu_count()
    sample = (f"{workload_name(args, n)}; each step = one process run of 1 step (--csv-detailed); "
              f"{ref_descr(exe, threads) if exe else 'oracle port (-Ofast, OpenMP)'}")
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.precision == "float" else "f64",
            "data": "synthetic (galaxy model, mt19937{42})",
            "config": {"workload": workload_name(args, args.n), "sample": workload_name(args, n)},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# our arm
def make_state(n, dtype, dim):
    """Synthetic input: the reference's galaxy model from the PRODUCT's own generator (host driver `--dry-run --save pos`,
    bit-identical to src/models.h:112-136; tests/test_host_driver.py). Returns state_t arrays + dt, G."""
    import tempfile
    exe = os.path.join(ROOT, "stdpar-nbody_b200", "bin", f"nbody_d{dim}")
    if not os.access(exe, os.X_OK):
        raise SystemExit(f"bench.py: {exe} is missing — build with __graft_entry__.build()")
    dtype = np.dtype(dtype)
    need = (2 * dim + 1) * dtype.itemsize * n * 10  # room for every rank of an 8-GPU run generating at once
    shm = None
    try:
        st = os.statvfs("/dev/shm")
        if os.access("/dev/shm", os.W_OK) and st.f_bavail * st.f_frsize > need:
            shm = "/dev/shm"
    except OSError:
        pass
    with tempfile.TemporaryDirectory(dir=shm) as td:
        subprocess.run([exe, "-n", str(n), "-s", "1", "--workload", "galaxy", "--precision",
                        "float" if dtype == np.float32 else "double", "--dry-run", "--save", "pos"], cwd=td, check=True,
                       stdout=subprocess.DEVNULL)
        with open(os.path.join(td, "positions.bin"), "rb") as f:
            hdr = np.fromfile(f, np.uint32, 4)
            assert int(hdr[0]) == n and int(hdr[2]) == dtype.itemsize and int(hdr[3]) == dim, hdr
            x = np.fromfile(f, dtype, n * dim).reshape(n, dim)
            v = np.fromfile(f, dtype, n * dim).reshape(n, dim)
            m = np.fromfile(f, dtype, n)
    assert len(m) == n
    return dict(m=m, x=x, v=v, a=np.zeros_like(x), ao=np.zeros_like(x), dt=dtype.type(10.0), G=dtype.type(1e-4))


class Ctx:
    """Process-wide plumbing shared by every measurement of this run."""

    def __init__(self):
        import torch

        import _pkg
        self.torch = torch
        self.nbx = _pkg.load().nbx
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.nbx.device_count() < 1:
            raise SystemExit("bench.py: no CUDA device — libnbx has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        self.sm_count = torch.cuda.get_device_properties(self.local_rank).multi_processor_count
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            self.peaks = {}
        self.profile_consts = {}
        for name in ("r02_walk_profile.json",):
            try:
                self.profile_consts.update(json.load(open(os.path.join(ROOT, "profiles", name))))
            except (OSError, ValueError):
                pass

    def flush_l2(self):
        self.flush.zero_()
        self.torch.cuda.synchronize()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if not self.dist:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def new_engine(self, s, cfg, multi=True):
        nbx = self.nbx
        n, dim = s["x"].shape
        world, rank = (self.world, self.rank) if multi else (1, 0)
        eng = nbx.Engine(n, dim, s["x"].dtype, cfg.algorithm, s["dt"], s["G"], theta=cfg.theta, device=self.local_rank,
                         rank=rank, world_size=world)
        if world > 1:
            ids = [nbx.comm_unique_id() if rank == 0 else None]
            self.dist.broadcast_object_list(ids, src=0)
            eng.comm_init_rank(ids[0])
            if cfg.algorithm == "bvh" and os.environ.get("NBX_PEER", "0") == "1":  # opt-in: see DESIGN.md §7
                self.peer_setup(eng)
        eng.upload_state(s)
        return eng

    def peer_setup(self, eng):
        """Peer-memory exchange of the BVH accelerations (nbx_peer_export / nbx_peer_import): CUDA IPC handles gathered
        through torch.distributed. All ranks agree on the outcome; where IPC is not available the NCCL all-gather stays."""
        nbx, torch = self.nbx, self.torch
        ok = 1
        try:
            handles = [None] * self.world
            self.dist.all_gather_object(handles, eng.peer_export())
            eng.peer_import(handles)
        except nbx.NbxError as ex:
            ok = 0
            if self.rank == 0:
                print(f"bench.py: peer buffers unavailable ({ex}); using the NCCL all-gather", file=sys.stderr, flush=True)
        t = torch.tensor([ok], device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        self.peer = bool(int(t.item()))
        if not self.peer:
            eng.peer_import(None)


def hbm_phase_bytes(cfg, n, isz, world=1):
    """Algorithmic HBM bytes PER RANK of the phases that stream the body arrays once: what the kernels must
    move."""
    rec = 4 * isz
    out = {"accel": 7 * rec * n}  # leapfrog: read x,v,a,ao + write x,v,ao (system.h:52-60)
    if cfg.algorithm == "bvh":
        passes = 8
        # keys (read x, write key) + onesweep radix sort (one 8 B histogram read, then 12 B in + 12 B out per pass) +
        # gather of x,v,a,ao through the permutation
        sort = (8 + passes * 24) * n
        if world > 1 and n >= (1 << 18):  # sharded: owner histogram + one partition sweep over n, radix passes over n / world
            sort = 8 * n + 24 * n + (8 + passes * 24) * (n // world) + 4 * n
        out["sort"] = (rec + 8) * n + sort + (4 + 8 * rec) * n
    elif cfg.algorithm == "octree":
        passes = 8
        # path keys (read x, write key) + onesweep sort (keys kept: the cells are found on the sorted keys) + delta (read the
        # sorted keys, write delta + cell counts); from n = 4 M also the Hilbert lane order: coarse keys + a 4-pass sort
        out["sort"] = (rec + 8) * n + (8 + passes * 24) * n + 16 * n + ((16 + 8 + 4 * 24 - 8) * n if n >= (4 << 20) else 0)
    return out


def roofline_all_pairs(ctx, cfg, n, dim, dt, ph, world):
    nbx = ctx.nbx
    prec = nbx.F32 if dt == np.float32 else nbx.F64
    peak = nbx.measure_fma_peak(prec, ctx.local_rank)
    flops = n * (n - 1) * FLOP_PER_PAIR[dim] / world
    achieved = flops / (ph["force"] * 1e-3) / 1e12 if ph.get("force") else None
    sym = n >= 16384
    return {"bound": "fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if achieved and peak else None, "traffic": None,
            "kernel": ("all_pairs_sym_packed_kernel" if dt == np.float32 else "all_pairs_sym_kernel") if sym else
                      ("all_pairs_kernel" if cfg.algorithm == "all-pairs" else "collapsed_kernel"),
            "kernel_ms": ph.get("force"),
            "note": f"{FLOP_PER_PAIR[dim]:.0f} algorithmic flop per ORDERED pair x n(n-1)/ranks / force time "
                    "(pair kernel + partial-sum reduction); for n >= 16384 the kernel evaluates each unordered "
                    "pair once (Newton's third law, src/all_pairs.h:41-42 TODO) and applies it to both bodies, in float with "
                    "FP32x2 (FFMA2) pair arithmetic; peak = "
                    f"{'FFMA' if prec == nbx.F32 else 'DFMA'} microbenchmark measured in this run "
                    "(MEASURED_PEAKS.json has no FP32/FP64 FMA figure)"}


def roofline_tree(ctx, cfg, n, dim, dt, ph, st, targets, clk):
    """The walks are ISSUE-bound on cache-resident records (ncu: L2 hit 93-99 %, DRAM < 1 % of peak), so the roofline
    of the dominant kernel is warp-instructions/s against 4 schedulers x SMs x clock. Instructions per warp step come
    from the committed ncu capture of the same kernel (profiles/r02_walk_profile.json: smsp__inst_executed.sum / warp
    steps of that launch); warp steps are counted live by the counting twin of the walk."""
    isz = np.dtype(dt).itemsize
    key = f"{cfg.algorithm} {'f32' if isz == 4 else 'f64'} {dim}D"
    prof = ctx.profile_consts.get(key, {})
    mhz = (clk or {}).get("sm_mhz") or (clk or {}).get("sm_max_mhz") or ctx.peaks.get("sm_max_mhz") or 1965.0
    peak = ctx.sm_count * 4 * mhz * 1e6 / 1e9  # G warp-instructions/s
    # the constant belongs to one kernel instantiation: the bvh walk serves 32, 64 or 128 bodies per warp step depending on
    # the problem size, only the captured variant has a committed instruction count
    ips = prof.get("inst_per_warp_step") if prof.get("bodies_per_warp_step") == st["width"] else None
    t = ph.get("traverse")
    achieved = ips * st["warp_steps"] / (t * 1e-3) / 1e9 if ips and t else None
    hbm_peak = ctx.peaks.get("hbm_gbs", 6650.0)
    r = {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "Gwarp-inst/s",
         "frac": achieved / peak if achieved else None, "kernel": prof.get("kernel", f"{cfg.algorithm} walk"),
         "kernel_ms": t, "inst_per_step": ips, "inst_per_step_source": prof.get("source"),
         "warp_steps": st["warp_steps"], "tests_per_body": st["node_visits"] / max(1, targets),
         "interactions_per_body": st["interactions"] / max(1, targets),
         "bodies_per_warp_step": st["width"], "lane_utilisation": st["node_visits"] / max(1, st["width"] * st["warp_steps"]),
         "traffic": prof.get("dram_bytes") if prof.get("n") == n else None,
         "l2_hit_pct": prof.get("l2_hit_pct"),
         "note": f"peak = {ctx.sm_count} SMs x 4 schedulers x {mhz:.0f} MHz (SM clock sampled during the timed region); "
                 "achieved = instructions per warp step (ncu capture under profiles/) x warp steps counted live / walk time"}
    if r["traffic"] and t:
        r["dram_gbs"] = r["traffic"] / (t * 1e-3) / 1e9
        r["hbm_frac"] = r["dram_gbs"] / hbm_peak
    # the phases that really stream HBM: algorithmic bytes / phase time / measured HBM peak
    phases = []
    for name, b in hbm_phase_bytes(cfg, n, isz, ctx.world).items():
        ms = ph.get(name)
        if ms:
            gbs = b / (ms * 1e-3) / 1e9
            phases.append({"phase": name, "bound": "hbm", "bytes": int(b), "ms": ms, "achieved": gbs, "peak": hbm_peak,
                           "unit": "GB/s", "frac": gbs / hbm_peak})
    r["hbm_phases"] = phases
    return r


def measure(ctx, cfg, steps, warmup, want_e2e, want_cpu, s=None):
    """One bench measurement of `cfg` (argparse-like: algorithm, precision, dim, n, theta) on ctx.world GPUs."""
    torch, nbx = ctx.torch, ctx.nbx
    rank, world = ctx.rank, ctx.world
    dt = np.float32 if cfg.precision == "float" else np.float64
    dim = cfg.dim
    metric, unit = metric_unit(cfg)
    if s is None:
        s = make_state(cfg.n, dt, dim)
    n = len(s["m"])
    eng = ctx.new_engine(s, cfg)

    # ---- device-resident throughput -------------------------------------------------------------------------------
    with ClockSampler(ctx.local_rank) as clocks:
        for _ in range(warmup):
            eng.step(1)
        eng.sync()
        c0 = eng.counters()
        step_ms = []
        t_region0 = time.time()
        for _ in range(steps):
            ctx.flush_l2()
            ctx.barrier()
            step_ms.append(eng.step_timed(1))  # CUDA events on the engine stream around the whole step
        ctx.barrier()
        t_region1 = time.time()
    c1 = eng.counters()
    total_ms = ctx.max_over_ranks(sum(step_ms))
    value = units_per_step(cfg, n) * steps / (total_ms * 1e-3)
    launches = c1["kernel_launches"] - c0["kernel_launches"]

    # per-phase device times of one more step (dominant kernel duration for the roofline)
    eng.set_phase_timing(True)
    ctx.flush_l2()
    eng.step_timed(1)
    ph = eng.phase_ms()
    eng.set_phase_timing(False)

    tree_stats = None
    ts_, te_ = nbx.shard_bounds(n, rank, world)
    if cfg.algorithm in ("octree", "bvh"):
        tree_stats = eng.traversal_stats()  # counting re-run of the last tree's walk (outside every timed region)

    # ---- end to end through the C ABI with host buffers --------------------------------------------------------------
    e2e = None
    if want_e2e:
        import ctypes as C

        def pinned(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t, t.numpy()
        keep, host = {}, {}
        for k in ("m", "x", "v", "a", "ao"):
            keep[k], host[k] = pinned(s[k])
        lib = nbx.lib()

        def e2e_step():
            # H2D of the step's inputs / D2H of the result, straight into the pinned buffer the next step uploads from
            # (nbx_upload has returned, so the old contents are no longer needed). On N GPUs every rank moves ITS shard
            # of the bodies over PCIe (nbx_upload_shard all-gathers the shards over NVLink): the job still moves every
            # input byte to the devices and every result byte back once per step.
            if world > 1:
                eng.upload_shard(host["m"], host["x"], host["v"], host["a"], host["ao"])
                eng.step(1)
                rc = lib.nbx_download_shard(eng._h, None, host["x"].ctypes.data_as(C.c_void_p), None, None, None)
            else:
                eng.upload(host["m"], host["x"], host["v"], host["a"], host["ao"])
                eng.step(1)
                rc = lib.nbx_download(eng._h, None, host["x"].ctypes.data_as(C.c_void_p), None, None, None)
            assert rc == 0
        for _ in range(max(1, warmup - 1)):
            e2e_step()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            e2e_step()
        ctx.barrier()
        e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
        isz = np.dtype(dt).itemsize
        e2e = {"value": units_per_step(cfg, n) * steps / e2e_s, "unit": unit,
               "h2d_bytes_per_step": int(n * (1 + 4 * dim) * isz), "d2h_bytes_per_step": int(n * dim * isz),
               "ms_per_step": 1e3 * e2e_s / steps,
               "io": "nbx_upload + nbx_download, pinned host buffers" if world == 1 else
                     f"nbx_upload_shard + nbx_download_shard: each rank moves its 1/{world} of the bytes over PCIe, shards "
                     "all-gathered over NVLink (bytes below are the job totals)"}
    eng.close()

    if cfg.algorithm == "all-pairs" and n >= 16384:
        parallelism = f"block-pair units dealt round-robin x{world}, NCCL all-reduce of accelerations"
    elif cfg.algorithm.startswith("all-pairs"):
        parallelism = f"targets sharded x{world}, NCCL all-gather of accelerations"
    else:
        parallelism = (f"traversal sharded x{world}, accelerations stored into the peers' arrays by the walk kernel (P2P over "
                       "NVLink) + barrier; sharded Hilbert sort, replicated tree build" if cfg.algorithm == "bvh" and getattr(ctx, "peer", False)
                       else f"replicated tree build, traversal sharded x{world}, NCCL all-gather of accelerations")
    if rank != 0:
        return None, s
    clk = clocks.summary(t_region0, t_region1)
    if cfg.algorithm.startswith("all-pairs"):
        roofline = roofline_all_pairs(ctx, cfg, n, dim, dt, ph, world)
        try:  # DRAM traffic of the dominant kernel from the committed ncu capture of this exact workload (1 GPU), if any
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(workload_name(cfg, n))
            if tr and world == 1:
                roofline["traffic"] = tr["bytes"]
                roofline["traffic_source"] = "profiles/" + tr["source"]
                if roofline.get("kernel_ms"):  # what actually reaches HBM (ncu) over the kernel time measured here
                    roofline["dram_gbs"] = tr["bytes"] / (roofline["kernel_ms"] * 1e-3) / 1e9
        except (OSError, ValueError):
            pass
    else:
        roofline = roofline_tree(ctx, cfg, n, dim, dt, ph, tree_stats, te_ - ts_, clk)
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if dt == np.float32 else "f64",
            "data": "synthetic (reference galaxy model, mt19937{42}, generated on the host by the product's driver)",
            "config": {"workload": workload_name(cfg, n), "parallelism": parallelism,
                       "l2": "512 MiB flush write between timed steps", "phase_ms": ph},
            "clocks": clk, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline}
    if want_cpu and world == 1:
        line["cpu_baseline"] = cpu_baseline(cfg)
    return line, s


def cpu_one_thread(cfg, line):
    """Context only: the same reference source on ONE thread (serial PSTL backend), smaller sample."""
    _, unit = metric_unit(cfg)
    try:
        exe1 = next((os.path.join(ROOT, "oracle", "_ref", f"nbody_d{cfg.dim}{sfx}") for sfx in ("_native", "")
                     if os.access(os.path.join(ROOT, "oracle", "_ref", f"nbody_d{cfg.dim}{sfx}"), os.X_OK)), None)
        if exe1 and line["cpu_baseline"].get("cores", 1) > 1:
            n1 = sample_n(cfg, 1)
            try:
                secs, _ = run_reference_once(cfg, exe1, n1, 1, 1)
            except (OSError, subprocess.CalledProcessError):  # -march=native built elsewhere
                exe1 = os.path.join(ROOT, "oracle", "_ref", f"nbody_d{cfg.dim}")
                secs, _ = run_reference_once(cfg, exe1, n1, 1, 1)
            return {"value": units_per_step(cfg, n1) / secs, "unit": unit, "cores": 1, "kind": "reference",
                    "sample": f"{workload_name(cfg, n1)}, 1 step; {ref_descr(exe1, 1)}"}
    except Exception as ex:  # noqa: BLE001 - the extra figure must never break the bench line
        return {"error": str(ex)}
    return None


# The other BASELINE.json configurations that fit one GPU (the reference's own sweep times all four algorithms,
# ci/benchmark:64-98): measured after the headline's timed region, few steps each, reported under "configs".
EXTRA_CONFIGS = {
    "C3": dict(algorithm="all-pairs-collapsed", precision="double", dim=3, n=262144, theta=0.5),
    "C4": dict(algorithm="octree", precision="double", dim=3, n=10_000_000, theta=0.5),
    "bvh_f32_n10M": dict(algorithm="bvh", precision="float", dim=3, n=10_000_000, theta=0.5),
}


def is_headline(args):
    return (args.algorithm, args.precision, args.dim, args.n) == ("all-pairs", "float", 3, 1_000_000)


# ---- N > 1: results of the sharded run against a single-GPU run of the same problem ------------------------------------
VERIFY_CASES = (  # (algorithm, precision, n, dim, steps, octree lanes in Hilbert order)
    ("all-pairs", "float", 9001, 3, 3, 0), ("all-pairs", "float", 70001, 3, 3, 0), ("all-pairs-collapsed", "double", 3001, 3, 3, 0),
    ("bvh", "float", 50021, 3, 3, 0), ("octree", "double", 40009, 3, 3, 0), ("bvh", "double", 7001, 2, 3, 0),
    ("octree", "float", 30011, 3, 3, 1),
    ("all-pairs", "float", 1_000_000, 3, 1, 0),   # C2 shape
    ("bvh", "float", 10_000_000, 3, 1, 0),        # C5 shape (per-GPU share of the walk at N = 8 is 1.25 M targets)
)


def verify_multi(ctx, cases=VERIFY_CASES, log=None):
    """Every rank compares the replicated state of the N-GPU engine with a single-GPU engine run on its own device.
    Trees must agree bit-for-bit (keys, permutation, x, v, a, ao: the per-body arithmetic does not depend on the
    sharding); all-pairs agrees to rounding (the partial sums are grouped differently). Also sum_i m_i a_i = 0."""
    import argparse
    nbx, torch = ctx.nbx, ctx.torch
    results, ok_all = [], True
    for algo, prec, n, dim, steps, hil in cases:
        os.environ["NBX_OCT_HILBERT"] = "1" if hil else ("0" if algo == "octree" and n < (4 << 20) else "")
        if not os.environ["NBX_OCT_HILBERT"]:
            del os.environ["NBX_OCT_HILBERT"]
        cfg = argparse.Namespace(algorithm=algo, precision=prec, dim=dim, n=n, theta=0.5)
        dt = np.float32 if prec == "float" else np.float64
        s = make_state(n, dt, dim)
        tree = algo in ("bvh", "octree")
        extra_m = extra_s = None
        with ctx.new_engine(s, cfg, multi=True) as e:
            e.step(steps)
            multi = e.download()
            if algo == "bvh":
                extra_m = e.bvh_keys()
        with ctx.new_engine(s, cfg, multi=False) as e:
            e.step(steps)
            single = e.download()
            if algo == "bvh":
                extra_s = e.bvh_keys()
        good, detail = True, {}
        for k in ("x", "v", "a", "ao"):
            if tree:
                same = multi[k].tobytes() == single[k].tobytes()
                detail[k] = "bit-exact" if same else f"max abs diff {np.abs(multi[k].astype(np.float64) - single[k]).max():.3e}"
                good &= same
            else:
                a, b = multi[k].astype(np.float64), single[k].astype(np.float64)
                err = np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-300)
                rms = float(np.sqrt((err ** 2).mean()))
                tol = 2e-5 if dt == np.float32 else 1e-12
                detail[k] = f"rms rel diff {rms:.2e} (max {err.max():.2e})"
                good &= rms < tol
        if algo == "bvh":
            same = extra_m[0].tobytes() == extra_s[0].tobytes() and extra_m[1].tobytes() == extra_s[1].tobytes()
            detail["keys,perm"] = "bit-exact" if same else "DIFFER"
            good &= same
        # Newton's third law on the final state: sum m a = 0 up to rounding (all-pairs) / the multipole error (trees);
        # the collapsed algorithm never accumulates z (reference bug kept on purpose), so only x, y are checked there
        nc = 2 if algo == "all-pairs-collapsed" else dim
        ma = multi["m"].astype(np.float64)[:, None] * multi["a"].astype(np.float64)[:, :nc]
        p = float(np.linalg.norm(ma.sum(0)) / np.abs(ma).sum())
        detail["sum_m_a_rel"] = f"{p:.2e}"
        good &= p < (1e-2 if tree else (1e-4 if dt == np.float32 else 1e-10))
        good = bool(good)
        ok_all &= good
        name = f"{algo} {prec} {dim}-D n={n} steps={steps}" + (" hilbert-lanes" if hil else "")
        results.append({"case": name, "ok": good, **detail})
        if log and (ctx.rank == 0 or not good):
            log(f"[rank {ctx.rank}/{ctx.world}] {name}: {'OK' if good else 'FAIL'} {detail}")
    os.environ.pop("NBX_OCT_HILBERT", None)
    # sharded I/O: every rank uploads only its shard; afterwards every rank must hold the complete state
    for prec, dim, n in (("float", 3, 100003), ("double", 2, 5001)):
        cfg = argparse.Namespace(algorithm="all-pairs", precision=prec, dim=dim, n=n, theta=0.5)
        s = make_state(n, np.float32 if prec == "float" else np.float64, dim)
        s["a"] = s["v"] * 3
        s["ao"] = s["x"] * 0.5
        with ctx.new_engine({**s, **{k: np.zeros_like(s[k]) for k in ("m", "x", "v", "a", "ao")}}, cfg, multi=True) as e:
            e.upload_shard(s["m"], s["x"], s["v"], s["a"], s["ao"])
            full = e.download()
            part = {k: np.full_like(s[k], -7) for k in ("m", "x", "v", "a", "ao")}
            e.download_shard_into(**part)
        lo, hi = nbx.shard_bounds(n, ctx.rank, ctx.world)
        good = all(full[k].tobytes() == s[k].tobytes() for k in ("m", "x", "v", "a", "ao"))
        good &= all(part[k][lo:hi].tobytes() == s[k][lo:hi].tobytes() and (part[k][:lo] == -7).all() and (part[k][hi:] == -7).all()
                    for k in ("m", "x", "v", "a", "ao"))
        ok_all &= bool(good)
        name = f"shard upload/download {prec} {dim}-D n={n}"
        results.append({"case": name, "ok": bool(good)})
        if log and (ctx.rank == 0 or not good):
            log(f"[rank {ctx.rank}/{ctx.world}] {name}: {'OK' if good else 'FAIL'}")
    t = torch.tensor([1 if ok_all else 0], device="cuda")
    if ctx.dist:
        ctx.dist.all_reduce(t, op=ctx.dist.ReduceOp.MIN)
    exchange = ("bvh: accelerations stored into the peers' arrays by the walk kernel (CUDA IPC, P2P over NVLink) + barrier"
                if getattr(ctx, "peer", False) else "bvh: NCCL all-gather of the accelerations")
    if log and ctx.rank == 0:
        log(f"[rank 0/{ctx.world}] exchange paths checked: {exchange}; all-pairs: NCCL all-reduce / all-gather; octree: NCCL all-gather")
    return {"ok": bool(int(t.item())), "ranks": ctx.world, "cases": results, "exchange": exchange,
            "what": "N-GPU engine vs single-GPU engine on the same inputs, every rank checks the full replicated state"}


def main_nbx(args):
    import argparse
    ctx = Ctx()
    line, _ = measure(ctx, args, args.steps, args.warmup, not args.no_e2e, not args.no_cpu_baseline)
    if ctx.rank == 0 and line.get("cpu_baseline"):
        one = cpu_one_thread(args, line)
        if one:
            line["cpu_reference_1thread"] = one
    if ctx.world == 1 and is_headline(args) and not args.no_configs:
        line["configs"] = {}
        for name, kw in EXTRA_CONFIGS.items():
            cfg = argparse.Namespace(**kw)
            try:
                sub, _ = measure(ctx, cfg, min(args.steps, 3), 3, False, not args.no_cpu_baseline)
                line["configs"][name] = {k: sub[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "dtype",
                                                             "gpu_launches", "clocks", "roofline") if k in sub}
                line["configs"][name]["workload"] = sub["config"]["workload"]
                line["configs"][name]["phase_ms"] = sub["config"]["phase_ms"]
                if "cpu_baseline" in sub:
                    line["configs"][name]["cpu_baseline"] = sub["cpu_baseline"]
            except Exception as ex:  # noqa: BLE001 - an extra config must never lose the headline line
                line["configs"][name] = {"error": f"{type(ex).__name__}: {ex}"}
    if ctx.world > 1 and not args.no_verify:
        cases = VERIFY_CASES if is_headline(args) else VERIFY_CASES[:7]
        v = verify_multi(ctx, cases, log=lambda m: print(m, file=sys.stderr, flush=True))
        if ctx.rank == 0:
            line["verify"] = v
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    if ctx.dist:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_nbx(a))
