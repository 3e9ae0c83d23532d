#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json): all-pairs 3-D float galaxy, n = 1M bodies,
G pair-interactions/s, on 1/2/4/8 B200 (targets sharded by rank, positions all-gathered over NCCL each step).

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
def main_nbx(args):
    import torch

    import _pkg
    nbx = _pkg.load().nbx
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if nbx.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — libnbx has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dt = np.float32 if args.precision == "float" else np.float64
    n, dim = args.n, args.dim
    metric, unit = metric_unit(args)

    # synthetic input: the reference's galaxy model (bit-identical restatement; generated on the host once)
    from oracle import oracle as O  # checker library, used here only as the workload generator + cpu_baseline
    s = O.Oracle(fast=True).galaxy(n, dt, dim)
    n = len(s["m"])

    eng = nbx.Engine(n, dim, dt, args.algorithm, s["dt"], s["G"], theta=args.theta, device=local_rank, rank=rank,
                     world_size=world)
    if world > 1:
        ids = [nbx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eng.comm_init_rank(ids[0])
    eng.upload_state(s)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------------------------
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            eng.step(1)
        eng.sync()
        c0 = eng.counters()
        step_ms = []
        t_region0 = time.time()
        for _ in range(args.steps):
            flush_l2()
            barrier()
            step_ms.append(eng.step_timed(1))  # CUDA events on the engine stream around the whole step
        barrier()
        t_region1 = time.time()
    c1 = eng.counters()
    total_ms = max_over_ranks(sum(step_ms))
    value = units_per_step(args, n) * args.steps / (total_ms * 1e-3)
    launches = c1["kernel_launches"] - c0["kernel_launches"]

    # dominant kernel duration (force phase) for the roofline
    eng.set_phase_timing(True)
    flush_l2()
    eng.step_timed(1)
    ph = eng.phase_ms()
    eng.set_phase_timing(False)

    tree_stats = None
    ts_, te_ = nbx.shard_bounds(n, rank, world)
    if args.algorithm in ("octree", "bvh"):
        tree_stats = eng.traversal_stats()  # counting re-run of the last tree's walk (outside every timed region)

    # ---- end to end through the C ABI with host buffers --------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        def pinned(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t, t.numpy()
        keep, host = {}, {}
        for k in ("m", "x", "v", "a", "ao"):
            keep[k], host[k] = pinned(s[k])
        import ctypes as C
        lib = nbx.lib()

        def e2e_step():
            eng.upload(host["m"], host["x"], host["v"], host["a"], host["ao"])   # H2D of the step's inputs
            eng.step(1)
            # D2H of the result, straight into the pinned buffer the next step uploads from (nbx_upload has returned, so
            # the old contents are no longer needed)
            rc = lib.nbx_download(eng._h, None, host["x"].ctypes.data_as(C.c_void_p), None, None, None)
            assert rc == 0
        for _ in range(max(1, args.warmup - 1)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        isz = np.dtype(dt).itemsize
        e2e = {"value": units_per_step(args, n) * args.steps / e2e_s, "unit": unit,
               "h2d_bytes_per_step": int(n * (1 + 4 * dim) * isz), "d2h_bytes_per_step": int(n * dim * isz),
               "ms_per_step": 1e3 * e2e_s / args.steps}

    if args.algorithm == "all-pairs" and n >= 16384:
        parallelism = f"block-pair units dealt round-robin x{world}, NCCL all-reduce of accelerations"
    elif args.algorithm.startswith("all-pairs"):
        parallelism = f"targets sharded x{world}, NCCL all-gather of positions"
    else:
        parallelism = f"replicated tree build, traversal sharded x{world}, NCCL all-gather of accelerations"
    line = None
    if rank == 0:
        clk = clocks.summary(t_region0, t_region1)
        roofline = None
        if args.algorithm.startswith("all-pairs"):
            prec = nbx.F32 if dt == np.float32 else nbx.F64
            peak = nbx.measure_fma_peak(prec, local_rank)
            flops = n * (n - 1) * FLOP_PER_PAIR[dim] / world
            achieved = flops / (ph["force"] * 1e-3) / 1e12 if ph.get("force") else None
            roofline = {"bound": "fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak if achieved and peak else None, "traffic": None,
                        "kernel": "all_pairs_sym_kernel" if (n >= 16384 and args.algorithm == "all-pairs") else "all_pairs_kernel", "kernel_ms": ph.get("force"),
                        "note": f"{FLOP_PER_PAIR[dim]:.0f} algorithmic flop per ORDERED pair x n(n-1)/ranks / force time "
                                "(pair kernel + partial-sum reduction); for n >= 16384 the kernel evaluates each unordered "
                                "pair once (Newton's third law, src/all_pairs.h:41-42 TODO) and applies it to both bodies, in float with "
                                "FP32x2 (FFMA2) pair arithmetic; peak = "
                                f"{'FFMA' if prec == nbx.F32 else 'DFMA'} microbenchmark measured in this run "
                                "(MEASURED_PEAKS.json has no FP32/FP64 FMA figure)"}
        else:
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except OSError:
                pass
            peak = peaks.get("hbm_gbs", 6650.0)
            # dominant kernel = the traversal; algorithmic bytes = (body, node) tests x node record size
            # (octree monopole + link record 24/40 B, bvh monopole + width 20/40 B)
            isz = np.dtype(dt).itemsize
            rec = (4 * isz + 8) if args.algorithm == "octree" else (4 * isz + isz)
            st = tree_stats
            gbs = st["node_visits"] * rec / (ph["traverse"] * 1e-3) / 1e9 if ph.get("traverse") else None
            roofline = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                        "frac": gbs / peak if gbs else None, "traffic": None, "kernel": f"{args.algorithm}_force kernel",
                        "kernel_ms": ph.get("traverse"), "node_bytes": rec,
                        "visits_per_body": st["node_visits"] / max(1, te_ - ts_),
                        "interactions_per_body": st["interactions"] / max(1, te_ - ts_),
                        "lane_utilisation": st["node_visits"] / max(1, 32 * st["warp_steps"]),
                        "note": ("achieved = requested node bytes (visits x record) / traversal time: an optimistic bound, "
                                 "the records are served mostly from L1/L2; peak = "
                                 + ("MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"))}
        try:  # DRAM traffic of the dominant kernel from the committed ncu capture of this exact workload (1 GPU), if any
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(workload_name(args, n))
            if tr and world == 1:
                roofline["traffic"] = tr["bytes"]
                roofline["traffic_source"] = "profiles/" + tr["source"]
                if roofline.get("kernel_ms"):  # what actually reaches HBM (ncu) over the kernel time measured here
                    roofline["dram_gbs"] = tr["bytes"] / (roofline["kernel_ms"] * 1e-3) / 1e9
        except (OSError, ValueError):
            pass
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32" if dt == np.float32 else "f64",
                "data": "synthetic (reference galaxy model, mt19937{42}, generated on the host)",
                "config": {"workload": workload_name(args, n), "parallelism": parallelism,
                           "l2": "512 MiB flush write between timed steps", "phase_ms": ph},
                "clocks": clk, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args)
            try:  # context only: the same reference source on ONE thread (serial PSTL backend), smaller sample
                exe1 = next((os.path.join(ROOT, "oracle", "_ref", f"nbody_d{args.dim}{sfx}") for sfx in ("_native", "")
                             if os.access(os.path.join(ROOT, "oracle", "_ref", f"nbody_d{args.dim}{sfx}"), os.X_OK)), None)
                if exe1 and line["cpu_baseline"].get("cores", 1) > 1:
                    n1 = sample_n(args, 1)
                    try:
                        secs, _ = run_reference_once(args, exe1, n1, 1, 1)
                    except (OSError, subprocess.CalledProcessError):  # -march=native built elsewhere
                        exe1 = os.path.join(ROOT, "oracle", "_ref", f"nbody_d{args.dim}")
                        secs, _ = run_reference_once(args, exe1, n1, 1, 1)
                    line["cpu_reference_1thread"] = {"value": units_per_step(args, n1) / secs, "unit": unit, "cores": 1,
                                                     "kind": "reference",
                                                     "sample": f"{workload_name(args, n1)}, 1 step; {ref_descr(exe1, 1)}"}
            except Exception as ex:  # noqa: BLE001 - the extra figure must never break the bench line
                line["cpu_reference_1thread"] = {"error": str(ex)}
        print(json.dumps(line), flush=True)
    eng.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_nbx(a))
