#!/usr/bin/env python
"""The reference's benchmark sweep (ci/benchmark:64-98) run through the nbx host driver.

ci/benchmark times, per toolchain, `./ci/run <compiler> <algorithm> galaxy 3 double <bodies> 200` with NO_SAVE=1, i.e.
    nbody_d3 -s 200 -n <bodies> --save none --csv-total --algorithm <a> --workload galaxy --precision double
(ci/run:144-176) for all four algorithms at n = 100 000 and for octree + bvh at n = 1 000 000, after logging the GPU,
driver, CPU model, core count and hostname; ci/data.py:48-60 scrapes that log into one CSV. This script produces the SAME
log (same banner lines, `compiler:nbx`, the driver's own CSV header/rows) from stdpar-nbody_b200/bin/nbody_d3, so the
reference's scraper and plotting pipeline work on it unchanged:

    python tools/ref_sweep.py > sweep.log && python /root/reference/ci/data.py sweep.log

--reference also runs the unmodified reference CPU build (oracle/_ref/nbody_d3_omp, all host cores) over the same matrix
with fewer steps, as `compiler:gcc-omp-shim`. parse_log() restates ci/data.py for boxes where the reference tree is absent
(the GPU box); tests/test_ref_sweep.py checks it against the real ci/data.py on a committed log.
"""
import argparse
import os
import platform
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALGOS_ALL = ("all-pairs", "octree", "bvh", "all-pairs-collapsed")  # ci/benchmark:24
ALGOS_LARGE = ("octree", "bvh")                                    # ci/benchmark:80


def sh(cmd):
    try:
        return subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
    except (OSError, subprocess.TimeoutExpired):
        return ""


def banner(out):
    smi = sh(["nvidia-smi", "--query-gpu=gpu_name,driver_version", "--format=csv"])  # ci/benchmark:42-44
    if smi.strip():
        out(smi.strip().splitlines()[0])
        out(smi.strip().splitlines()[1])
    for ln in sh(["lscpu"]).splitlines():                                             # ci/benchmark:48-49
        if "Model name" in ln or "Core(s) per socket" in ln:
            out(ln)
    out(f"hostname:{platform.node()}")                                                # ci/benchmark:50


def run_matrix(exe, label, out, steps, small, large, precision, threads=None):
    env = dict(os.environ)
    if threads:
        env.update(OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
    for bodies, algos in ((small, ALGOS_ALL), (large, ALGOS_LARGE)):
        if not bodies:
            continue
        for a in algos:
            out(f"compiler:{label}")
            cmd = [exe, "-s", str(steps), "-n", str(bodies), "--save", "none", "--csv-total", "--algorithm", a,
                   "--workload", "galaxy", "--precision", precision]
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if r.returncode != 0:
                raise SystemExit(f"{' '.join(cmd)} failed:\n{r.stdout}{r.stderr}")
            for ln in r.stdout.splitlines():
                out(ln)


def parse_log(lines):
    """Restatement of /root/reference/ci/data.py:20-60: log lines -> CSV lines."""
    res, parsed_gpu, found = [], 0, False
    gpu = driver = cpu = cores = compiler = hostname = None
    sequential = False
    for l in lines:
        if not l.endswith("\n"):
            l += "\n"
        if l.startswith("+"):
            continue
        elif parsed_gpu == 1:
            gpu, driver, parsed_gpu = l.split(", ")[0].strip(), l.split(", ")[1].strip(), 2
        elif l.startswith("Vendor"):
            continue
        elif l.startswith("Model name:"):
            cpu = l.split("Model name:")[1].strip()
        elif l.startswith("Core"):
            cores = l.split("Core(s) per socket:")[1].strip()
        elif l.startswith("name"):
            parsed_gpu = 1
        elif l.startswith("sequential"):
            sequential = True
        elif l.startswith("algorithm"):
            if not found:
                found = True
                res.append(f"gpu,driver,cpu,#cores,seq,compiler,hostname,{l.strip()}")
        elif l.startswith("compiler"):
            compiler = l.split(":")[1].strip()
        elif l.startswith("hostname") or l.startswith("node"):
            hostname = l.split(":")[1].strip().replace(".nvidia.com", "")
        elif l.split(",")[0] in ["octree", "all-pairs", "all-pairs-collapsed", "bvh"]:
            res.append(f"{gpu},{driver},{cpu},{cores},{sequential},{compiler},{hostname},{l.strip()}")
    return res


def main():
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("--steps", type=int, default=200)            # ci/benchmark:16
    ap.add_argument("--small", type=int, default=100_000)        # ci/benchmark:64
    ap.add_argument("--large", type=int, default=1_000_000)      # ci/benchmark:87
    ap.add_argument("--precision", default="double")             # ci/benchmark:14
    ap.add_argument("--dim", type=int, default=3, choices=[2, 3])
    ap.add_argument("--reference", action="store_true", help="also sweep the unmodified reference CPU build")
    ap.add_argument("--reference-steps", type=int, default=11)     # 10 hidden warm-up steps + 1 timed
    ap.add_argument("--reference-small", type=int, default=20_000, help="CPU arm: a bounded sample of the sizes")
    ap.add_argument("--reference-large", type=int, default=200_000)
    ap.add_argument("--reference-only", action="store_true")
    ap.add_argument("--csv", action="store_true", help="print the scraped CSV (ci/data.py format) instead of the raw log")
    args = ap.parse_args()
    log = []

    def out(line):
        log.append(line)
        if not args.csv:
            print(line, flush=True)

    banner(out)
    if not args.reference_only:
        exe = os.path.join(ROOT, "stdpar-nbody_b200", "bin", f"nbody_d{args.dim}")
        run_matrix(exe, "nbx", out, args.steps, args.small, args.large, args.precision)
    if args.reference or args.reference_only:
        exe = os.path.join(ROOT, "oracle", "_ref", f"nbody_d{args.dim}_omp")
        threads = len(os.sched_getaffinity(0))
        run_matrix(exe, "gcc-omp-shim", out, args.reference_steps, args.reference_small, args.reference_large, args.precision, threads)
    if args.csv:
        print("\n".join(parse_log(log)))


if __name__ == "__main__":
    main()
