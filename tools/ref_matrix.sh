#!/bin/bash
# Reference CPU arm (all host cores, oracle/_ref/nbody_d*_omp) over the BASELINE configs -> gpurun_out/ref_matrix.jsonl
out=gpurun_out/ref_matrix.jsonl; : > $out
nproc >> gpurun_out/ref_matrix_host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/ref_matrix_host.txt
run() { python bench.py --impl reference --steps 1 --warmup 1 "$@" | tail -1 >> $out; }
run --algorithm all-pairs -n 10000 --dim 2 --precision float
run --algorithm all-pairs -n 1000000 --dim 3 --precision float
run --algorithm all-pairs -n 262144 --dim 3 --precision double
run --algorithm all-pairs-collapsed -n 262144 --dim 3 --precision double
run --algorithm octree -n 10000000 --dim 3 --precision double
run --algorithm octree -n 10000000 --dim 3 --precision float
run --algorithm bvh -n 10000000 --dim 3 --precision float
run --algorithm bvh -n 10000000 --dim 3 --precision double
python - <<'PY'
import json
for ln in open("gpurun_out/ref_matrix.jsonl"):
    d = json.loads(ln)
    print(f"{d['config']['sample']:55s} {d['value']:10.4f} {d['unit']:14s} {d['ms_per_step']:10.1f} ms/step cores={d['cpu_baseline']['cores']}")
PY
