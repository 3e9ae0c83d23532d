set -x
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/p_ap.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:all_pairs_sym_kernel -s 2 -c 1 -f -o gpurun_out/r01d_allpairs_sym_packed_f32_n1M $CMD > gpurun_out/n_ap.log 2>&1; tail -2 gpurun_out/n_ap.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/p_ap2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_allpairs_sym_packed_f32_n1M.csv $CMD > gpurun_out/n_ap2.log 2>&1; tail -2 gpurun_out/n_ap2.log
