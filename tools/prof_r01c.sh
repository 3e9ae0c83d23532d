set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct"
CMD="python bench.py --algorithm bvh -n 10000000 --precision float --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/p_bvh.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bvh_force_key_kernel -s 1 -c 1 -f -o gpurun_out/r01c_bvh_force_key_f32_n10M $CMD > gpurun_out/n_bvh.log 2>&1; tail -2 gpurun_out/n_bvh.log
$CMD > gpurun_out/p_bvh2.log 2>&1 && ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches_bvh_f32_n10M.csv $CMD > gpurun_out/n_bvh2.log 2>&1; tail -2 gpurun_out/n_bvh2.log
CMD="python bench.py --algorithm octree -n 10000000 --precision float --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/p_octf.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:octree_force_kernel -s 1 -c 1 -f -o gpurun_out/r01c_octree_force_f32_n10M $CMD > gpurun_out/n_octf.log 2>&1; tail -2 gpurun_out/n_octf.log
