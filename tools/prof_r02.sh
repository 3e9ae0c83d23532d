# Round-2 ncu captures (run on the GPU box through gpurun; every command first runs once WITHOUT ncu and must exit 0).
# Usage: bash tools/prof_r02.sh <tag> <kernel regex> <skip> -- <bench.py args...>
# Writes gpurun_out/<tag>_ncu_full.csv (tools/ncu_summary.py) and gpurun_out/<tag>_source.csv (hottest source lines).
set -u
tag=$1; kern=$2; skip=$3; shift 4
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-configs $*"
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f -o gpurun_out/${tag} $CMD > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
python tools/ncu_summary.py gpurun_out/${tag}.ncu-rep gpurun_out/${tag}_ncu_full.csv > /dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
rm -f gpurun_out/${tag}.ncu-rep
grep -E "gpu__time_duration|smsp__inst_executed.sum|issue_active|pipe_alu|pipe_fma|pipe_xu|pipe_fp64|registers_per_thread|warps_active|dram__bytes_(read|write).sum,|lts__t_sector_hit|thread_inst_executed_per" gpurun_out/${tag}_ncu_full.csv
