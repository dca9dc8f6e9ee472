#!/bin/bash
# The ncu evidence kept under profiles/ for one round (run on the GPU box through gpurun):
#   1. launch list of the bench command itself (gpu__time_duration only)
#   2. DRAM bytes + time of every kernel of one train step (scripts/prof_step.py)
#   3. one --set full capture of the dominant kernel ($1 = kernel regex, default attn_hpn_bwd)
K=${1:-attn_hpn_bwd}
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 || { tail -5 gpurun_out/plain_bench.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python scripts/prof_step.py > gpurun_out/plain_step.log 2>&1 || { tail -5 gpurun_out/plain_step.log; exit 1; }
# all three steps of prof_step.py (scripts/ncu_traffic.py divides by 3)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -c 600 --csv --log-file gpurun_out/step_metrics.csv python scripts/prof_step.py > gpurun_out/ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 3 --launch-count 1 -o gpurun_out/prof_top -f \
    python scripts/prof_step.py > gpurun_out/ncu_top.log 2>&1
ncu -i gpurun_out/prof_top.ncu-rep --page raw --csv > gpurun_out/prof_top_raw.csv
ncu -i gpurun_out/prof_top.ncu-rep --page source --csv > gpurun_out/prof_top_src.csv
rm -f gpurun_out/prof_top.ncu-rep
tail -2 gpurun_out/ncu_top.log
