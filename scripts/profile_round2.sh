#!/bin/bash
# The ncu evidence kept under profiles/ for round 2 (run on the GPU box through gpurun; every ncu pass comes
# after the same command has exited 0 without ncu):
#   1. launch list of the bench command itself (gpu__time_duration only)         -> r02_launches_bench.csv
#   2. DRAM bytes + time of every kernel of three train steps (prof_step.py)      -> r02_step_metrics.csv
#   3. --set full of the 12 tcgen05 GEMM launches of one step (6 news + 6 user)   -> r02_gemm_raw.csv / _src.csv
#   4. --set full of the 4 attention launches of one step                         -> r02_attn_raw.csv / _src.csv
set -x
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/plain_bench.log 2>&1 || { tail -5 $O/plain_bench.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launch.log 2>&1
python scripts/prof_step.py > $O/plain_step.log 2>&1 || { tail -5 $O/plain_step.log; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -c 800 --csv --log-file $O/r02_step_metrics.csv python scripts/prof_step.py > $O/ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ig_gemm_kernel --launch-skip 24 --launch-count 12 -o $O/prof_gemm -f \
    python scripts/prof_step.py > $O/ncu_gemm.log 2>&1
ncu -i $O/prof_gemm.ncu-rep --page raw --csv > $O/r02_gemm_raw.csv
ncu -i $O/prof_gemm.ncu-rep --page source --csv > $O/r02_gemm_src.csv
rm -f $O/prof_gemm.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:attn_hp --launch-skip 8 --launch-count 4 -o $O/prof_attn -f \
    python scripts/prof_step.py > $O/ncu_attn.log 2>&1
ncu -i $O/prof_attn.ncu-rep --page raw --csv > $O/r02_attn_raw.csv
ncu -i $O/prof_attn.ncu-rep --page source --csv > $O/r02_attn_src.csv
rm -f $O/prof_attn.ncu-rep
tail -2 $O/ncu_gemm.log $O/ncu_attn.log
ls -la $O/r02_*.csv
