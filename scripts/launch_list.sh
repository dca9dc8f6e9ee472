#!/bin/bash
# per-launch device times of one train step (ncu, gpu__time_duration only) -> gpurun_out/launches_step.csv
python scripts/prof_step.py > gpurun_out/plain_step.log 2>&1 || { tail -5 gpurun_out/plain_step.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step.csv python scripts/prof_step.py > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
