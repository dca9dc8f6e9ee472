#!/bin/bash
# ncu --set full of the kernels added at the end of round 2 (run under gpurun; each ncu pass after the same
# command has exited 0 without ncu):
#   1. cfg5's 48-row attention tiles (attn_hpn_{fwd,bwd}_kernel<1, 48>: the news encoder's launches of a step)
#   2. the sibling plugin's masked attention (masked_attn_{fwd,bwd}_kernel<2, 2>, batch 512)
set -x
O=gpurun_out
CONFIG=cfg5 STEPS=2 python scripts/prof_step.py > $O/plain_step_cfg5.log 2>&1 || { tail -5 $O/plain_step_cfg5.log; exit 1; }
CONFIG=cfg5 STEPS=2 ncu --set full --clock-control none --import-source on -k regex:attn_hpn --launch-skip 2 --launch-count 2 \
    -o $O/prof_attn48 -f python scripts/prof_step.py > $O/ncu_attn48.log 2>&1
ncu -i $O/prof_attn48.ncu-rep --page raw --csv > $O/r02_attn48_raw.csv
ncu -i $O/prof_attn48.ncu-rep --page source --csv > $O/r02_attn48_src.csv
rm -f $O/prof_attn48.ncu-rep
python scripts/variant_bench.py 512 3 2 > $O/plain_variant.log 2>&1 || { tail -5 $O/plain_variant.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:masked_attn --launch-skip 4 --launch-count 2 \
    -o $O/prof_masked -f python scripts/variant_bench.py 512 3 2 > $O/ncu_masked.log 2>&1
ncu -i $O/prof_masked.ncu-rep --page raw --csv > $O/r02_masked_raw.csv
ncu -i $O/prof_masked.ncu-rep --page source --csv > $O/r02_masked_src.csv
rm -f $O/prof_masked.ncu-rep
tail -2 $O/ncu_attn48.log $O/ncu_masked.log
ls -la $O/r02_attn48_*.csv $O/r02_masked_*.csv
