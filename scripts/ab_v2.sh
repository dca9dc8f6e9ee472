#!/bin/bash
# A/B of the second-generation pooling / reducer kernels (NRMS_V2 bit mask, csrc/abi.cu) in ONE box:
# GPU tests with the default, then the cfg2 bench line with the mask 0 (first-generation kernels everywhere),
# the default (reducer everywhere, pooling for small launches) and 7 (second generation everywhere).
# (profiles/r02_v2_ab.txt was taken with an earlier form of the mask: 8 = four gather items in flight.)
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/v2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v2_pytest.log
tail -4 gpurun_out/v2_pytest.log
for m in 0 default 7; do
  if [ $m = default ]; then unset NRMS_V2; else export NRMS_V2=$m; fi
  timeout 90 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/v2_bench_$m.json 2> gpurun_out/v2_bench_$m.err
  echo "== NRMS_V2=$m rc=$?"
  python scripts/show_bench.py gpurun_out/v2_bench_$m.json 2>&1 | grep -E "^value|pool_|reduce_|gather|sum"
done
