#!/bin/bash
# usage: scripts/ab.sh "ENV1=.. ENV2=.." ...  -> one bench run per argument, prints attention/gemm lines
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 30 --warmup 5 --no-cpu-baseline > /tmp/ab.log 2>/tmp/ab.err || tail -3 /tmp/ab.err
  python scripts/show_bench.py /tmp/ab.log 2>/dev/null | grep -E "^value|attn_|gemm_fwd_qkv"
done
