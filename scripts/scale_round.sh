#!/bin/bash
# Multi-GPU measurements of a round on ONE 8-GPU box (run under gpurun --gpus 8):
#   cfg2 at 1 rank, at 8 ranks with the sharded (reduce-scatter + Adam on V/G rows + all-gather) and the dense
#   (all-reduce + replicated Adam) table exchange, sharded at 4 and 2 ranks; cfg3 at 1 and 8; cfg5 at 1, 2, 4, 8;
#   the data-parallel == single-process check at 8 ranks in both modes.
# Every line lands in gpurun_out/r02_scale_*.json (copied to profiles/ afterwards).
cd "$(dirname "$0")/.."
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29500
run() {  # name nproc args...
  local name=$1 n=$2; shift 2
  port=$((port+1))
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 "$@" > $OUT/r02_scale_$name.json 2> $OUT/r02_scale_$name.err
  else
    timeout 300 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n "$@" > $OUT/r02_scale_$name.json 2> $OUT/r02_scale_$name.err
  fi
  echo "$name rc=$? $(grep -o '"value": [0-9.]*' $OUT/r02_scale_$name.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' $OUT/r02_scale_$name.json | head -1)"
}
COMMON="--no-extras --no-cpu-baseline --steps 40"
run cfg2_n1 1 $COMMON
run cfg2_n8_sharded 8 $COMMON --table-sync sharded
run cfg2_n8_dense 8 $COMMON --table-sync dense
run cfg2_n4_sharded 4 $COMMON --table-sync sharded
run cfg2_n2_sharded 2 $COMMON --table-sync sharded
run cfg3_n1 1 --config cfg3 $COMMON
run cfg3_n8_sharded 8 --config cfg3 $COMMON --table-sync sharded
run cfg5_n1 1 --config cfg5 --no-extras --no-cpu-baseline --steps 15
run cfg5_n2 2 --config cfg5 --no-extras --no-cpu-baseline --steps 15 --table-sync sharded
run cfg5_n4 4 --config cfg5 --no-extras --no-cpu-baseline --steps 15 --table-sync sharded
run cfg5_n8 8 --config cfg5 --no-extras --no-cpu-baseline --steps 15 --table-sync sharded
for mode in dense sharded; do
  port=$((port+1))
  TABLE_SYNC=$mode timeout 120 $TR --nproc-per-node 8 --master-port $port scripts/dp_check.py > $OUT/r02_dp_check_${mode}_n8.json 2> $OUT/r02_dp_check_${mode}_n8.err
  echo "dp_check $mode rc=$? $(cat $OUT/r02_dp_check_${mode}_n8.json | tail -1)"
done
