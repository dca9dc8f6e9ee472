#!/bin/bash
# Final single-GPU evidence of round 2, last build (run under gpurun): tests, the bench lines of every config,
# the CPU reference arm, cached-vector scoring, batch assembly, the sibling plugin, then the two cheap ncu
# passes (launch list of the bench command; DRAM bytes + time of every kernel of three steps).  The
# `--set full` captures of the GEMM / attention kernels (scripts/profile_round2.sh passes 3-4) are not
# repeated: the fp32-grade instantiations they profile did not change after they were taken.
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1; tail -3 $O/r02_pytest_gpu.log
timeout 400 python bench.py --steps 100 --warmup 5 > $O/r02_bench_cfg2_n1.json 2> $O/r02_bench_cfg2_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err
timeout 300 python bench.py --zipf --steps 50 --no-extras --no-cpu-baseline > $O/r02_bench_cfg2_zipf_n1.json 2>/dev/null
timeout 300 python bench.py --config cfg3 --steps 20 --no-extras --no-cpu-baseline > $O/r02_bench_cfg3_n1.json 2>/dev/null
timeout 300 python bench.py --config cfg5 --steps 15 --no-extras --no-cpu-baseline > $O/r02_bench_cfg5_n1.json 2>/dev/null
timeout 300 python bench.py --config cfg5 --gemm-mode 1 --steps 15 --no-extras --no-cpu-baseline > $O/r02_bench_cfg5_fp32_n1.json 2>/dev/null
timeout 400 python scripts/eval_bench.py 1000000 8192 1 > $O/r02_eval_cfg4.json 2> $O/r02_eval_cfg4.err
timeout 400 python scripts/eval_bench.py 1000000 8192 2 > $O/r02_eval_cfg4_bf16.json 2> $O/r02_eval_cfg4_bf16.err
timeout 200 python scripts/data_bench.py > $O/r02_data_bench.json 2>/dev/null
timeout 200 python scripts/variant_bench.py 64 50 5 > $O/r02_variant_nrms_b64.json 2>/dev/null
timeout 200 python scripts/variant_bench.py 512 40 5 > $O/r02_variant_nrms_b512.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/plain_bench.log 2>&1 || { tail -5 $O/plain_bench.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/ncu_launch.log 2>&1
python scripts/prof_step.py > $O/plain_step.log 2>&1 || { tail -5 $O/plain_step.log; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -c 800 --csv --log-file $O/r02_step_metrics.csv python scripts/prof_step.py > $O/ncu_step.log 2>&1
ls -la $O/r02_*.csv $O/r02_*.json | tail -20
