#!/bin/bash
# Final single-GPU evidence of round 2 (run under gpurun): tests, the bench lines of every config, the CPU
# reference arm, cached-vector scoring, batch assembly, then the ncu passes (scripts/profile_round2.sh).
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
bash scripts/profile_round2.sh > $O/r02_profile.log 2>&1
tail -3 $O/r02_profile.log
