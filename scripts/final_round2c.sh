#!/bin/bash
# Last evidence run of round 2 (one gpurun call, one B200): GPU tests, smoke, the default bench line, the
# routing A/B for 49..64-token sequences (NRMS_HPL_MIN_FWD=65: the user encoder's 50-slot history on the
# 64-row attention_hpn forward instead of the key-tiled one) with the parity tests under that routing, and
# the cfg3 / cfg5 single-GPU lines of the final build.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -3 gpurun_out/f_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f_smoke.log
timeout 200 python bench.py --steps 100 --warmup 5 > gpurun_out/f_bench_cfg2.json 2> gpurun_out/f_bench_cfg2.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/f_bench_cfg2.json | head -3
NRMS_HPL_MIN_FWD=65 timeout 90 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/f_bench_hpl65.json 2> gpurun_out/f_bench_hpl65.err
python - <<'P'
import json
for f in ("f_bench_cfg2", "f_bench_hpl65"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    k = d["kernel_breakdown"]["attn_fwd"]
    print(f, "ms/step", round(d["ms_per_step"], 4), "attn_fwd", round(k["ms_per_step"], 4), "user", k.get("user_encoder_ms_per_step"))
P
NRMS_HPL_MIN_FWD=65 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/f_pytest_hpl65.log 2>&1; echo "pytest(hpl65) rc=$?"; tail -1 gpurun_out/f_pytest_hpl65.log
for c in cfg3 cfg5; do
  timeout 120 python bench.py --config $c --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/f_bench_$c.json 2> gpurun_out/f_bench_$c.err; echo "$c rc=$?"
  python scripts/show_bench.py gpurun_out/f_bench_$c.json | head -1
done
