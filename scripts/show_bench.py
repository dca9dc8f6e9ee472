#!/usr/bin/env python
"""Pretty-print the last JSON line of a bench.py log: headline numbers + kernel breakdown."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:.1f}  launches {d.get('gpu_launches')}")
print("roofline", d.get("roofline"))
print("clocks", d.get("clocks"), "cpu", d.get("cpu_baseline"))
if d.get("eager_gpu_baseline"):
    print("eager_gpu_baseline", d["eager_gpu_baseline"])
tot = 0.0
for k, v in sorted((d.get("kernel_breakdown") or {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
    tot += v["ms_per_step"]
    print(f"  {k:24s} {v['ms_per_step']:.4f} ms  {100 * v['share']:5.1f}%  x{v['launches_per_step']:.0f}")
print(f"  sum {tot:.4f} ms")
