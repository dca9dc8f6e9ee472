#!/usr/bin/env python
"""Turns an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
launch list of `scripts/prof_step.py` (one train step captured) into profiles/ncu_traffic.json:
DRAM bytes and serialized device time per step for each library kernel label bench.py reports.

    python scripts/ncu_traffic.py gpurun_out/step_metrics.csv profiles/ncu_traffic.json [steps_in_capture]
"""
import csv
import json
import sys

LABELS = [   # (substring of the ncu function name, bench.py label)
    ("attn_mma_bwd", "attn_bwd"), ("attn_bwd_kernel", "attn_bwd"), ("attn_hpn_bwd", "attn_bwd"), ("attn_hpl_bwd", "attn_bwd"),
    ("attn_mma_fwd", "attn_fwd"), ("attn_fwd_kernel", "attn_fwd"), ("attn_hp_fwd", "attn_fwd"), ("attn_hpn_fwd", "attn_fwd"),
    ("attn_hpl_fwd", "attn_fwd"),
    ("ig_gemm_kernel<0, 0, 240, 0", "gemm_fwd_qkv"), ("ig_gemm_kernel<0, 0, 256, 6", "gemm_fwd_qkv"),
    ("ig_gemm_kernel<0, 0, 208, 1", "gemm_fwd_additive"),
    ("ig_gemm_kernel<0, 1, 320, 5", "gemm_dgrad_additive"), ("ig_gemm_kernel<0, 1, 64, 5", "gemm_dgrad_additive"),
    ("ig_gemm_kernel<0, 1, 320, 3", "gemm_dgrad_qkv"), ("ig_gemm_kernel<0, 1, 64, 3", "gemm_dgrad_qkv"),
    ("ig_gemm_kernel<1, 1, 320, 4", "gemm_wgrad"), ("gather_rows_img", "gather"), ("pool_fwd", "pool_fwd"),
    ("pool_bwd", "pool_bwd"), ("adam_kernel", "adam"), ("embgrad_reduce", "embgrad_reduce"), ("embgrad_fixup", "embgrad_reduce"), ("rsort_", "plan"), ("reduce_wgrad", "reduce_wgrad"),
    ("reduce_rows", "reduce_rows"), ("img_pack", "img_pack"), ("score_kernel", "score_1"), ("zero_kernel", "zero"), ("plan_", "plan"),
]


def label(name):
    for sub, lab in LABELS:
        if sub in name:
            return lab
    return None


def main(src, dst, steps=1):
    rows = [r for r in csv.reader(open(src)) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    out = {}
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv:
            continue
        lab = label(r[kn])
        if lab is None:
            continue
        rec = out.setdefault(lab, {"dram_bytes_per_step": 0.0, "device_us_per_step": 0.0, "launches_per_step": 0})
        val = float(r[mv].replace(",", ""))
        unit = r[mu].lower()
        if r[mn].startswith("dram__bytes"):
            scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
            rec["dram_bytes_per_step"] += val * scale
        elif r[mn].startswith("gpu__time_duration"):
            scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1.0)
            rec["device_us_per_step"] += val * scale
            rec["launches_per_step"] += 1
    for rec in out.values():
        rec["dram_bytes_per_step"] /= steps
        rec["device_us_per_step"] /= steps
        rec["launches_per_step"] /= steps
    json.dump({"source": src, "steps_in_capture": steps, "note": "one fused train step (cfg2), ncu per-launch metrics, cold-cache and serialised",
               "kernels": out}, open(dst, "w"), indent=1, sort_keys=True)
    for k, v in sorted(out.items(), key=lambda kv: -kv[1]["device_us_per_step"]):
        print(f"{k:22s} {v['device_us_per_step']:9.1f} us  {v['dram_bytes_per_step'] / 1e6:9.1f} MB  x{v['launches_per_step']}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 1)
