"""Numerics study (CPU, oracle only): what would single-pass tensor-core formats cost in accuracy?

VERDICT r01 asks to compare a single-pass `kind::tf32` GEMM against the 1e-3 logits bound before
keeping the 3-term bf16 split.  The arithmetic of every candidate scheme is emulated here by
rounding the operands of the dense projections (W_Q/W_K/W_V, additive linear) in the oracle and
measuring the eval-mode logits against the fp32 oracle at cfg2's shape.  Run:
    python tests/precision_study.py            -> profiles/r02_precision_study.json
Schemes: operands rounded to tf32 (10-bit mantissa, round-to-nearest and truncation), fp16, bf16,
and the bf16x3 / fp16 "A split, B single" variants.
"""
import json, os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nrms_oracle as O
from pytorch_news_recommender_b200 import synthetic as S


def round_mant(x, bits, trunc=False):
    """keep `bits` explicit mantissa bits of fp32 (tf32: 10)"""
    i = x.contiguous().view(torch.int32)
    drop = 23 - bits
    if trunc:
        return (i & ~((1 << drop) - 1)).view(torch.float32)
    half = 1 << (drop - 1)
    return ((i + half) & ~((1 << drop) - 1)).view(torch.float32)


def q_tf32(x): return round_mant(x, 10)
def q_tf32_trunc(x): return round_mant(x, 10, trunc=True)
def q_fp16(x): return x.half().float()
def q_bf16(x): return x.bfloat16().float()
def ident(x): return x
def q_bf16x2(x):
    hi = x.bfloat16().float()
    return hi + (x - hi).bfloat16().float()


SCHEMES = {
    "fp32": (ident, ident),
    "bf16x3 (hi+lo both operands; drops lo*lo)": (q_bf16x2, q_bf16x2),
    "tf32 single pass, operands rounded to nearest": (q_tf32, q_tf32),
    "tf32 single pass, operands truncated (hardware reads the top 19 bits)": (q_tf32_trunc, q_tf32_trunc),
    "fp16 single pass": (q_fp16, q_fp16),
    "fp16: activations split hi+lo, weights single": (ident, q_fp16),
    "bf16 single pass (gemm_mode 2)": (q_bf16, q_bf16),
}


def main():
    torch.manual_seed(0)
    V, B = 70000, 32
    cfg = O.OracleConfig(30, 50, 4, 300, 10, 200, 0.0, 1e-3)
    sd = O.init_state_dict(cfg, S.make_embedding_table(V, 300, seed=0), seed=42)
    pool = S.make_news_pool(65000, 30, V, seed=0)
    batch = S.make_train_batch(pool, B, 50, 4, seed=7)
    real = batch["candidate_mask"].bool()
    orig = F.linear
    out = {}
    ref = None
    for name, (qa, qb) in SCHEMES.items():
        def lin(x, w, b=None, qa=qa, qb=qb):
            return orig(qa(x), qb(w), b)
        O.F.linear = lin
        try:
            with torch.no_grad():
                lg = O.model_forward(sd, batch, cfg, training=False, per_slot=False)
        finally:
            O.F.linear = orig
        if ref is None:
            ref = lg.double()
            continue
        d = (lg.double() - ref).abs()[real]
        r = ref.abs()[real]
        out[name] = {"max_abs": float(d.max()), "max_rel_to_max": float(d.max() / r.max()),
                     "max_elementwise_rel(clamp 1e-3)": float((d / r.clamp_min(1e-3)).max()),
                     "rms_rel": float(d.pow(2).mean().sqrt() / r.pow(2).mean().sqrt())}
        print(name, out[name])
    json.dump({"shape": "cfg2 (T=30 H=50 K=4 D=300 V=70k), 32 impressions, eval mode", "logits_error_vs_fp32_oracle": out},
              open(os.path.join(ROOT, "profiles", "r02_precision_study.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
