"""Runs the forward-orientation GEMM self-test at the benchmark's shape, once per variant given on the command
line (0 = one CTA per SM, 3 = CTA pairs), a few times each, so that ncu can capture them and CUDA events can
time them:   python scripts/micro/gemm_pair_probe.py 0 3"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from pytorch_news_recommender_b200 import ops, _build
_build.build()
M, N, K = 105600, 960, 300
torch.manual_seed(0)
A = torch.randn(M, K, device="cuda")
B = torch.randn(N, K, device="cuda")
for v in map(int, sys.argv[1:] or ["0", "3"]):
    for _ in range(3):
        out = ops.gemm_selftest(v, A, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = ops.gemm_selftest(v, A, B)
    e1.record()
    torch.cuda.synchronize()
    print(f"variant {v}: {e0.elapsed_time(e1) / 10:.4f} ms per call (pack + GEMM)", flush=True)
print("ok")
