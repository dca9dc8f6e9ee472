import os, sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as e; e.build()
from pytorch_news_recommender_b200 import ops
torch.manual_seed(0)
dev = "cuda:0"
D, Q, h = 300, 200, 10
V = 5000
table = torch.randn(V, D, device=dev); table[0] = 0
params = torch.randn(ops.encoder_param_count(D, Q), device=dev) * 0.05
for (n_seq, L) in ((4, 30), (4, 32), (8, 32), (5, 32), (37, 30), (3520, 30), (4, 30)):
    ids = torch.randint(0, V, (n_seq, L), device=dev)
    shape = ops.EncoderShape(n_seq, L, D, h, Q, V)
    saved0 = torch.zeros(ops.saved_bytes(shape), dtype=torch.uint8, device=dev)
    saved1 = torch.zeros_like(saved0)
    ops.news_encoder_fwd(shape, ids, table, params, saved0, 0.0, 7, 0)
    ops.news_encoder_fwd(shape, ids, table, params, saved1, 0.0, 7, 1)
    torch.cuda.synchronize()
    M = n_seq * L
    qkv0 = saved0[:M*900*4].view(torch.float32).view(M, 900)
    qkv1 = saved1[:M*900*4].view(torch.float32).view(M, 900)
    d = (qkv0-qkv1).abs()
    badrows = (d.amax(1) > 1e-3).nonzero().flatten()
    badcols = (d.amax(0) > 1e-3).nonzero().flatten()
    print((n_seq, L), "M", M, "maxdiff", d.max().item(), "n bad rows", badrows.numel(), badrows[:8].tolist(), badrows[-4:].tolist(),
          "n bad cols", badcols.numel(), badcols[:6].tolist(), badcols[-3:].tolist(), flush=True)
    if badrows.numel():
        r = badrows[0].item()
        print("  row", r, "ref", qkv0[r, :4].tolist(), "got", qkv1[r, :4].tolist())
