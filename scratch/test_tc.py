import os, sys, tempfile, torch, time
sys.path.insert(0, '/root/repo')
import __graft_entry__ as e; e.build()
from pytorch_news_recommender_b200 import ops
torch.manual_seed(0)
dev = "cuda:0"
D, Q, h, L = 300, 200, 10, 30
for n_seq in (4, 37, 3520):
    V = 5000
    table = torch.randn(V, D, device=dev); table[0] = 0
    ids = torch.randint(0, V, (n_seq, L), device=dev)
    params = torch.randn(ops.encoder_param_count(D, Q), device=dev) * 0.05
    shape = ops.EncoderShape(n_seq, L, D, h, Q, V)
    saved0 = torch.empty(ops.saved_bytes(shape), dtype=torch.uint8, device=dev)
    saved1 = torch.empty_like(saved0)
    for p in (0.0, 0.2):
        out0 = ops.news_encoder_fwd(shape, ids, table, params, saved0, p, 7, 0)
        out1 = ops.news_encoder_fwd(shape, ids, table, params, saved1, p, 7, 1)
        torch.cuda.synchronize()
        M = n_seq * L
        qkv0 = saved0[:M*900*4].view(torch.float32).view(M, 900)
        qkv1 = saved1[:M*900*4].view(torch.float32).view(M, 900)
        print(n_seq, p, "qkv maxabs diff", (qkv0-qkv1).abs().max().item(), "scale", qkv0.abs().max().item(),
              "out diff", (out0-out1).abs().max().item(), "out scale", out0.abs().max().item(), flush=True)
    # backward
    d_out = torch.randn(n_seq, D, device=dev)
    scratch = torch.empty(ops.scratch_bytes(shape), dtype=torch.uint8, device=dev)
    res = []
    for mode, saved in ((0, saved0), (1, saved1)):
        ops.news_encoder_fwd(shape, ids, table, params, saved, 0.2, 7, mode)
        dp = torch.empty_like(params); dr = torch.empty(n_seq*L, D, device=dev)
        ops.news_encoder_bwd(shape, ids, table, params, d_out, saved, scratch, dp, dr, 0.2, 7, mode)
        torch.cuda.synchronize()
        res.append((dp.clone(), dr.clone()))
    print(n_seq, "bwd d_params rel", ((res[0][0]-res[1][0]).norm()/res[0][0].norm()).item(),
          "d_rows rel", ((res[0][1]-res[1][1]).norm()/res[0][1].norm()).item(), flush=True)
# timing
n_seq = 3520
shape = ops.EncoderShape(n_seq, L, D, h, Q, V)
ids = torch.randint(0, V, (n_seq, L), device=dev)
for mode in (0, 1):
    saved = torch.empty(ops.saved_bytes(shape), dtype=torch.uint8, device=dev)
    for _ in range(3): ops.news_encoder_fwd(shape, ids, table, params, saved, 0.2, 7, mode)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): ops.news_encoder_fwd(shape, ids, table, params, saved, 0.2, 7, mode)
    torch.cuda.synchronize(); print("mode", mode, "fwd ms", (time.perf_counter()-t)*100)
