import os, sys, tempfile, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as e; e.build()
from pytorch_news_recommender_b200 import synthetic as S, ops
from pytorch_news_recommender_b200.config import Config
from pytorch_news_recommender_b200.engine import FusedTrainer
from pytorch_news_recommender_b200.model import NRMS_V0
cfg = Config("X"); cfg.__nrms__()
cfg.n_words_title, cfg.history_len, cfg.sample_size, cfg.dropout = 30, 50, 4, 0.2
V, B = 70000, 64
tmp = tempfile.mkdtemp()
S.save_embedding_npz(os.path.join(tmp, "emb.npz"), S.make_embedding_table(V, 300, 0))
cfg.data_path, cfg.word_embedding_pretrained, cfg.device = tmp + "/", "emb.npz", torch.device("cuda:0")
pool = S.make_news_pool(65000, 30, V, seed=0)
batch = S.make_train_batch(pool, B, 50, 4, seed=0)
torch.manual_seed(42)
model = NRMS_V0(cfg).to(cfg.device); model.train()
tr = FusedTrainer(model)
t0 = model.state_dict()["news_encoder.word_embedding.0.weight"].clone()
for i in range(1):
    tr.step(batch)
t1 = model.state_dict()["news_encoder.word_embedding.0.weight"]
touched = torch.zeros(V, dtype=torch.bool)
touched[batch["browsed_titles"].reshape(-1)] = True
touched[batch["candidate_titles"].reshape(-1)] = True
touched[0] = False
moved = ((t1 - t0).abs().amax(dim=1) > 0).cpu()
g = tr.table_grad.cpu()
gn = (g.abs().amax(dim=1) > 0)
print("touched", touched.sum().item(), "moved", moved.sum().item(), "grad nonzero rows", gn.sum().item())
print("moved&~touched", (moved & ~touched).sum().item(), "touched&~moved", (touched & ~moved).sum().item())
print("gradnz&~touched", (gn & ~touched).sum().item(), "touched&~gradnz", (touched & ~gn).sum().item())
idx = (touched & ~gn).nonzero().flatten()[:10]
print("examples touched without grad:", idx.tolist())
ids = torch.cat([batch["candidate_titles"].reshape(-1, 30), batch["browsed_titles"].reshape(-1, 30)])
for v in idx[:3].tolist():
    where = (ids == v).nonzero()
    print(v, "occurs at title rows", where[:5].tolist(), "cand_mask of those?", )
ref = torch.zeros(V, 300).index_add_(0, ids.reshape(-1), tr._bufs[(64,5,50,30)]["d_rows"].cpu())
ref[0] = 0
print("max abs diff dense grad vs index_add of d_rows:", (ref - g).abs().max().item())
uniq = ops.embedding_plan_unique(tr.blobs.get("plan", 1, "cuda:0"), V)
print("unique", uniq.item())
