"""Generates tests/golden/bert_*.npz by running the UNMODIFIED reference `model/nrms.py` in this container.

    python tests/golden/make_golden_bert.py          # needs /root/reference (build container only)

Injected around (not into) the reference module so that it imports and is reproducible here:
  * `torchsnooper` and `tools` (absent / importing matplotlib) are answered with empty stand-ins — the
    module only imports them (nrms.py:5,10; every `@snoop()` is commented out);
  * for the train-mode cases `torch.nn.functional.dropout` hands out explicit multiplier tensors in the
    order the module draws them: candidate vectors (nrms.py:339 -> :254), history vectors (:343), attention
    probabilities (:349 -> :45-47), so the CUDA path and the oracle can be run with the very same masks.
The fixtures hold the inputs, the tiny model in full and sampled entries + norms of the real-size one.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/MIND_2020"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import nrms_bert_oracle as OB  # noqa: E402
from pytorch_news_recommender_b200 import synthetic as S  # noqa: E402
from make_golden import DropoutInjector, summarize  # noqa: E402


def load_reference():
    snoop = types.ModuleType("torchsnooper")
    snoop.snoop = lambda *a, **k: (lambda f: f)
    tools = types.ModuleType("tools")
    tools.log_exec_time = lambda f: f
    saved = {k: sys.modules.get(k) for k in ("torchsnooper", "tools")}
    sys.modules["torchsnooper"], sys.modules["tools"] = snoop, tools
    try:
        spec = importlib.util.spec_from_file_location("ref_nrms_bert", os.path.join(REF, "model", "nrms.py"))
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ref


class RefConfig:
    """The attributes nrms.Model reads (config.py:19,57,66-74)."""

    def __init__(self, cfg: OB.BertOracleConfig, data_path: str, npz: str):
        self.data_path = data_path
        self.bert_embedding_pretrained = npz
        self.bert_embed_size = cfg.bert_embed_size
        self.news_feature_size = cfg.news_feature_size
        self.user_heads_num = cfg.user_heads_num
        self.query_vector_dim_large = cfg.query_vector_dim_large
        self.dropout = cfg.dropout
        self.device = torch.device("cpu")


def make_mults(rng, B, S_, H, E, heads, p):
    def m(shape):
        return torch.from_numpy((rng.random(shape) >= p).astype(np.float32) / np.float32(1.0 - p))
    return {"cand": m((B, S_, E)), "hist": m((B, H, E)), "attn": m((B, heads, H, H))}


def run_case(ref, name, cfg: OB.BertOracleConfig, B, n_news, store_full, seed):
    out = {}
    H, C, E = cfg.history_len, cfg.sample_size + 1, cfg.bert_embed_size
    table = S.make_news_vector_table(n_news, E, seed=seed)
    pool = S.make_news_pool(n_news, 4, 10, seed=seed)
    batch = S.make_train_batch(pool, B, H, cfg.sample_size, seed=seed, short_tail=0.34, min_hist=1)
    tmp = tempfile.mkdtemp()
    S.save_embedding_npz(os.path.join(tmp, "bert.npz"), table)
    rc = RefConfig(cfg, tmp + "/", "bert.npz")

    torch.manual_seed(42)
    model = ref.Model(rc)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert list(sd0.keys()) == OB.state_dict_keys(), list(sd0.keys())
    sd_or = OB.init_state_dict(cfg, table, seed=42)
    for k in sd0:
        assert torch.equal(sd0[k], sd_or[k]), f"init mismatch {k}"
    out["meta/dims"] = np.array([B, H, C, E, cfg.user_heads_num, cfg.query_vector_dim_large, n_news], dtype=np.int64)
    out["meta/dropout"] = np.float64(cfg.dropout)
    out["meta/lr"] = np.float64(cfg.learning_rate)
    out["meta/seed"] = np.int64(seed)
    for k in ("browsed_ids", "candidate_ids", "browsed_mask", "candidate_mask"):
        out[f"in/{k}"] = batch[k].numpy().astype(np.int32 if "ids" in k else np.uint8)
    if store_full:
        for k, v in sd0.items():
            out[f"sd0/{k}"] = v.numpy()
    summarize("sd0sum", sd0, out)

    # ---- eval-mode forward -----------------------------------------------------------------
    model.eval()
    with torch.no_grad():
        logits = model(batch)
        cand_vec = model.news_encoder((batch["candidate_ids"], None))
        hist_vec = model.news_encoder((batch["browsed_ids"], None))
        user_vec = model.user_encoder(hist_vec, batch["browsed_mask"])
    out["eval/logits"] = logits.numpy()
    out["eval/user_vec"] = user_vec.numpy()
    out["eval/cand_vec"] = cand_vec.numpy()

    # ---- eval-mode loss + grads ---------------------------------------------------------------
    model.zero_grad()
    loss = torch.nn.CrossEntropyLoss()(model(batch), torch.zeros(B).long())
    loss.backward()
    out["evalgrad/loss"] = np.float32(loss.item())
    g = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    if store_full:
        for k, v in g.items():
            out[f"evalgrad/full/{k}"] = v.numpy()
    summarize("evalgrad", g, out)

    # ---- train mode with injected dropout multipliers: 2 Adam steps (train_eval.py:166-205) ----
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate)
    crit = torch.nn.CrossEntropyLoss()
    rng = np.random.default_rng(2000 + seed)
    for step in range(2):
        mults = make_mults(rng, B, C, H, E, cfg.user_heads_num, cfg.dropout)
        with DropoutInjector() as inj:
            inj.queue = [mults["cand"], mults["hist"], mults["attn"]]
            outputs = model(batch)
            assert not inj.queue
        model.zero_grad()
        loss = crit(outputs, torch.zeros(len(outputs)).long())
        loss.backward()
        out[f"train/step{step}/loss"] = np.float32(loss.item())
        out[f"train/step{step}/logits"] = outputs.detach().numpy()
        g = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        summarize(f"train/step{step}/grad", g, out)
        if store_full and step == 0:
            for k, v in g.items():
                out[f"train/step0/gradfull/{k}"] = v.numpy()
        opt.step()
        summarize(f"train/step{step}/param", dict(model.state_dict()), out)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    torch.set_num_threads(4)
    ref = load_reference()
    tiny = OB.BertOracleConfig(history_len=7, sample_size=2, bert_embed_size=48, news_feature_size=48,
                               user_heads_num=4, query_vector_dim_large=20, dropout=0.2)
    run_case(ref, "bert_tiny", tiny, B=4, n_news=40, store_full=True, seed=21)
    real = OB.BertOracleConfig(history_len=50, sample_size=4, bert_embed_size=512, news_feature_size=512,
                               user_heads_num=8, query_vector_dim_large=400, dropout=0.2)
    run_case(ref, "bert_mind", real, B=6, n_news=900, store_full=False, seed=23)


if __name__ == "__main__":
    main()
