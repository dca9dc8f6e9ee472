"""Generates tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py          # needs /root/reference (build container only)

It imports /root/reference/MIND_2020/model/nrms_v0.py and evaluation.py by path.  Two things
are injected around (not into) the reference code so it can run here and be reproduced:
  * `torch.device('cuda')` is answered with the CPU device (nrms_v0.py:248,250,272 hard-code
    CUDA; there is no GPU in the build container) — a proxy for the module-global `torch`;
  * for the train-mode cases `torch.nn.functional.dropout` pops explicit multiplier tensors
    from a queue (same call order as nrms_v0.py:137,171-173 inside the loops :255-260), so
    the masks are known and the CUDA path / oracle can be run with the very same masks.
The fixtures are small (ids, a tiny model in full, sampled entries + norms for the real-size
model) and are committed; the GPU box never needs /root/reference.
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

from oracle import nrms_oracle as O  # noqa: E402
from pytorch_news_recommender_b200 import synthetic as S  # noqa: E402

N_SAMPLES = 48


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_nrms_v0", os.path.join(REF, "model", "nrms_v0.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    class TorchProxy(types.ModuleType):
        def __init__(self, real):
            super().__init__("torch_proxy")
            self.__dict__["_real"] = real

        def __getattr__(self, k):
            return getattr(self.__dict__["_real"], k)

        def device(self, name, *a):
            real = self.__dict__["_real"]
            return real.device("cpu") if str(name).startswith("cuda") else real.device(name, *a)

    ref.torch = TorchProxy(torch)
    spec2 = importlib.util.spec_from_file_location("ref_evaluation", os.path.join(REF, "evaluation.py"))
    ev = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(ev)
    return ref, ev


class RefConfig:
    """The attributes nrms_v0.Model reads from config.py (Config + __nrms__)."""

    def __init__(self, cfg: O.OracleConfig, data_path: str, npz: str):
        self.data_path = data_path
        self.word_embedding_pretrained = npz
        self.word_embed_size = cfg.word_embed_size
        self.num_attention_heads = cfg.num_attention_heads
        self.query_vector_dim = cfg.query_vector_dim
        self.dropout = cfg.dropout
        self.device = torch.device("cpu")


class DropoutInjector:
    def __init__(self):
        self.queue = []
        self.real = torch.nn.functional.dropout

    def __call__(self, input, p=0.5, training=True, inplace=False):
        if training and self.queue:
            m = self.queue.pop(0)
            assert m.shape == input.shape, (m.shape, input.shape)
            return input * m
        return self.real(input, p, training, inplace)

    def __enter__(self):
        torch.nn.functional.dropout = self
        return self

    def __exit__(self, *a):
        torch.nn.functional.dropout = self.real


def make_masks(rng, n_titles, T, D, p):
    keep = rng.random((2, n_titles, T, D)) >= p
    m = keep.astype(np.float32) / np.float32(1.0 - p)
    return torch.from_numpy(m[0]), torch.from_numpy(m[1])


def queue_for(masks, B, C, H):
    cand_rows, hist_rows = O.flat_title_rows(B, C, H)
    q = []
    for c in range(C):
        q += [masks[0][cand_rows[:, c]], masks[1][cand_rows[:, c]]]
    for h in range(H):
        q += [masks[0][hist_rows[:, h]], masks[1][hist_rows[:, h]]]
    return q


def sample_idx(n):
    return np.unique(np.linspace(0, n - 1, N_SAMPLES).astype(np.int64))


def summarize(prefix, sd, out):
    for k, t in sd.items():
        a = t.detach().numpy().astype(np.float64).ravel()
        out[f"{prefix}/{k}/norm"] = np.float64(np.sqrt((a * a).sum()))
        out[f"{prefix}/{k}/sum"] = np.float64(a.sum())
        out[f"{prefix}/{k}/samples"] = t.detach().numpy().ravel()[sample_idx(a.size)].astype(np.float32)


def run_case(ref, name, cfg: O.OracleConfig, vocab, B, n_news, store_full, seed):
    out = {}
    T, H, C = cfg.n_words_title, cfg.history_len, cfg.sample_size + 1
    table = S.make_embedding_table(vocab, cfg.word_embed_size, seed=seed)
    pool = S.make_news_pool(n_news, T, vocab, seed=seed)
    batch = S.make_train_batch(pool, B, H, cfg.sample_size, seed=seed, short_tail=0.34)
    tmp = tempfile.mkdtemp()
    S.save_embedding_npz(os.path.join(tmp, "emb.npz"), table)
    rc = RefConfig(cfg, tmp + "/", "emb.npz")

    torch.manual_seed(42)                      # run_demo.py:22
    model = ref.Model(rc)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert list(sd0.keys()) == O.state_dict_keys(), list(sd0.keys())
    # oracle init must replay the same RNG stream
    sd_or = O.init_state_dict(cfg, table, seed=42)
    for k in sd0:
        assert torch.equal(sd0[k], sd_or[k]), f"init mismatch {k}"

    out["meta/vocab"] = np.int64(vocab)
    out["meta/dims"] = np.array([B, T, H, C, cfg.word_embed_size, cfg.num_attention_heads,
                                 cfg.query_vector_dim, n_news], dtype=np.int64)
    out["meta/dropout"] = np.float64(cfg.dropout)
    out["meta/lr"] = np.float64(cfg.learning_rate)
    out["meta/seed"] = np.int64(seed)
    out["in/browsed_titles"] = batch["browsed_titles"].numpy().astype(np.int32)
    out["in/candidate_titles"] = batch["candidate_titles"].numpy().astype(np.int32)
    out["in/candidate_mask"] = batch["candidate_mask"].numpy()
    if store_full:
        for k, v in sd0.items():
            out[f"sd0/{k}"] = v.numpy()
    summarize("sd0sum", sd0, out)

    # ---- eval-mode forward (train_eval.py:230,242) -------------------------------------
    model.eval()
    with torch.no_grad():
        logits = model(batch)
        cand_vec = torch.stack([model.news_encoder(x) for x in batch["candidate_titles"].permute(1, 0, 2)], 1)
        hist_vec = torch.stack([model.news_encoder(x) for x in batch["browsed_titles"].permute(1, 0, 2)], 1)
        user_vec = model.user_encoder(hist_vec)
    out["eval/logits"] = logits.numpy()
    out["eval/cand_vec"] = cand_vec.numpy()
    out["eval/user_vec"] = user_vec.numpy()
    out["eval/hist_vec_sample"] = hist_vec.numpy()[:, ::7]

    # ---- eval-mode (no dropout) loss + grads --------------------------------------------
    model.zero_grad()
    loss = torch.nn.CrossEntropyLoss()(model(batch), torch.zeros(B).long())
    loss.backward()
    out["evalgrad/loss"] = np.float32(loss.item())
    g = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    if store_full:
        for k, v in g.items():
            out[f"evalgrad/full/{k}"] = v.numpy()
    summarize("evalgrad", g, out)

    # ---- train mode with injected dropout masks: 2 Adam steps (train_eval.py:166-205) ---
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate)
    crit = torch.nn.CrossEntropyLoss()
    rng = np.random.default_rng(1000 + seed)
    n_titles = B * (C + H)
    for step in range(2):
        masks = make_masks(rng, n_titles, T, cfg.word_embed_size, cfg.dropout)
        out[f"train/step{step}/mask_seed"] = np.int64(1000 + seed)
        with DropoutInjector() as inj:
            inj.queue = queue_for(masks, B, C, H)
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
        summarize(f"train/step{step}/param", {k: v for k, v in model.state_dict().items()}, out)
    if store_full:
        for k, v in model.state_dict().items():
            out[f"train/final/{k}"] = v.detach().numpy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def run_metrics(ev):
    rng = np.random.default_rng(7)
    ys, ss, res = [], [], []
    cases = [([0, 1, 0, 0, 1], [.1, .9, .3, .2, .25]), ([1, 0, 0, 0, 0, 0], [.2, .5, .1, .3, .05, 0]),
             ([1, 0, 1, 0], [.5, .5, .5, .1])]
    for n in list(rng.integers(2, 16, size=40)) + list(rng.integers(17, 300, size=40)):
        y = (rng.random(n) < 0.15).astype(np.int64)
        y[rng.integers(0, n)] = 1
        y[(np.flatnonzero(y)[0] + 1) % n] = 0
        s = rng.standard_normal(n).astype(np.float32)
        if n <= 16 and rng.random() < 0.5:      # ties only where np.argsort is stable (n <= 16)
            s = np.round(s * 2) / 2
        cases.append((y.tolist(), s.tolist()))
    import warnings
    for y, s in cases:
        y = np.asarray(y)
        s32 = np.asarray(s, dtype=np.float32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = [ev.auc_score(y, s32), ev.mrr_score(y, s32), ev.ndcg_score(y, s32, 5), ev.ndcg_score(y, s32, 10)]
        ys.append(y)
        ss.append(s32)
        res.append(r)
    # single-class impressions (SURVEY §8c KAT 4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for y in ([0, 0, 0], [1, 1]):
            s32 = np.linspace(0, 1, len(y)).astype(np.float32)
            y = np.asarray(y)
            r = [ev.auc_score(y, s32), ev.mrr_score(y, s32), ev.ndcg_score(y, s32, 5), ev.ndcg_score(y, s32, 10)]
            ys.append(y)
            ss.append(s32)
            res.append([float(v) for v in r])
    offsets = np.cumsum([0] + [len(y) for y in ys]).astype(np.int64)
    path = os.path.join(HERE, "metrics.npz")
    np.savez_compressed(path, labels=np.concatenate(ys).astype(np.uint8), scores=np.concatenate(ss),
                        offsets=offsets, expected=np.asarray(res, dtype=np.float64))
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    torch.set_num_threads(4)
    ref, ev = load_reference()
    tiny = O.OracleConfig(n_words_title=6, history_len=5, sample_size=2, word_embed_size=24,
                          num_attention_heads=4, query_vector_dim=8, dropout=0.2, learning_rate=1e-3)
    run_case(ref, "tiny", tiny, vocab=50, B=3, n_news=40, store_full=True, seed=11)
    mind = O.OracleConfig(n_words_title=30, history_len=50, sample_size=4, word_embed_size=300,
                          num_attention_heads=10, query_vector_dim=200, dropout=0.2, learning_rate=1e-3)
    run_case(ref, "mind", mind, vocab=500, B=4, n_news=300, store_full=False, seed=5)
    long_ = O.OracleConfig(n_words_title=48, history_len=200, sample_size=8, word_embed_size=300,
                           num_attention_heads=10, query_vector_dim=200, dropout=0.2, learning_rate=1e-3)
    run_case(ref, "long", long_, vocab=400, B=2, n_news=500, store_full=False, seed=9)
    run_metrics(ev)


if __name__ == "__main__":
    main()
