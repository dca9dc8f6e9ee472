"""Synthetic MIND-shaped data: the real MIND files are not shipped with the
reference (.gitignore:1-6) and there is no network, so every test and benchmark runs on
seeded synthetic inputs with the layout `MyDataset.__getitem__` produces
(data_handler.py:185-250): int64 ids, uint8 masks, front-aligned zero padding.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch


def make_embedding_table(vocab: int, dim: int = 300, seed: int = 0) -> np.ndarray:
    """default_rng(seed).standard_normal((V, dim)) float32, row 0 (padding_idx) zeroed."""
    t = np.random.default_rng(seed).standard_normal((vocab, dim)).astype(np.float32)
    t[0] = 0.0
    return t


def make_news_vector_table(n_news: int, dim: int = 512, seed: int = 0, scale: float = 0.5) -> np.ndarray:
    """[n_news + 1, dim] stand-in for the pre-computed BERT news vectors of the `nrms` variant
    (model/nrms.py:222-224, `news_embeds_512.npz`).  Row 0 — the id history padding uses — is an ordinary
    row: the reference builds this table without padding_idx."""
    return (np.random.default_rng(seed + 7).standard_normal((n_news + 1, dim)) * scale).astype(np.float32)


def save_embedding_npz(path: str, table: np.ndarray) -> None:
    """The npz the model constructor reads: key "embeddings" (nrms_v0.py:134-135)."""
    np.savez(path, embeddings=table)


@dataclass
class NewsPool:
    """`n_news` synthetic titles; row i is news id i+1 (0 = pad, data_handler.py:88,100)."""
    titles: np.ndarray  # [n_news, T] int64, front-aligned, zero padded

    @property
    def n_news(self) -> int:
        return self.titles.shape[0]

    def title_table(self) -> np.ndarray:
        """[n_news + 1, T]: row 0 is the all-zero pad title, so ids index it directly."""
        return np.concatenate([np.zeros((1, self.titles.shape[1]), np.int64), self.titles], 0)


def make_news_pool(n_news: int, n_words_title: int, vocab: int, seed: int = 0,
                   zipf: bool = False, min_len: int = 5) -> NewsPool:
    rng = np.random.default_rng(seed + 1)
    T = n_words_title
    lo = min(min_len, T)
    lens = rng.integers(lo, T + 1, size=n_news)
    if zipf:
        # Zipf(s=1.0) over ranks 1..V-1 via inverse CDF of the truncated harmonic series
        ranks = np.arange(1, vocab, dtype=np.float64)
        cdf = np.cumsum(1.0 / ranks)
        cdf /= cdf[-1]
        tok = np.searchsorted(cdf, rng.random((n_news, T))) + 1
    else:
        tok = rng.integers(1, vocab, size=(n_news, T))
    keep = np.arange(T)[None, :] < lens[:, None]
    return NewsPool((tok * keep).astype(np.int64))


def make_train_batch(pool: NewsPool, batch: int, history_len: int, n_neg: int, seed: int = 0,
                     short_tail: float = 0.05, min_hist: int = 5) -> Dict[str, torch.Tensor]:
    """One training batch in the MyDataset layout (type=0: S = sample_size + 1 slots,
    positive first: data_processor.py:526-527; histories >= 5 clicks: data_handler.py:90-93)."""
    rng = np.random.default_rng(seed + 2)
    H, C, T = history_len, n_neg + 1, pool.titles.shape[1]
    tt = pool.title_table()
    hist_len = rng.integers(min(min_hist, H), H + 1, size=batch)
    browsed_ids = rng.integers(1, pool.n_news + 1, size=(batch, H))
    browsed_mask = (np.arange(H)[None, :] < hist_len[:, None])
    browsed_ids = browsed_ids * browsed_mask
    cand_ids = rng.integers(1, pool.n_news + 1, size=(batch, C))
    n_real = np.full(batch, C)
    short = rng.random(batch) < short_tail
    if C > 1:
        n_real[short] = rng.integers(1, C, size=int(short.sum()))
    cand_mask = (np.arange(C)[None, :] < n_real[:, None])
    cand_ids = cand_ids * cand_mask
    return {
        "browsed_lens": torch.from_numpy(hist_len.astype(np.int64)),
        "browsed_ids": torch.from_numpy(browsed_ids.astype(np.int64)),
        "browsed_titles": torch.from_numpy(tt[browsed_ids].astype(np.int64)),
        "browsed_mask": torch.from_numpy(browsed_mask.astype(np.uint8)),
        "candidate_ids": torch.from_numpy(cand_ids.astype(np.int64)),
        "candidate_titles": torch.from_numpy(tt[cand_ids].astype(np.int64)),
        "candidate_mask": torch.from_numpy(cand_mask.astype(np.uint8)),
    }


def make_eval_impressions(pool: NewsPool, n_impr: int, history_len: int,
                          max_candidate_size: int = 300, seed: int = 0,
                          mean_candidates: float = 37.0):
    """Eval impressions (MyDataset type=1: S = max_candidate_size padded slots,
    data_handler.py:174-177): candidates per impression ~ clipped lognormal (mean ~37) in
    [2, max], labels Bernoulli with >= 1 positive and >= 1 negative.
    Returns (browsed_ids [N,H], candidate_ids [N,S], candidate_mask [N,S], y_true lists)."""
    rng = np.random.default_rng(seed + 3)
    H, S = history_len, max_candidate_size
    sigma = 0.8
    mu = np.log(mean_candidates) - 0.5 * sigma * sigma
    n_c = np.clip(np.rint(rng.lognormal(mu, sigma, size=n_impr)), 2, S).astype(np.int64)
    hist_len = rng.integers(1, H + 1, size=n_impr)
    browsed_ids = rng.integers(1, pool.n_news + 1, size=(n_impr, H)) * \
        (np.arange(H)[None, :] < hist_len[:, None])
    cmask = np.arange(S)[None, :] < n_c[:, None]
    cand_ids = rng.integers(1, pool.n_news + 1, size=(n_impr, S)) * cmask
    lab = (rng.random((n_impr, S)) < 0.1) & cmask
    # force one positive and one negative inside the real range
    pos_at = (rng.random(n_impr) * n_c).astype(np.int64)
    neg_at = (pos_at + 1 + (rng.random(n_impr) * (n_c - 1)).astype(np.int64)) % n_c
    lab[np.arange(n_impr), pos_at] = True
    lab[np.arange(n_impr), neg_at] = False
    y_true = [lab[i, :n_c[i]].astype(np.int64).tolist() for i in range(n_impr)] \
        if n_impr <= 200000 else None
    return {
        "browsed_ids": torch.from_numpy(browsed_ids.astype(np.int64)),
        "candidate_ids": torch.from_numpy(cand_ids.astype(np.int64)),
        "candidate_mask": torch.from_numpy(cmask.astype(np.uint8)),
        "labels": torch.from_numpy(lab.astype(np.uint8)),
        "n_candidates": torch.from_numpy(n_c),
        "y_true": y_true,
    }


def make_sample_lists(pool: NewsPool, n_samples: int, history_len: int, n_candidates_lo: int,
                      n_candidates_hi: int, seed: int = 0, min_hist: int = 5, n_categ: int = 18,
                      n_subcateg: int = 270) -> List[list]:
    """Sample lists in the `idx_*.pkl` layout (data_handler.py:100-103):
    [history_idx, categ_idx, subcateg_idx, imp_idx, imp_categ_idx, imp_subcateg_idx], news ids =
    table row + 1.  Candidate counts are uniform in [n_candidates_lo, n_candidates_hi]."""
    rng = np.random.default_rng(seed + 4)
    out = []
    for _ in range(n_samples):
        x = int(rng.integers(min(min_hist, history_len), history_len + 1))
        y = int(rng.integers(n_candidates_lo, n_candidates_hi + 1))
        hist = rng.integers(1, pool.n_news + 1, size=x).tolist()
        imp = rng.integers(1, pool.n_news + 1, size=y).tolist()
        out.append([hist, rng.integers(1, n_categ, size=x).tolist(), rng.integers(1, n_subcateg, size=x).tolist(),
                    imp, rng.integers(1, n_categ, size=y).tolist(), rng.integers(1, n_subcateg, size=y).tolist()])
    return out


def write_demo_files(path: str, config, n_news: int = 500, vocab: int = 2000, n_train: int = 256,
                     n_dev: int = 64, seed: int = 0, n_words_abst: Optional[int] = None) -> Dict[str, object]:
    """Writes every file `run_demo.py` resolves under `path` (run_demo.py:30-45,
    train_eval.py:157-159), with the literal names the reference uses:
    demo_word_embedding.npz, demo_news_title.pkl, demo_news_abst.pkl, idx_small_train.pkl,
    idx_small_dev.pkl, small_dev_behaviors.csv.  Returns the in-memory objects."""
    import os
    import pickle

    os.makedirs(path, exist_ok=True)
    rng = np.random.default_rng(seed + 5)
    table = make_embedding_table(vocab, config.word_embed_size, seed=seed)
    save_embedding_npz(os.path.join(path, 'demo_word_embedding.npz'), table)
    pool = make_news_pool(n_news, config.n_words_title, vocab, seed=seed)
    A = n_words_abst if n_words_abst is not None else getattr(config, 'n_words_abst', 40)
    absts = make_news_pool(n_news, A, vocab, seed=seed + 100).titles
    title_dict = {i: pool.titles[i].tolist() for i in range(n_news)}
    abst_dict = {i: absts[i].tolist() for i in range(n_news)}
    with open(os.path.join(path, 'demo_news_title.pkl'), 'wb') as f:
        pickle.dump(title_dict, f)
    with open(os.path.join(path, 'demo_news_abst.pkl'), 'wb') as f:
        pickle.dump(abst_dict, f)
    S = config.sample_size + 1
    train = make_sample_lists(pool, n_train, config.history_len, max(1, S - 2), S, seed=seed)   # <= S: see pack_samples
    dev = make_sample_lists(pool, n_dev, config.history_len, 2, min(40, config.max_candidate_size), seed=seed + 1,
                            min_hist=1)
    with open(os.path.join(path, 'idx_small_train.pkl'), 'wb') as f:
        pickle.dump(train, f)
    with open(os.path.join(path, 'idx_small_dev.pkl'), 'wb') as f:
        pickle.dump(dev, f)
    y_true = []
    for d in dev:
        n = len(d[3])
        lab = (rng.random(n) < 0.2).astype(np.int64)
        lab[int(rng.integers(0, n))] = 1
        lab[(int(np.argmax(lab)) + 1) % n] = 0
        y_true.append(lab.tolist())
    with open(os.path.join(path, 'small_dev_behaviors.csv'), 'w') as f:
        f.write('impression_id,y_true\n')
        for i, lab in enumerate(y_true):
            f.write('%d,%s\n' % (i, ' '.join(str(v) for v in lab)))
    return {'table': table, 'pool': pool, 'title_dict': title_dict, 'abst_dict': abst_dict,
            'train': train, 'dev': dev, 'y_true': y_true}
