"""Cached-vector scoring (BASELINE cfg4): encode every news title ONCE, then score impressions
from `browsed_ids` / `candidate_ids` with gathers of the cached vectors.

The reference re-encodes all 300 + 50 titles of every dev batch (train_eval.py:238-251: 350
encoder calls per batch) although its model already exposes the three hooks this needs
(`get_news_vector / get_user_vector / get_prediction`, nrms_v0.py:278-312, never called).  The
arithmetic is identical — the news vector of a title does not depend on the impression it
appears in (eval mode: no dropout) — so scores match the uncached forward bit-for-bit up to
kernel launch grouping.

News id convention (data_handler.py:88,100): id = row in the news table + 1, id 0 = padded slot
whose title is all zeros.  Padded slots are NOT masked inside the encoders, so
cache row 0 holds the encoding of the all-zero title.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import ops
from ._lib import NrmsError
from .parallel import GradientExchange, shard_range


class CachedScorer:
    def __init__(self, model, title_table: torch.Tensor, chunk: int = 65536):
        """title_table: int64 [n_news + 1, T], row 0 = all-zero pad title (ids index it directly)."""
        self.model = model
        table = model.news_encoder.word_embedding[0].weight
        if not table.is_cuda:
            raise NrmsError("CachedScorer needs the model on a CUDA device (no CPU fallback)")
        self.device = table.device
        self.titles = title_table.to(self.device, dtype=torch.int64).contiguous()
        self.chunk = chunk
        self.news_vecs: Optional[torch.Tensor] = None
        self._cache_version = None      # engine.weights_version(model) the cache was built from

    @torch.no_grad()
    def build_cache(self, rank: int = 0, world: int = 1) -> torch.Tensor:
        """[n_news + 1, D] news vectors.  With world > 1 the encoding is sharded by news id and
        all-gathered once."""
        was_training = self.model.training
        self.model.eval()
        n = self.titles.shape[0]
        lo, hi = shard_range(n, rank, world)
        parts = [self.model.get_news_vector(self.titles[i:min(i + self.chunk, hi)])
                 for i in range(lo, hi, self.chunk)]
        mine = torch.cat(parts, 0) if parts else torch.empty((0, self.model.config.word_embed_size), device=self.device)
        if world > 1:
            import torch.distributed as dist
            sizes = [shard_range(n, r, world) for r in range(world)]
            bufs = [torch.empty((b - a, mine.shape[1]), dtype=mine.dtype, device=self.device) for a, b in sizes]
            dist.all_gather(bufs, mine)
            mine = torch.cat(bufs, 0)
        self.news_vecs = mine
        from .engine import weights_version
        self._cache_version = weights_version(self.model)
        self.model.train(was_training)
        return mine

    @torch.no_grad()
    def score(self, browsed_ids: torch.Tensor, candidate_ids: torch.Tensor,
              candidate_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B, H] and [B, S] news ids (+ optional uint8 mask) -> logits [B, S] (padded = -1e9)."""
        from .engine import weights_version
        if self.news_vecs is None or self._cache_version != weights_version(self.model):
            self.build_cache()          # first use, or the weights changed since the cache was built
        dev = self.device
        b_ids = browsed_ids.to(dev, dtype=torch.int64, non_blocking=True).contiguous()
        c_ids = candidate_ids.to(dev, dtype=torch.int64, non_blocking=True).contiguous()
        B, H = b_ids.shape
        S = c_ids.shape[1]
        D = self.news_vecs.shape[1]
        hist = ops.gather_rows(self.news_vecs, b_ids.view(-1), base=0).view(B, H, D)
        cand = ops.gather_rows(self.news_vecs, c_ids.view(-1), base=0).view(B, S, D)
        was_training = self.model.training
        self.model.eval()
        user = self.model.get_user_vector(hist)
        self.model.train(was_training)
        mask = None if candidate_mask is None else candidate_mask.to(dev, dtype=torch.uint8, non_blocking=True).contiguous()
        return ops.score_fwd(cand, user, mask)

    @torch.no_grad()
    def evaluate(self, impressions: Dict[str, torch.Tensor], batch: int = 4096, group=None) -> Dict[str, float]:
        """Scores `impressions` (browsed_ids, candidate_ids, candidate_mask, labels [N, S] uint8,
        n_candidates [N]) in batches and returns the mean AUC / MRR / nDCG@5 / nDCG@10 over the
        impressions whose metric is defined, plus counts.  Metrics never leave the GPU until the
        final 4 sums; with a process group every rank evaluates its shard of impressions and the
        sums are all-reduced."""
        ex = GradientExchange(group)
        N = impressions["browsed_ids"].shape[0]
        lo, hi = shard_range(N, ex.rank, ex.world)
        dev = self.device
        sums = torch.zeros(4, dtype=torch.float64, device=dev)
        cnts = torch.zeros(4, dtype=torch.float64, device=dev)
        for i in range(lo, hi, batch):
            j = min(i + batch, hi)
            logits = self.score(impressions["browsed_ids"][i:j], impressions["candidate_ids"][i:j],
                                impressions["candidate_mask"][i:j])
            n_c = impressions["n_candidates"][i:j].to(dev, dtype=torch.int64)
            S = logits.shape[1]
            off = torch.zeros(j - i + 1, dtype=torch.int64, device=dev)
            torch.cumsum(n_c, 0, out=off[1:])
            lab = impressions["labels"][i:j].to(dev)
            keep = torch.arange(S, device=dev)[None, :] < n_c[:, None]
            m = ops.rank_metrics(logits, lab[keep].contiguous(), off, max_len=S, row_stride=S)
            ok = ~torch.isnan(m)
            sums += torch.where(ok, m, torch.zeros_like(m)).sum(0)
            cnts += ok.sum(0)
        both = ex.sum_over_ranks(torch.cat([sums, cnts]))
        sums, cnts = both[:4].cpu().numpy(), both[4:].cpu().numpy()
        mean = sums / np.maximum(cnts, 1)
        return {"auc": float(mean[0]), "mrr": float(mean[1]), "ndcg5": float(mean[2]), "ndcg10": float(mean[3]),
                "n_impressions": int(hi - lo) if ex.world == 1 else N, "n_defined": int(cnts[0])}
