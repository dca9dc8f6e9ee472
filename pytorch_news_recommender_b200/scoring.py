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
        """[B, H] and [B, S] news ids (+ optional uint8 mask) -> logits [B, S] (padded = -1e9).

        Tensor-core GEMM modes: two C-ABI calls per batch — the user encoder gathers its rows from the
        cache by id (`nrms_user_encoder_fwd_gather`), and the scorer dots the cached candidate vectors
        with the user vector by id (`nrms_score_cached`); neither the [B, H, D] history tensor nor the
        [B, S, D] candidate tensor (S = 300 padded slots) is materialised.  gemm_mode 0 keeps the three
        explicit steps (gather, get_user_vector, score)."""
        from .engine import _blobs, encoder_param_list, pack_params, weights_version
        from .model.nrms_v0 import _gemm_mode
        if self.news_vecs is None or self._cache_version != weights_version(self.model):
            self.build_cache()          # first use, or the weights changed since the cache was built
        dev = self.device
        b_ids = browsed_ids.to(dev, dtype=torch.int64, non_blocking=True).contiguous()
        c_ids = candidate_ids.to(dev, dtype=torch.int64, non_blocking=True).contiguous()
        mask = None if candidate_mask is None else candidate_mask.to(dev, dtype=torch.uint8, non_blocking=True).contiguous()
        B, H = b_ids.shape
        S = c_ids.shape[1]
        D = self.news_vecs.shape[1]
        cfg = self.model.config
        gm = _gemm_mode(cfg)
        if gm >= 1:
            params = encoder_param_list(self.model.user_encoder)
            shape = ops.EncoderShape(B, H, D, cfg.num_attention_heads, params[8].numel(), self.news_vecs.shape[0])
            blob = _blobs.get("infer_saved", ops.saved_bytes(shape, gm), dev)
            user = ops.user_encoder_fwd_gather(shape, b_ids, self.news_vecs, pack_params(params), blob, gm)
            return ops.score_cached(self.news_vecs, c_ids, user, mask)
        hist = ops.gather_rows(self.news_vecs, b_ids.view(-1), base=0).view(B, H, D)
        cand = ops.gather_rows(self.news_vecs, c_ids.view(-1), base=0).view(B, S, D)
        was_training = self.model.training
        self.model.eval()
        user = self.model.get_user_vector(hist)
        self.model.train(was_training)
        return ops.score_fwd(cand, user, mask)

    @torch.no_grad()
    def evaluate(self, impressions: Dict[str, torch.Tensor], batch: int = 4096, group=None) -> Dict[str, float]:
        """Scores `impressions` (browsed_ids, candidate_ids, candidate_mask, labels [N, S] uint8,
        n_candidates [N]; host or device tensors) in batches and returns the mean AUC / MRR / nDCG@5 /
        nDCG@10 over the impressions whose metric is defined, plus counts.  Per batch: the two scoring
        calls above and one metrics launch that reads the padded score / label rows in place; the
        per-impression metrics land in one [N, 4] float64 buffer that is reduced once at the end, so
        nothing but the final 8 numbers leaves the GPU.  Host inputs are staged batch by batch on a copy
        stream, the next batch's H2D running underneath the current batch's kernels.  With a process
        group every rank evaluates its shard of impressions and the sums are all-reduced."""
        ex = GradientExchange(group)
        N = impressions["browsed_ids"].shape[0]
        lo, hi = shard_range(N, ex.rank, ex.world)
        dev = self.device
        keys = ("browsed_ids", "candidate_ids", "candidate_mask", "labels", "n_candidates")
        dtypes = dict(browsed_ids=torch.int64, candidate_ids=torch.int64, candidate_mask=torch.uint8,
                      labels=torch.uint8, n_candidates=torch.int64)
        metrics = torch.empty((max(hi - lo, 1), 4), dtype=torch.float64, device=dev)
        main = torch.cuda.current_stream(dev)
        copy = getattr(self, "_copy_stream", None)
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(device=dev)

        def stage(i):
            """batch [i, j) on the device: slices of resident tensors, or an async copy of host slices"""
            j = min(i + batch, hi)
            if all(impressions[k].is_cuda for k in keys):
                return {k: impressions[k][i:j] for k in keys}, None
            copy.wait_stream(main)       # the staging buffers of two batches back are free again
            with torch.cuda.stream(copy):
                out = {k: impressions[k][i:j].to(dev, dtype=dtypes[k], non_blocking=True) for k in keys}
                ev = torch.cuda.Event()
                ev.record(copy)
            return out, ev

        starts = list(range(lo, hi, batch))
        nxt = stage(starts[0]) if starts else None
        for n_b, i in enumerate(starts):
            cur, ev = nxt
            nxt = stage(starts[n_b + 1]) if n_b + 1 < len(starts) else None
            if ev is not None:
                main.wait_event(ev)
                for t in cur.values():
                    t.record_stream(main)
            logits = self.score(cur["browsed_ids"], cur["candidate_ids"], cur["candidate_mask"])
            ops.rank_metrics_rows(logits, cur["labels"].contiguous(), cur["n_candidates"].contiguous(),
                                  out=metrics[i - lo:i - lo + logits.shape[0]])
        m = metrics[:hi - lo]
        ok = ~torch.isnan(m)
        sums = torch.where(ok, m, torch.zeros_like(m)).sum(0)
        cnts = ok.sum(0).to(torch.float64)
        both = ex.sum_over_ranks(torch.cat([sums, cnts]))
        sums, cnts = both[:4].cpu().numpy(), both[4:].cpu().numpy()
        mean = sums / np.maximum(cnts, 1)
        return {"auc": float(mean[0]), "mrr": float(mean[1]), "ndcg5": float(mean[2]), "ndcg10": float(mean[3]),
                "n_impressions": int(hi - lo) if ex.world == 1 else N, "n_defined": int(cnts[0])}
