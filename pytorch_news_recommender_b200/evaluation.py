"""Ranking metrics of the reference's `evaluation.py:6-27` — same function names and argument
meaning (`y_true`: 0/1 labels of one impression, `y_score`: its scores), computed by the
`rank_metrics` kernel on the GPU in float64 instead of numpy/sklearn on the host.

Semantics reproduced (SURVEY.md §8c KATs): DCG gains 2^y-1 with log2(rank+1) discounts over
`argsort(score)[::-1]` (ties resolve highest-index-first), nDCG = DCG/ideal DCG, MRR over ALL
positives, AUC = roc_auc_score (midrank ties); NaN where the reference yields NaN
(single-class impressions; sklearn raises a warning and returns NaN).

The per-impression functions exist for call-site compatibility; the batched entry points
(`impression_metrics`, `evaluate_scores`) are what the scoring path uses.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import ops
from ._lib import NrmsError

_COL = {"auc": 0, "mrr": 1, "ndcg5": 2, "ndcg10": 3}


def _device(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise NrmsError("ranking metrics run on a CUDA device only (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def impression_metrics(scores: torch.Tensor, labels: torch.Tensor, offsets: torch.Tensor,
                       max_len: int, row_stride=None) -> torch.Tensor:
    """[N, 4] float64 = AUC, MRR, nDCG@5, nDCG@10 per impression (device tensors in and out).
    Ragged `scores` [sum n_i] (row_stride None) or padded rows [N, row_stride]; `labels` ragged
    uint8; `offsets` int64 [N+1]."""
    return ops.rank_metrics(scores, labels, offsets, max_len, row_stride)


def _one(y_true, y_score, device=None) -> np.ndarray:
    dev = _device(device)
    y = torch.as_tensor(np.asarray(y_true).astype(np.uint8), device=dev)
    s = torch.as_tensor(np.asarray(y_score, dtype=np.float32), device=dev)
    if y.numel() != s.numel() or y.numel() == 0:
        raise ValueError("y_true and y_score must be non-empty and of equal length")
    off = torch.tensor([0, y.numel()], dtype=torch.int64, device=dev)
    return ops.rank_metrics(s, y, off, max_len=int(y.numel())).cpu().numpy()[0]


def ndcg_score(y_true, y_score, k=10):
    """evaluation.py:14-17.  The kernel serves the two cut-offs the path reports (5 and 10)."""
    if k not in (5, 10):
        raise ValueError("ndcg_score: k must be 5 or 10 (the cut-offs of the scoring path)")
    return float(_one(y_true, y_score)[_COL["ndcg5" if k == 5 else "ndcg10"]])


def dcg_score(y_true, y_score, k=10):
    """evaluation.py:6-11: DCG@k = nDCG@k x ideal DCG@k (the ideal DCG of 0/1 labels is closed
    form: sum of 1/log2(i+2) over the first min(k, #positives) ranks)."""
    y = np.asarray(y_true)
    n_pos = int(min(k, int((y > 0).sum())))
    ideal = float(np.sum(1.0 / np.log2(np.arange(n_pos) + 2)))
    return ndcg_score(y_true, y_score, k) * ideal if n_pos else 0.0


def mrr_score(y_true, y_score):
    """evaluation.py:20-24."""
    return float(_one(y_true, y_score)[_COL["mrr"]])


def auc_score(y_true, y_pred):
    """evaluation.py:26-27 (sklearn.metrics.roc_auc_score)."""
    return float(_one(y_true, y_pred)[_COL["auc"]])


def evaluate_scores(rank_score: torch.Tensor, y_true_lists: Sequence[Sequence[int]]) -> torch.Tensor:
    """train_eval.py:219-227 for all impressions at once: impression i owns
    rank_score[i][:len(y_true[i])] (positional pairing, padded slots dropped).  rank_score is
    the [N, S] device tensor of concatenated model outputs; returns [N, 4] float64 on device."""
    if not rank_score.is_cuda:
        raise NrmsError("evaluate_scores expects the scores on the CUDA device they were computed on")
    n, stride = rank_score.shape
    if len(y_true_lists) != n:
        raise ValueError(f"{len(y_true_lists)} label lists for {n} score rows")
    lens = np.fromiter((len(y) for y in y_true_lists), dtype=np.int64, count=n)
    if lens.max(initial=0) > stride:
        raise ValueError("an impression has more labels than scored candidate slots "
                         "(the reference fails on this too: scores are truncated at max_candidate_size)")
    off = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    flat = np.fromiter((v for y in y_true_lists for v in y), dtype=np.uint8, count=int(off[-1]))
    dev = rank_score.device
    return ops.rank_metrics(rank_score.contiguous(), torch.from_numpy(flat).to(dev), torch.from_numpy(off).to(dev),
                            max_len=stride, row_stride=stride)
