"""CPU oracle for the NRMS train + scoring hot path — TEST INFRASTRUCTURE ONLY.

This file restates, op by op, what the reference computes on the path named by
BASELINE.json:north_star.  It is NOT part of the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
it, and only as the checker / the timed CPU arm.  The product
(`pytorch_news_recommender_b200`) never imports it and has no CPU fallback.

Parity pin: the reference has no tests or golden vectors of its own, so this
oracle is pinned against the REFERENCE ITSELF, run in the build container by
`tests/golden/make_golden.py` (which imports /root/reference/MIND_2020/model/nrms_v0.py and
evaluation.py) — the resulting fixtures live in `tests/golden/*.npz` and
`tests/test_oracle_golden.py` checks this file against them.

All paths cited below are relative to the reference's `MIND_2020/` directory.
Arithmetic: float32 torch CPU ops (the same ATen ops the reference calls); ids int64.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]


@dataclass
class OracleConfig:
    """The config.py knobs the path reads (config.py:30-35,52,57,71,88)."""
    n_words_title: int = 30
    history_len: int = 50
    sample_size: int = 4           # negatives; C = sample_size + 1 in training batches
    word_embed_size: int = 300
    num_attention_heads: int = 10
    query_vector_dim: int = 200
    dropout: float = 0.2
    learning_rate: float = 1e-3


NEWS = "news_encoder."
USER = "user_encoder."
TABLE_KEY = "news_encoder.word_embedding.0.weight"


def state_dict_keys() -> List[str]:
    """The 19 tensors of nrms_v0.Model.state_dict() in registration order."""
    keys = [TABLE_KEY]
    for enc in (NEWS, USER):
        for lin in ("W_Q", "W_K", "W_V"):
            keys += [f"{enc}multihead_self_attention.{lin}.weight",
                     f"{enc}multihead_self_attention.{lin}.bias"]
        keys += [f"{enc}additive_attention.attention_query_vector",
                 f"{enc}additive_attention.linear.weight",
                 f"{enc}additive_attention.linear.bias"]
    return keys


def init_state_dict(cfg: OracleConfig, table: np.ndarray, seed: Optional[int] = 42) -> StateDict:
    """Replays the RNG draw order of Model.__init__ (nrms_v0.py:223-228):
    per encoder — Linear W_Q, W_K, W_V (default init: weight then bias, nrms_v0.py:35-37),
    xavier_uniform_(gain=1) over the three weights (nrms_v0.py:41-44), additive Linear
    (nrms_v0.py:91), query vector U(-0.1, 0.1) (nrms_v0.py:92-93)."""
    if seed is not None:
        torch.manual_seed(seed)
    D, Q = cfg.word_embed_size, cfg.query_vector_dim
    sd: StateDict = {TABLE_KEY: torch.tensor(np.asarray(table, dtype=np.float32))}
    for enc in (NEWS, USER):
        lins = [torch.nn.Linear(D, D) for _ in range(3)]
        for m in lins:
            torch.nn.init.xavier_uniform_(m.weight, gain=1)
        add = torch.nn.Linear(D, Q)
        qv = torch.empty(Q).uniform_(-0.1, 0.1)
        for name, m in zip(("W_Q", "W_K", "W_V"), lins):
            sd[f"{enc}multihead_self_attention.{name}.weight"] = m.weight.detach().clone()
            sd[f"{enc}multihead_self_attention.{name}.bias"] = m.bias.detach().clone()
        sd[f"{enc}additive_attention.attention_query_vector"] = qv
        sd[f"{enc}additive_attention.linear.weight"] = add.weight.detach().clone()
        sd[f"{enc}additive_attention.linear.bias"] = add.bias.detach().clone()
    return {k: sd[k] for k in state_dict_keys()}


# ------------------------------------------------------------------------------------------
# model
# ------------------------------------------------------------------------------------------
def scaled_dot_product_attention(Q, K, V, d_k: int):
    """nrms_v0.py:13-23 — softmax(Q K^T / sqrt(d_k)) V; attn_mask is ignored there."""
    scores = torch.matmul(Q, K.transpose(-1, -2)) / np.sqrt(d_k)
    attn = F.softmax(scores, dim=-1)
    return torch.matmul(attn, V)


def multihead_self_attention(x, sd: StateDict, enc: str, n_heads: int):
    """nrms_v0.py:46-76 with K=V=Q=x, length=None (no mask), no output projection."""
    B = x.size(0)
    d_model = x.size(-1)
    d_k = d_model // n_heads
    p = enc + "multihead_self_attention."
    q = F.linear(x, sd[p + "W_Q.weight"], sd[p + "W_Q.bias"]).view(B, -1, n_heads, d_k).transpose(1, 2)
    k = F.linear(x, sd[p + "W_K.weight"], sd[p + "W_K.bias"]).view(B, -1, n_heads, d_k).transpose(1, 2)
    v = F.linear(x, sd[p + "W_V.weight"], sd[p + "W_V.bias"]).view(B, -1, n_heads, d_k).transpose(1, 2)
    ctx = scaled_dot_product_attention(q, k, v, d_k)
    return ctx.transpose(1, 2).contiguous().view(B, -1, n_heads * d_k)


def additive_attention(x, sd: StateDict, enc: str):
    """nrms_v0.py:100-126."""
    p = enc + "additive_attention."
    temp = torch.tanh(F.linear(x, sd[p + "linear.weight"], sd[p + "linear.bias"]))
    w = F.softmax(torch.matmul(temp, sd[p + "attention_query_vector"]), dim=1)
    return torch.bmm(w.unsqueeze(1), x).squeeze(1)


def news_encoder(ids, sd: StateDict, cfg: OracleConfig, training: bool,
                 masks: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """nrms_v0.py:154-176.  ids [n, T] int64 -> [n, D].
    masks = (m_emb, m_ctx), each [n, T, D] holding 0 or 1/(1-p): explicit dropout multipliers
    (used to compare against a CUDA run with the same Philox masks); None -> torch's own
    dropout RNG when training (what the reference does)."""
    table = sd[TABLE_KEY]
    x = F.embedding(ids, table, padding_idx=0)          # nrms_v0.py:134-137,166
    if masks is not None:
        x = x * masks[0]
    else:
        x = F.dropout(x, p=cfg.dropout, training=training)
    ctx = multihead_self_attention(x, sd, NEWS, cfg.num_attention_heads)   # :170
    if masks is not None:
        ctx = ctx * masks[1]
    else:
        ctx = F.dropout(ctx, p=cfg.dropout, training=training)            # :171-173
    return additive_attention(ctx, sd, NEWS)                               # :175


def user_encoder(x, sd: StateDict, cfg: OracleConfig):
    """nrms_v0.py:188-199 — no dropout, no history mask."""
    return additive_attention(multihead_self_attention(x, sd, USER, cfg.num_attention_heads), sd, USER)


def click_predictor(cand, user):
    """nrms_v0.py:205-216."""
    return torch.bmm(cand, user.unsqueeze(-1)).squeeze(-1)


def flat_title_rows(B: int, C: int, H: int):
    """Row numbering shared with the CUDA path: candidate title (b,c) is row b*C+c,
    clicked title (b,h) is row B*C + b*H + h."""
    cand = torch.arange(B * C).view(B, C)
    hist = B * C + torch.arange(B * H).view(B, H)
    return cand, hist


def model_forward(sd: StateDict, batch: Dict[str, torch.Tensor], cfg: OracleConfig,
                  training: bool = False,
                  masks: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                  per_slot: bool = True, return_parts: bool = False):
    """Model.forward nrms_v0.py:230-276.  per_slot=True follows the reference's loops over
    the C candidate slots and H history slots (nrms_v0.py:255-260); per_slot=False encodes
    all titles in one flattened call (bit-identical in eval mode)."""
    clicked = batch["browsed_titles"].long()       # [B, H, T]
    cands = batch["candidate_titles"].long()       # [B, C, T]
    B, H, T = clicked.shape
    C = cands.shape[1]
    cand_rows, hist_rows = flat_title_rows(B, C, H)

    def enc(ids, rows):
        m = None
        if masks is not None:
            m = (masks[0][rows], masks[1][rows])
        return news_encoder(ids, sd, cfg, training, m)

    if per_slot:
        cand_vec = torch.stack([enc(cands[:, c], cand_rows[:, c]) for c in range(C)], dim=1)
        hist_vec = torch.stack([enc(clicked[:, h], hist_rows[:, h]) for h in range(H)], dim=1)
    else:
        cand_vec = enc(cands.reshape(B * C, T), cand_rows.reshape(-1)).view(B, C, -1)
        hist_vec = enc(clicked.reshape(B * H, T), hist_rows.reshape(-1)).view(B, H, -1)
    user_vec = user_encoder(hist_vec, sd, cfg)                       # :266
    logits = click_predictor(cand_vec, user_vec)                     # :269
    mask = batch.get("candidate_mask")
    if mask is not None:
        logits = logits.masked_fill(mask == 0, -1e9)                 # :272-274
    if return_parts:
        return logits, cand_vec, hist_vec, user_vec
    return logits


def cross_entropy_vs_zero(logits):
    """train_eval.py:181,194-195 — nn.CrossEntropyLoss()(outputs, zeros(B).long())."""
    y = torch.zeros(len(logits), dtype=torch.long, device=logits.device)
    return F.cross_entropy(logits, y)


# ------------------------------------------------------------------------------------------
# optimizer — torch.optim.Adam defaults (train_eval.py:167,205), restated
# ------------------------------------------------------------------------------------------
@dataclass
class AdamState:
    step: int
    m: StateDict
    v: StateDict


def adam_init(sd: StateDict) -> AdamState:
    return AdamState(0, {k: torch.zeros_like(t) for k, t in sd.items()},
                     {k: torch.zeros_like(t) for k, t in sd.items()})


def adam_step(sd: StateDict, grads: StateDict, st: AdamState, lr: float, beta1: float = 0.9,
              beta2: float = 0.999, eps: float = 1e-8) -> None:
    """In-place Adam as torch/optim/adam.py::_single_tensor_adam (no weight decay, no amsgrad)."""
    st.step += 1
    bc1 = 1.0 - beta1 ** st.step
    bc2 = 1.0 - beta2 ** st.step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for k, p in sd.items():
        g = grads[k]
        st.m[k].lerp_(g, 1.0 - beta1)
        st.v[k].mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        denom = (st.v[k].sqrt() / bc2_sqrt).add_(eps)
        p.addcdiv_(st.m[k], denom, value=-step_size)


def loss_and_grads(sd: StateDict, batch, cfg: OracleConfig, training: bool,
                   masks=None, per_slot: bool = True):
    """forward + CE + loss.backward() (train_eval.py:189-204) on leaf copies of sd."""
    leaves = {k: t.detach().clone().requires_grad_(True) for k, t in sd.items()}
    logits = model_forward(leaves, batch, cfg, training=training, masks=masks, per_slot=per_slot)
    loss = cross_entropy_vs_zero(logits)
    loss.backward()
    grads = {k: (t.grad if t.grad is not None else torch.zeros_like(t)) for k, t in leaves.items()}
    return loss.detach(), logits.detach(), grads


def train_step(sd: StateDict, st: AdamState, batch, cfg: OracleConfig, training: bool = True,
               masks=None, per_slot: bool = True) -> float:
    """One iteration of the hot loop train_eval.py:187-205."""
    loss, _, grads = loss_and_grads(sd, batch, cfg, training, masks, per_slot)
    adam_step(sd, grads, st, cfg.learning_rate)
    return float(loss)


# ------------------------------------------------------------------------------------------
# metrics — evaluation.py:6-27, numpy restatement
# ------------------------------------------------------------------------------------------
def dcg_score(y_true, y_score, k=10):
    """evaluation.py:6-11 (ties: stable ascending argsort reversed — what np.argsort gives for
    n <= 16 and for tie-free scores)."""
    order = np.argsort(y_score, kind="stable")[::-1]
    y_true = np.take(y_true, order[:k])
    gains = 2 ** y_true - 1
    discounts = np.log2(np.arange(len(y_true)) + 2)
    return np.sum(gains / discounts)


def ndcg_score(y_true, y_score, k=10):
    """evaluation.py:14-17."""
    best = dcg_score(y_true, y_true, k)
    actual = dcg_score(y_true, y_score, k)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.float64(actual) / np.float64(best)


def mrr_score(y_true, y_score):
    """evaluation.py:20-24."""
    order = np.argsort(y_score, kind="stable")[::-1]
    y_true = np.take(y_true, order)
    rr_score = y_true / (np.arange(len(y_true)) + 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.float64(np.sum(rr_score)) / np.float64(np.sum(y_true))


def auc_score(y_true, y_score):
    """evaluation.py:26-27 -> sklearn.metrics.roc_auc_score (scikit-learn 1.9.0 in the build
    image; the reference pins no version).  Restated as the Mann-Whitney statistic with
    midrank ties, which is what the trapezoidal ROC area equals; NaN for single-class input
    (sklearn 1.9 warns and returns NaN)."""
    y = np.asarray(y_true)
    s = np.asarray(y_score, dtype=np.float64)
    pos, neg = s[y == 1], s[y != 1]
    if len(pos) == 0 or len(neg) == 0:
        return float("nan")
    gt = (pos[:, None] > neg[None, :]).sum()
    eq = (pos[:, None] == neg[None, :]).sum()
    return (gt + 0.5 * eq) / (len(pos) * len(neg))


def impression_metrics(y_true: Sequence[int], y_score: Sequence[float]) -> np.ndarray:
    """[AUC, MRR, nDCG@5, nDCG@10] of one impression (train_eval.py:219-227 computes AUC only;
    the other three are the imported-but-commented calls at train_eval.py:263-270)."""
    y = np.asarray(y_true)
    s = np.asarray(y_score, dtype=np.float32)
    return np.array([auc_score(y, s), mrr_score(y, s), ndcg_score(y, s, 5), ndcg_score(y, s, 10)],
                    dtype=np.float64)


def evaluate_scores(rank_score: np.ndarray, y_true_lists: Sequence[Sequence[int]]) -> np.ndarray:
    """train_eval.py:255-271: per-impression metrics on rank_score[i][:len(y_true[i])], then
    the mean over impressions (NaN propagates exactly as np.mean does there)."""
    per = np.stack([impression_metrics(y, rank_score[i][:len(y)]) for i, y in enumerate(y_true_lists)])
    return per.mean(axis=0), per
