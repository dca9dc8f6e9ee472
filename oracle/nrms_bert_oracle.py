"""CPU oracle for the `nrms` sibling variant — TEST INFRASTRUCTURE ONLY.

The reference's `model/nrms.py:297-366` Model: news vectors come from a trainable table of pre-computed
BERT vectors followed by one Linear (`:216-256`), the user encoder is a multi-head self-attention WITH a
key/query padding mask, dropout on the attention PROBABILITIES and an output projection (`:26-86`), then an
additive attention with a padding mask (`:88-117,258-271`); the click score is the dot product with
`candidate_mask` filled by -1e9 (`:361-363`).

Same rules as `oracle/nrms_oracle.py`: only `tests/`, `smoke()` and the CPU arms of the benches import this
file; the product never does.  Pinned against the reference module itself by
`tests/golden/make_golden_bert.py` -> `tests/golden/bert_*.npz` (`tests/test_oracle_golden.py`).
All paths cited are relative to the reference's `MIND_2020/`.  float32 torch CPU ops; ids int64.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]

TABLE_KEY = "news_encoder.news_embedding.weight"
DENSE = "news_encoder.news_dense.0."
MHSA = "user_encoder.multi_head_self_attention."
ADD = "user_encoder.additive_attention."


@dataclass
class BertOracleConfig:
    """config.py:30-35,57 and Config.__nrms__ (config.py:66-74).  The shipped defaults
    (news_feature_size 800 vs bert_embed_size 512) do not compose — `UserEncoder` (nrms.py:262-267) is
    built on news_feature_size while `BertNewsEncoder` emits bert_embed_size columns — so a runnable
    configuration sets them equal; both knobs are kept."""
    history_len: int = 50
    sample_size: int = 4
    bert_embed_size: int = 512
    news_feature_size: int = 512
    user_heads_num: int = 8
    query_vector_dim_large: int = 400
    dropout: float = 0.2
    learning_rate: float = 1e-3


def state_dict_keys() -> List[str]:
    """nrms.Model.state_dict() in registration order (a module's own parameters come before its
    children's: `query_vector` precedes `linear.*`)."""
    keys = [TABLE_KEY, DENSE + "weight", DENSE + "bias"]
    for i in range(3):
        keys += [f"{MHSA}linear_layers.{i}.weight", f"{MHSA}linear_layers.{i}.bias"]
    keys += [MHSA + "output_linear.weight", MHSA + "output_linear.bias",
             ADD + "query_vector", ADD + "linear.weight", ADD + "linear.bias"]
    return keys


def init_state_dict(cfg: BertOracleConfig, table: np.ndarray, seed: Optional[int] = 42) -> StateDict:
    """Replays the RNG draws of Model.__init__ (nrms.py:301-302): news_dense Linear (:226-230); the three
    projection Linears then output_linear (:65-66); the additive Linear (:91) and the query vector
    U(-0.1, 0.1) (:96)."""
    if seed is not None:
        torch.manual_seed(seed)
    Eb, E, Q = cfg.bert_embed_size, cfg.news_feature_size, cfg.query_vector_dim_large
    dense = torch.nn.Linear(Eb, Eb)
    lins = [torch.nn.Linear(E, E) for _ in range(3)]
    out = torch.nn.Linear(E, E)
    add = torch.nn.Linear(E, Q)
    qv = torch.empty(Q).uniform_(-0.1, 0.1)
    sd: StateDict = {TABLE_KEY: torch.tensor(np.asarray(table, dtype=np.float32)),
                     DENSE + "weight": dense.weight.detach().clone(), DENSE + "bias": dense.bias.detach().clone()}
    for i, m in enumerate(lins):
        sd[f"{MHSA}linear_layers.{i}.weight"] = m.weight.detach().clone()
        sd[f"{MHSA}linear_layers.{i}.bias"] = m.bias.detach().clone()
    sd[MHSA + "output_linear.weight"] = out.weight.detach().clone()
    sd[MHSA + "output_linear.bias"] = out.bias.detach().clone()
    sd[ADD + "query_vector"] = qv
    sd[ADD + "linear.weight"] = add.weight.detach().clone()
    sd[ADD + "linear.bias"] = add.bias.detach().clone()
    return {k: sd[k] for k in state_dict_keys()}


def news_encoder(ids, sd: StateDict, mult: Optional[torch.Tensor] = None):
    """nrms.py:235-256 — table lookup (no padding_idx: row 0 is an ordinary trainable row), Linear,
    dropout.  `mult` = explicit dropout multipliers (0 or 1/(1-p)) of the output's shape, None = eval."""
    x = F.linear(sd[TABLE_KEY][ids], sd[DENSE + "weight"], sd[DENSE + "bias"])
    return x if mult is None else x * mult


def masked_self_attention(x, mask, sd: StateDict, n_heads: int, attn_mult: Optional[torch.Tensor] = None):
    """nrms.py:26-86.  x [B, L, E], mask [B, L] (1 = real slot).  The score of (query i, key j) is
    replaced by -1e9 unless BOTH slots are real (:38-41), so a padded query row attends uniformly over all
    L keys; dropout multiplies the probabilities (:45-47); an output projection follows (:86)."""
    B, L, E = x.shape
    dk = E // n_heads
    q, k, v = [F.linear(x, sd[f"{MHSA}linear_layers.{i}.weight"], sd[f"{MHSA}linear_layers.{i}.bias"])
               .view(B, L, n_heads, dk).transpose(1, 2) for i in range(3)]
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    if mask is not None:
        m = mask.to(torch.float32)
        both = (m.unsqueeze(1) * m.unsqueeze(2)).unsqueeze(1)          # [B, 1, L, L]
        scores = scores.masked_fill(both.expand(B, n_heads, L, L) == 0, -1e9)
    p = F.softmax(scores, dim=-1)
    if attn_mult is not None:
        p = p * attn_mult
    ctx = torch.matmul(p, v).transpose(1, 2).contiguous().view(B, L, E)
    return F.linear(ctx, sd[MHSA + "output_linear.weight"], sd[MHSA + "output_linear.bias"])


def masked_additive_attention(x, mask, sd: StateDict):
    """nrms.py:98-117 — softmax over the slots of tanh(x W^T + b) . q with padded slots at -1e9."""
    t = torch.tanh(F.linear(x, sd[ADD + "linear.weight"], sd[ADD + "linear.bias"]))
    s = torch.matmul(t, sd[ADD + "query_vector"])
    if mask is not None:
        s = s.masked_fill(mask == 0, -1e9)
    w = F.softmax(s, dim=1)
    return torch.bmm(w.unsqueeze(1), x).squeeze(1)


def user_encoder(hist_vec, browsed_mask, sd: StateDict, cfg: BertOracleConfig, attn_mult=None):
    """nrms.py:269-272."""
    a = masked_self_attention(hist_vec, browsed_mask, sd, cfg.user_heads_num, attn_mult)
    return masked_additive_attention(a, browsed_mask, sd)


def model_forward(sd: StateDict, batch, cfg: BertOracleConfig, mults: Optional[dict] = None, return_parts=False):
    """nrms.py:317-365.  batch: browsed_ids [B,H], candidate_ids [B,S], browsed_mask [B,H], candidate_mask
    [B,S].  mults (train mode) = {"cand": [B,S,E], "hist": [B,H,E], "attn": [B,h,H,H]} in the order the
    reference draws them (candidates first, :339; history, :343; probabilities, :349)."""
    mults = mults or {}
    cand = news_encoder(batch["candidate_ids"], sd, mults.get("cand"))
    hist = news_encoder(batch["browsed_ids"], sd, mults.get("hist"))
    user = user_encoder(hist, batch["browsed_mask"], sd, cfg, mults.get("attn"))
    pred = torch.sum(user.unsqueeze(1) * cand, 2)
    pred = pred.masked_fill(batch["candidate_mask"] == 0, -1e9)
    if return_parts:
        return pred, {"cand_vec": cand, "hist_vec": hist, "user_vec": user}
    return pred


def loss_and_grads(sd: StateDict, batch, cfg: BertOracleConfig, mults: Optional[dict] = None):
    """CrossEntropyLoss against label 0 (train_eval.py:189-199) and every parameter's gradient."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    logits = model_forward(leaf, batch, cfg, mults)
    loss = F.cross_entropy(logits, torch.zeros(logits.size(0), dtype=torch.long))
    loss.backward()
    return loss.detach(), logits.detach(), {k: v.grad for k, v in leaf.items()}
