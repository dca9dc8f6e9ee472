"""NRMS on pre-computed BERT news vectors with a masked user encoder — B200-native drop-in for the
reference's sibling plugin `model/nrms.py`.

Same plugin contract as `nrms_v0`: `Model(config)`, `forward(batch) -> [B, S]` logits on the device, the
reference's sub-module / parameter names (`state_dict()` keys are interchangeable) and its RNG draw order at
construction.  What the reference computes (nrms.py:297-366): news vector = row of a trainable BERT-vector
table -> Linear -> dropout (:216-256); user vector = multi-head self-attention over the history WITH the
padding mask, dropout on the attention probabilities and an output projection (:26-86), then additive
attention with the same mask (:88-117); click score = dot product, padded candidates -1e9 (:361-363).

Every arithmetic op is one of this library's sm_100a kernels behind the C-ABI (include/nrms_b200.h, last
section): the Linears on the tcgen05 image GEMMs, attention / pooling / dropout / gather / table gradient /
scoring as CUDA kernels.  torch carries memory and the autograd tape, nothing else; there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import engine, ops
from .._lib import NrmsError


# ------------------------------------------------------------------------------------------------------
# autograd nodes: one per C-ABI pair
# ------------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """nn.Linear over the last dim (nrms_linear_fwd / nrms_linear_bwd)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        ctx.save_for_backward(x2, weight, bias)
        ctx.in_shape = x.shape
        return ops.linear_fwd(x2, weight, bias).view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight, bias = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1]).contiguous()
        dx, dW, db = ops.linear_bwd(x2, weight, dy2, need_dx=ctx.needs_input_grad[0], need_dbias=bias is not None)
        return (dx.view(ctx.in_shape) if dx is not None else None), dW, db


class DropoutFn(torch.autograd.Function):
    """nn.Dropout with the library's Philox masks (nrms_dropout_apply); the backward is the same op."""

    @staticmethod
    def forward(ctx, x, seed, stream_id, p):
        ctx.args = (seed, stream_id, p)
        return ops.dropout_apply(x, seed, stream_id, p)

    @staticmethod
    def backward(ctx, dy):
        seed, stream_id, p = ctx.args
        return ops.dropout_apply(dy.contiguous(), seed, stream_id, p), None, None, None


class MaskedAttentionFn(torch.autograd.Function):
    """Attention.forward of every head (nrms.py:26-49) on the fused Q|K|V projection."""

    @staticmethod
    def forward(ctx, qkv, mask, heads, p, seed):
        qkv = qkv.contiguous()
        out, probs = ops.masked_attention_fwd(qkv, mask, heads, p, seed)
        ctx.save_for_backward(qkv, probs)
        ctx.mask, ctx.args = mask, (heads, p, seed)
        return out

    @staticmethod
    def backward(ctx, d_out):
        qkv, probs = ctx.saved_tensors
        heads, p, seed = ctx.args
        return ops.masked_attention_bwd(qkv, ctx.mask, probs, d_out.contiguous(), heads, p, seed), None, None, None, None


class MaskedPoolFn(torch.autograd.Function):
    """AdditiveAttention.forward after its Linear (nrms.py:107-117)."""

    @staticmethod
    def forward(ctx, t, query_vector, x, mask):
        t, x = t.contiguous(), x.contiguous()
        out, alpha = ops.masked_pool_fwd(t, query_vector, x, mask)
        ctx.save_for_backward(t, query_vector, x, alpha)
        ctx.mask = mask
        return out

    @staticmethod
    def backward(ctx, d_out):
        t, qv, x, alpha = ctx.saved_tensors
        d_t, d_x, d_qv = ops.masked_pool_bwd(t, qv, x, ctx.mask, alpha, d_out.contiguous())
        return d_t, d_qv, d_x, None


class NewsVectorFn(torch.autograd.Function):
    """nn.Embedding lookup of the BERT-vector table (nrms.py:249) and its dense gradient: gather kernel
    forward; backward = the library's deterministic de-duplicated row reduction (stable radix sort of
    (id, row), no atomics).  The table has no padding_idx (nrms.py:222-224), while the reduction kernels
    drop id 0 and handle rows of <= 384 floats: ids are shifted by one and a row is reduced as two halves."""

    @staticmethod
    def forward(ctx, ids, table):
        flat = ids.reshape(-1).contiguous()
        ctx.save_for_backward(flat)
        ctx.table_shape = table.shape
        return ops.gather_rows(table, flat).view(*ids.shape, table.shape[1])

    @staticmethod
    def backward(ctx, d_rows):
        (flat,) = ctx.saved_tensors
        V, E = ctx.table_shape
        d_rows = d_rows.reshape(-1, E).contiguous()
        parts = 1 if E <= 384 else 2
        if E % (4 * parts):
            raise NrmsError(f"news vector width {E} must be a multiple of {4 * parts}")
        n = flat.numel() * parts
        if parts == 1:
            ids = flat + 1
        else:
            ids = ((flat + 1) * parts).unsqueeze(1) + torch.arange(parts, device=flat.device)
            ids = ids.reshape(-1)
        vocab = (V + 1) * parts
        plan = torch.empty(ops.embedding_plan_bytes(n, vocab), dtype=torch.uint8, device=flat.device)
        ops.embedding_plan(ids, vocab, plan)
        d_table = torch.empty((V + 1, E), dtype=torch.float32, device=flat.device)
        ops.embedding_grad_dense(plan, d_rows, n, vocab, E // parts, d_table)
        return None, d_table[1:]


# ------------------------------------------------------------------------------------------------------
# modules (names and construction order = the reference's: state_dict keys, RNG stream)
# ------------------------------------------------------------------------------------------------------
class MultiHeadSelfAttention(nn.Module):
    """nrms.py:52-86: three input projections, masked scaled-dot-product attention with dropout on the
    probabilities, output projection."""

    def __init__(self, h, d_model, dropout):
        super().__init__()
        if d_model % h:
            raise ValueError("d_model must be divisible by the number of heads")
        self.d_k, self.h = d_model // h, h
        self.linear_layers = nn.ModuleList([nn.Linear(d_model, d_model) for _ in range(3)])
        self.output_linear = nn.Linear(d_model, d_model)
        self.dropout = nn.Dropout(p=dropout)

    def forward(self, query, key=None, value=None, mask=None, seed=0):
        if (key is not None and key is not query) or (value is not None and value is not query):
            raise NrmsError("self-attention only: query, key and value are the same tensor on this path (nrms.py:270)")
        if not query.is_cuda:
            raise NrmsError("runs on a CUDA device only (no CPU fallback)")
        # one [3E, E] GEMM for the three projections: their outputs sit side by side, as the attention kernel reads them
        W = torch.cat([l.weight for l in self.linear_layers], 0)
        b = torch.cat([l.bias for l in self.linear_layers], 0)
        qkv = LinearFn.apply(query, W, b)
        p = float(self.dropout.p) if self.training else 0.0
        ctx = MaskedAttentionFn.apply(qkv, mask, self.h, p, seed)
        return LinearFn.apply(ctx, self.output_linear.weight, self.output_linear.bias)


class AdditiveAttention(nn.Module):
    """nrms.py:88-117."""

    def __init__(self, query_vector_dim, input_vector_dim):
        super().__init__()
        self.linear = nn.Linear(input_vector_dim, query_vector_dim)
        self.query_vector = nn.Parameter(torch.empty(query_vector_dim).uniform_(-0.1, 0.1))

    def forward(self, input, mask=None):
        if not input.is_cuda:
            raise NrmsError("runs on a CUDA device only (no CPU fallback)")
        t = LinearFn.apply(input, self.linear.weight, self.linear.bias)
        return MaskedPoolFn.apply(t, self.query_vector, input, mask)


class BertNewsEncoder(nn.Module):
    """nrms.py:216-256: BERT-vector table -> Linear -> dropout."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        emb = np.load(config.data_path + config.bert_embedding_pretrained)["embeddings"].astype("float32")
        if emb.shape[1] != config.bert_embed_size:
            raise ValueError("news vector width %d != config.bert_embed_size %d" % (emb.shape[1], config.bert_embed_size))
        self.news_embedding = nn.Embedding.from_pretrained(torch.tensor(emb), freeze=False).to(config.device)
        self.news_dense = nn.Sequential(nn.Linear(config.bert_embed_size, config.bert_embed_size))
        self.dropout = nn.Dropout(p=config.dropout)

    def encode(self, news_ids):
        """Table lookup + Linear (no dropout): ids of any shape -> [..., bert_embed_size]."""
        table = self.news_embedding.weight
        if not table.is_cuda:
            raise NrmsError("BertNewsEncoder runs on a CUDA device only (no CPU fallback); move the model with .to('cuda')")
        ids = news_ids.to(table.device, dtype=torch.int64)
        x = NewsVectorFn.apply(ids, table)
        return LinearFn.apply(x, self.news_dense[0].weight, self.news_dense[0].bias)

    def drop(self, x, seed, stream_id):
        if self.training and self.dropout.p > 0:
            return DropoutFn.apply(x, seed, stream_id, float(self.dropout.p))
        return x

    def forward(self, _input, seed=0, stream_id=ops.DROP_CAND_VEC):
        """_input = (news_ids [B, n], anything) as in the reference (:246); -> [B, n, bert_embed_size]."""
        news_ids = _input[0] if isinstance(_input, (tuple, list)) else _input
        return self.drop(self.encode(news_ids), seed, stream_id)


class UserEncoder(nn.Module):
    """nrms.py:258-272."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.multi_head_self_attention = MultiHeadSelfAttention(config.user_heads_num, config.news_feature_size,
                                                                config.dropout)
        self.additive_attention = AdditiveAttention(config.query_vector_dim_large, config.news_feature_size)

    def forward(self, news_vectors, attn_masks, seed=0):
        if attn_masks is not None:
            attn_masks = attn_masks.to(news_vectors.device, dtype=torch.uint8).contiguous()
        a = self.multi_head_self_attention(news_vectors, news_vectors, news_vectors, mask=attn_masks, seed=seed)
        return self.additive_attention(a, attn_masks)


class ClickPredictor(nn.Module):
    """nrms.py:274-280 (kept for the plugin surface; Model.forward scores all slots with the masked scorer)."""

    def forward(self, news_vector, user_vector):
        return engine.ScoreFn.apply(news_vector.unsqueeze(1), user_vector, None).flatten()


class Model(nn.Module):
    """nrms.py:297-366."""

    def __init__(self, config):
        super().__init__()
        for k in ("bert_embed_size", "news_feature_size", "user_heads_num", "query_vector_dim_large"):
            if not hasattr(config, k):
                raise AttributeError("config.__nrms__() must be called before building the model (missing config.%s)" % k)
        if config.news_feature_size != config.bert_embed_size:
            # the reference constructs but cannot run such a model: Linear(news_feature_size, .) applied to
            # bert_embed_size columns raises at the first forward (nrms.py:262-270 vs :226-231)
            raise ValueError("config.news_feature_size (%d) must equal config.bert_embed_size (%d): the user encoder "
                             "consumes the news encoder's output" % (config.news_feature_size, config.bert_embed_size))
        self.news_encoder = BertNewsEncoder(config)
        self.user_encoder = UserEncoder(config)
        self.config = config
        self.device = config.device

    def forward(self, batch):
        """batch: browsed_ids [B, H], candidate_ids [B, S], browsed_mask [B, H], candidate_mask [B, S]
        (data_handler.py:236-250; browsed_titles / candidate_titles are read by the reference and ignored by
        its encoder, nrms.py:318-331, so they are optional here).  Returns logits [B, S] on the device."""
        table = self.news_encoder.news_embedding.weight
        if not table.is_cuda:
            raise NrmsError("Model.forward runs on a CUDA device only (no CPU fallback); use model.to('cuda')")
        dev = table.device
        seed = engine._next_seed(self) if self.training else 0
        cand_ids = batch["candidate_ids"].to(dev, dtype=torch.int64, non_blocking=True)
        hist_ids = batch["browsed_ids"].to(dev, dtype=torch.int64, non_blocking=True)
        (B, S), H = cand_ids.shape, hist_ids.shape[1]
        # one lookup + one Linear for all B * (S + H) news (the reference runs the encoder twice, :339,343)
        vec = self.news_encoder.encode(torch.cat([cand_ids.reshape(-1), hist_ids.reshape(-1)]))
        cand = self.news_encoder.drop(vec[:B * S].view(B, S, -1), seed, ops.DROP_CAND_VEC)
        hist = self.news_encoder.drop(vec[B * S:].view(B, H, -1), seed, ops.DROP_HIST_VEC)
        user = self.user_encoder(hist, batch["browsed_mask"], seed)
        mask = batch["candidate_mask"]
        if mask is not None:
            mask = mask.to(dev, dtype=torch.uint8, non_blocking=True).contiguous()
        return engine.ScoreFn.apply(cand, user, mask)
