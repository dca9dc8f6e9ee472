"""NRMS (title-only, "v0") — B200-native drop-in for the reference's `model/nrms_v0.py`.

Same plugin contract: `Model(config)`, `forward(batch) -> [B, S]` logits on
the device, `get_news_vector / get_user_vector / get_prediction`, the same sub-module and
parameter names (so `state_dict()` keys and reference checkpoints are interchangeable) and the
same RNG draw order at construction (so `torch.manual_seed(42)` gives the same initial
weights, reference nrms_v0.py:223-228).  None of the reference's ATen op sequence is used:
each encoder call is a handful of our sm_100a kernels behind the C-ABI (include/nrms_b200.h).

The sub-modules below only HOLD parameters; the arithmetic of a whole encoder is fused, so
`MultiHeadSelfAttention` / `AdditiveAttention` cannot be called on their own.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import engine
from .._lib import NrmsError


class MultiHeadSelfAttention(nn.Module):
    """Parameter holder for W_Q / W_K / W_V (reference nrms_v0.py:27-44: three
    Linear(d_model, d_model) with bias, xavier-uniform weights, no output projection)."""

    def __init__(self, d_model, num_attention_heads):
        super().__init__()
        if d_model % num_attention_heads:
            raise ValueError("d_model must be divisible by num_attention_heads")
        self.d_model, self.num_attention_heads = d_model, num_attention_heads
        self.d_k = self.d_v = d_model // num_attention_heads
        self.W_Q = nn.Linear(d_model, d_model)
        self.W_K = nn.Linear(d_model, d_model)
        self.W_V = nn.Linear(d_model, d_model)
        for lin in (self.W_Q, self.W_K, self.W_V):
            nn.init.xavier_uniform_(lin.weight, gain=1)

    def forward(self, *a, **k):
        raise NrmsError("fused into NewsEncoder/UserEncoder; call the encoder instead")


class AdditiveAttention(nn.Module):
    """Parameter holder for the pooling head (reference nrms_v0.py:84-93)."""

    def __init__(self, query_vector_dim, candidate_vector_dim):
        super().__init__()
        self.linear = nn.Linear(candidate_vector_dim, query_vector_dim)
        self.attention_query_vector = nn.Parameter(torch.empty(query_vector_dim).uniform_(-0.1, 0.1))

    def forward(self, *a, **k):
        raise NrmsError("fused into NewsEncoder/UserEncoder; call the encoder instead")


def _gemm_mode(config) -> int:
    """config.gemm_mode when set; otherwise the tcgen05 path (1) whenever its tile shapes cover
    the model dims (include/nrms_b200.h), else the exact-fp32 CUDA-core GEMMs (0).  Both are
    implementations of the same contractions on the GPU — neither is a fallback off the device."""
    gm = getattr(config, "gemm_mode", None)
    if gm is not None:
        return int(gm)
    d, q, h = config.word_embed_size, config.query_vector_dim, config.num_attention_heads
    return 1 if (d <= 316 and q <= 208 and (d // h) % 2 == 0) else 0


class NewsEncoder(nn.Module):
    """Title encoder (reference nrms_v0.py:130-176)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        emb = np.load(config.data_path + config.word_embedding_pretrained)["embeddings"].astype("float32")
        if emb.shape[1] != config.word_embed_size:
            raise ValueError("embedding width %d != config.word_embed_size %d" % (emb.shape[1], config.word_embed_size))
        self.word_embedding = nn.Sequential(
            nn.Embedding.from_pretrained(torch.tensor(emb), freeze=False, padding_idx=0).to(config.device),
            nn.Dropout(p=config.dropout, inplace=False))
        self.multihead_self_attention = MultiHeadSelfAttention(config.word_embed_size, config.num_attention_heads)
        self.additive_attention = AdditiveAttention(config.query_vector_dim, config.word_embed_size)

    def forward(self, news):
        """news: [n, num_words_title] int64 -> [n, word_embed_size]."""
        table = self.word_embedding[0].weight
        if not table.is_cuda:
            raise NrmsError("NewsEncoder runs on a CUDA device only (no CPU fallback); move the model with .to('cuda')")
        ids = news.to(table.device, dtype=torch.int64)
        p = float(self.config.dropout) if self.training else 0.0
        if p == 0.0 and not torch.is_grad_enabled():
            # inference: one shared activation blob, chunked (engine.news_encode_nograd)
            return engine.news_encode_nograd(ids, table, _gemm_mode(self.config), self.config.num_attention_heads,
                                             engine.encoder_param_list(self))
        seed = engine._next_seed(self) if p > 0 else 0
        return engine.NewsEncodeFn.apply(ids, table, p, seed, _gemm_mode(self.config),
                                         self.config.num_attention_heads,
                                         *engine.encoder_param_list(self))


class UserEncoder(nn.Module):
    """Clicked-history encoder (reference nrms_v0.py:179-199): no dropout, no history mask."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.multihead_self_attention = MultiHeadSelfAttention(config.word_embed_size, config.num_attention_heads)
        self.additive_attention = AdditiveAttention(config.query_vector_dim, config.word_embed_size)

    def forward(self, user_vector):
        """user_vector: [B, num_clicked, D] -> [B, D]."""
        if not user_vector.is_cuda:
            raise NrmsError("UserEncoder runs on a CUDA device only (no CPU fallback)")
        if not torch.is_grad_enabled():
            return engine.user_encode_nograd(user_vector, _gemm_mode(self.config), self.config.num_attention_heads,
                                             engine.encoder_param_list(self))
        return engine.UserEncodeFn.apply(user_vector, _gemm_mode(self.config),
                                         self.config.num_attention_heads,
                                         *engine.encoder_param_list(self))


class DotProductClickPredictor(nn.Module):
    """reference nrms_v0.py:201-216."""

    def forward(self, candidate_news_vector, user_vector, mask=None):
        if not candidate_news_vector.is_cuda:
            raise NrmsError("DotProductClickPredictor runs on a CUDA device only (no CPU fallback)")
        return engine.ScoreFn.apply(candidate_news_vector, user_vector, mask)


class Model(nn.Module):
    """NRMS network: 1 + K candidate titles and the clicked titles in, click logits out."""

    def __init__(self, config, pretrained_word_embedding=None):
        super().__init__()
        if not hasattr(config, "query_vector_dim"):
            raise AttributeError("config.__nrms__() must be called before building the model "
                                 "(the reference needs it too: run_v0.py:50)")
        self.config = config
        self.news_encoder = NewsEncoder(config)
        self.user_encoder = UserEncoder(config)
        self.click_predictor = DotProductClickPredictor()

    def forward(self, batch):
        """batch: the collated dict of MyDataset (data_handler.py:236-250), CPU or device
        tensors; reads browsed_titles [B,H,T], candidate_titles [B,S,T], candidate_mask [B,S].
        Returns logits [B,S] on the device, padded slots = -1e9 (reference nrms_v0.py:230-276).
        All B*(S+H) titles go through ONE fused encoder launch sequence instead of S+H
        per-slot sub-graphs; the result is independent of that grouping."""
        table = self.news_encoder.word_embedding[0].weight
        if not table.is_cuda:
            raise NrmsError("Model.forward runs on a CUDA device only (no CPU fallback); use model.to('cuda')")
        dev = table.device
        cand = batch["candidate_titles"].to(dev, dtype=torch.int64, non_blocking=True)
        clicked = batch["browsed_titles"].to(dev, dtype=torch.int64, non_blocking=True)
        B, S, T = cand.shape
        H = clicked.shape[1]
        ids = torch.cat([cand.reshape(B * S, T), clicked.reshape(B * H, T)], dim=0)
        news_vec = self.news_encoder(ids)
        cand_vec = news_vec[:B * S].view(B, S, -1)
        clicked_vec = news_vec[B * S:].view(B, H, -1)
        user_vec = self.user_encoder(clicked_vec)
        mask = batch.get("candidate_mask") if hasattr(batch, "get") else batch["candidate_mask"]
        if mask is not None:
            mask = mask.to(dev, dtype=torch.uint8, non_blocking=True).contiguous()
        return self.click_predictor(cand_vec, user_vec, mask)

    def get_news_vector(self, news):
        """[n, T] ids -> [n, D] (reference nrms_v0.py:278-289)."""
        return self.news_encoder(news)

    def get_user_vector(self, clicked_news_vector):
        """[B, H, D] -> [B, D] (reference nrms_v0.py:291-299)."""
        return self.user_encoder(clicked_news_vector)

    def get_prediction(self, news_vector, user_vector):
        """[S, D], [D] -> [S] (reference nrms_v0.py:301-312)."""
        return self.click_predictor(news_vector.unsqueeze(0), user_vector.unsqueeze(0)).squeeze(0)
