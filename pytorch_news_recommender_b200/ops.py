"""Thin host wrappers over the C-ABI (include/nrms_b200.h): torch CUDA tensors in, torch CUDA
tensors out, every byte of device memory owned by torch, every kernel ours.

torch is plumbing here (allocation, streams, autograd bookkeeping); no op below has a
PyTorch/CPU fallback — a non-CUDA tensor or a missing library raises.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import EncoderDims, NrmsError, check, ptr

DROP_EMBEDDING = 1
DROP_CONTEXT = 2


def _stream(stream=None) -> int:
    """cudaStream_t of `stream` (a torch.cuda.Stream), or of the current stream.  The fused step passes its
    side streams explicitly: a `with torch.cuda.stream(...)` block costs more host time than the launch."""
    return torch.cuda.current_stream().cuda_stream if stream is None else stream.cuda_stream


def _require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NrmsError("this path runs on a CUDA device only (sm_100a kernels); "
                            "got a CPU tensor and there is no CPU fallback")


def _cf32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise NrmsError(f"expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


@dataclass(frozen=True)
class EncoderShape:
    n_seq: int
    seq_len: int
    d_model: int
    n_heads: int
    d_query: int
    vocab: int = 0

    def dims(self, dropout_p: float = 0.0, seed: int = 0, gemm_mode: int = 0) -> EncoderDims:
        return _dims_cached(self, float(dropout_p), int(seed) & (2**64 - 1), int(gemm_mode))


@functools.lru_cache(maxsize=256)
def _dims_cached(shape: "EncoderShape", dropout_p: float, seed: int, gemm_mode: int) -> EncoderDims:
    """One ctypes struct per (shape, step): a step passes the same dims to its forward and both backward
    phases, and the host side of the fused step is on the critical path of the end-to-end loop."""
    return EncoderDims(shape.n_seq, shape.seq_len, shape.d_model, shape.n_heads, shape.d_query,
                       shape.vocab, dropout_p, gemm_mode, seed)


def encoder_param_count(d_model: int, d_query: int) -> int:
    return int(_lib.load().nrms_encoder_param_count(d_model, d_query))


def saved_bytes(shape: EncoderShape, gemm_mode: int = 0) -> int:
    n = int(_lib.load().nrms_encoder_saved_bytes(shape.dims(gemm_mode=gemm_mode)))
    if n < 0:
        check(-1, "nrms_encoder_saved_bytes")
    return n


def scratch_bytes(shape: EncoderShape, gemm_mode: int = 0) -> int:
    n = int(_lib.load().nrms_encoder_scratch_bytes(shape.dims(gemm_mode=gemm_mode)))
    if n < 0:
        check(-1, "nrms_encoder_scratch_bytes")
    return n


class BlobCache:
    """Caller-owned device blobs (saved activations / scratch), grown on demand and reused."""

    def __init__(self):
        self._blobs: Dict[str, torch.Tensor] = {}

    def get(self, name: str, nbytes: int, device) -> torch.Tensor:
        b = self._blobs.get(name)
        if b is None or b.numel() < nbytes or b.device != torch.device(device):
            b = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            self._blobs[name] = b
        return b

    def clear(self):
        self._blobs.clear()


def news_encoder_fwd(shape: EncoderShape, ids, table, params, saved, dropout_p=0.0, seed=0,
                     gemm_mode=0, out=None):
    _require_cuda(ids, table, params, saved)
    if ids.dtype != torch.int64 or not ids.is_contiguous():
        raise NrmsError("ids must be contiguous int64")
    if out is None:
        out = torch.empty((shape.n_seq, shape.d_model), dtype=torch.float32, device=table.device)
    d = shape.dims(dropout_p, seed, gemm_mode)
    check(_lib.load().nrms_news_encoder_fwd(d, ptr(ids), ptr(table), ptr(params), ptr(out),
                                            ptr(saved), saved.numel(), _stream()),
          "nrms_news_encoder_fwd")
    return out


BWD_DATA, BWD_PARAMS = 1, 2


def news_encoder_bwd(shape: EncoderShape, ids, table, params, d_out, saved, scratch, d_params,
                     d_rows, dropout_p=0.0, seed=0, gemm_mode=0, phase: Optional[int] = None):
    """phase None: the whole backward; BWD_DATA / BWD_PARAMS: its two halves (include/nrms_b200.h)."""
    _require_cuda(ids, table, params, d_out, saved, scratch, d_params, d_rows)
    d = shape.dims(dropout_p, seed, gemm_mode)
    lib = _lib.load()
    if phase is None:
        check(lib.nrms_news_encoder_bwd(d, ptr(ids), ptr(table), ptr(params), ptr(d_out), ptr(saved),
                                        saved.numel(), ptr(scratch), scratch.numel(), ptr(d_params),
                                        ptr(d_rows), _stream()), "nrms_news_encoder_bwd")
    else:
        check(lib.nrms_news_encoder_bwd_phase(d, ptr(ids), ptr(table), ptr(params), ptr(d_out), ptr(saved),
                                              saved.numel(), ptr(scratch), scratch.numel(), ptr(d_params),
                                              ptr(d_rows), int(phase), _stream()), "nrms_news_encoder_bwd_phase")


def user_encoder_fwd(shape: EncoderShape, x, params, saved, gemm_mode=0, out=None):
    _require_cuda(x, params, saved)
    if out is None:
        out = torch.empty((shape.n_seq, shape.d_model), dtype=torch.float32, device=x.device)
    d = shape.dims(0.0, 0, gemm_mode)
    check(_lib.load().nrms_user_encoder_fwd(d, ptr(x), ptr(params), ptr(out), ptr(saved),
                                            saved.numel(), _stream()), "nrms_user_encoder_fwd")
    return out


def user_encoder_fwd_gather(shape: EncoderShape, ids, table, params, saved, gemm_mode=1, out=None):
    """UserEncoder.forward over rows gathered by id from a vector table: x[s, l] = table[ids[s, l]]
    (shape.vocab = rows of the table).  No [n_seq, L, D] input tensor is materialised."""
    _require_cuda(ids, table, params, saved)
    if ids.dtype != torch.int64 or not ids.is_contiguous():
        raise NrmsError("ids must be contiguous int64")
    if out is None:
        out = torch.empty((shape.n_seq, shape.d_model), dtype=torch.float32, device=table.device)
    d = shape.dims(0.0, 0, gemm_mode)
    check(_lib.load().nrms_user_encoder_fwd_gather(d, ptr(ids), ptr(table), ptr(params), ptr(out), ptr(saved),
                                                   saved.numel(), _stream()), "nrms_user_encoder_fwd_gather")
    return out


def user_encoder_bwd(shape: EncoderShape, x, params, d_out, saved, scratch, d_params, d_x,
                     gemm_mode=0, phase: Optional[int] = None, stream=None):
    """phase None: the whole backward; BWD_DATA / BWD_PARAMS: its two halves (include/nrms_b200.h)."""
    _require_cuda(x, params, d_out, saved, scratch, d_params, d_x)
    d = shape.dims(0.0, 0, gemm_mode)
    if phase is None:
        check(_lib.load().nrms_user_encoder_bwd(d, ptr(x), ptr(params), ptr(d_out), ptr(saved),
                                                saved.numel(), ptr(scratch), scratch.numel(),
                                                ptr(d_params), ptr(d_x), _stream(stream)),
              "nrms_user_encoder_bwd")
    else:
        check(_lib.load().nrms_user_encoder_bwd_phase(d, ptr(x), ptr(params), ptr(d_out), ptr(saved),
                                                      saved.numel(), ptr(scratch), scratch.numel(),
                                                      ptr(d_params), ptr(d_x), int(phase), _stream(stream)),
              "nrms_user_encoder_bwd_phase")


def score_fwd(cand, user, mask):
    _require_cuda(cand, user, mask)
    B, Cn, D = cand.shape
    logits = torch.empty((B, Cn), dtype=torch.float32, device=cand.device)
    check(_lib.load().nrms_score_fwd(B, Cn, D, ptr(cand), ptr(user), ptr(mask), ptr(logits),
                                     _stream()), "nrms_score_fwd")
    return logits


def score_cached(vecs, cand_ids, user, mask, out=None):
    """logits[b, c] = vecs[cand_ids[b, c]] . user[b] (padded slots -1e9) straight from the vector cache."""
    _require_cuda(vecs, cand_ids, user, mask)
    B, S = cand_ids.shape
    D = vecs.shape[1]
    if out is None:
        out = torch.empty((B, S), dtype=torch.float32, device=vecs.device)
    check(_lib.load().nrms_score_cached(B, S, D, ptr(vecs), vecs.shape[0], ptr(cand_ids), ptr(user), ptr(mask),
                                        ptr(out), _stream()), "nrms_score_cached")
    return out


def score_bwd(cand, user, mask, d_logits, d_cand=None, d_user=None):
    _require_cuda(cand, user, mask, d_logits)
    B, Cn, D = cand.shape
    if d_cand is None:
        d_cand = torch.empty_like(cand)
    if d_user is None:
        d_user = torch.empty_like(user)
    check(_lib.load().nrms_score_bwd(B, Cn, D, ptr(cand), ptr(user), ptr(mask), ptr(d_logits),
                                     ptr(d_cand), ptr(d_user), _stream()), "nrms_score_bwd")
    return d_cand, d_user


def score_ce_fwd_bwd(cand, user, mask, b_global, logits, loss_rows, d_cand, d_user, loss_mean=None,
                     ticket=None):
    """loss_mean [1] float32 (optional) receives mean(loss_rows); ticket [1] int32, zero-initialised
    once by the caller, is the kernel's last-CTA counter (include/nrms_b200.h)."""
    _require_cuda(cand, user, mask, logits, loss_rows, d_cand, d_user, loss_mean, ticket)
    B, Cn, D = cand.shape
    check(_lib.load().nrms_score_ce_fwd_bwd(B, Cn, D, int(b_global), ptr(cand), ptr(user),
                                            ptr(mask), ptr(logits), ptr(loss_rows), ptr(d_cand),
                                            ptr(d_user), ptr(loss_mean), ptr(ticket), _stream()),
          "nrms_score_ce_fwd_bwd")


def embedding_plan_bytes(n_rows: int, vocab: int) -> int:
    return int(_lib.load().nrms_embedding_plan_bytes(n_rows, vocab))


def embedding_plan(ids, vocab: int, plan, stream=None):
    _require_cuda(ids, plan)
    check(_lib.load().nrms_embedding_plan(ptr(ids), ids.numel(), vocab, ptr(plan), plan.numel(),
                                          _stream(stream)), "nrms_embedding_plan")


def embedding_grad_dense(plan, d_rows, n_rows: int, vocab: int, D: int, d_table, stream=None):
    _require_cuda(plan, d_rows, d_table)
    check(_lib.load().nrms_embedding_grad_dense(ptr(plan), plan.numel(), ptr(d_rows), n_rows, vocab,
                                                D, ptr(d_table), _stream(stream)),
          "nrms_embedding_grad_dense")


def embedding_plan_unique(plan, vocab: int) -> torch.Tensor:
    out = torch.zeros(1, dtype=torch.int32, device=plan.device)
    check(_lib.load().nrms_embedding_plan_unique(ptr(plan), plan.numel(), vocab, ptr(out), _stream()),
          "nrms_embedding_plan_unique")
    return out


def adam_step(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, stream=None):
    _require_cuda(p, g, m, v)
    check(_lib.load().nrms_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), int(step), float(lr),
                                     float(beta1), float(beta2), float(eps), float(grad_scale),
                                     _stream(stream)), "nrms_adam_step")


def rank_metrics(scores, labels, offsets, max_len: int, row_stride: Optional[int] = None):
    """scores ragged [sum n_i] (row_stride None) or padded [N, row_stride]; labels uint8 ragged;
    offsets int64 [N+1].  Returns float64 [N, 4] = AUC, MRR, nDCG@5, nDCG@10."""
    _require_cuda(scores, labels, offsets)
    n = offsets.numel() - 1
    out = torch.empty((n, 4), dtype=torch.float64, device=scores.device)
    lib = _lib.load()
    if row_stride is None:
        check(lib.nrms_rank_metrics(ptr(scores), ptr(labels), ptr(offsets), n, int(max_len), ptr(out),
                                    _stream()), "nrms_rank_metrics")
    else:
        check(lib.nrms_rank_metrics_padded(ptr(scores), int(row_stride), ptr(labels), ptr(offsets), n,
                                           int(max_len), ptr(out), _stream()),
              "nrms_rank_metrics_padded")
    return out


def rank_metrics_rows(scores, labels, lens, out=None):
    """Padded on both sides: scores float32 [N, S], labels uint8 [N, S], lens int64 [N] real candidates
    per impression.  Returns float64 [N, 4] = AUC, MRR, nDCG@5, nDCG@10 (`out` may be a preallocated slice)."""
    _require_cuda(scores, labels, lens)
    n, S = scores.shape
    if out is None:
        out = torch.empty((n, 4), dtype=torch.float64, device=scores.device)
    check(_lib.load().nrms_rank_metrics_rows(ptr(scores), scores.stride(0), ptr(labels), labels.stride(0), ptr(lens),
                                             n, S, ptr(out), _stream()), "nrms_rank_metrics_rows")
    return out


def assemble_batch(index, browsed_ids, browsed_lens, candidate_ids, candidate_lens, titles):
    """One-launch batch assembly (include/nrms_b200.h: nrms_assemble_batch).  All inputs int64 CUDA
    tensors; returns the dict of batch tensors `MyDataset` + default_collate would produce for the
    keys the NRMS path reads."""
    _require_cuda(index, browsed_ids, browsed_lens, candidate_ids, candidate_lens, titles)
    B, H, S, T = index.numel(), browsed_ids.shape[1], candidate_ids.shape[1], titles.shape[1]
    dev = index.device
    out = {'browsed_lens': torch.empty(B, dtype=torch.int64, device=dev),
           'browsed_ids': torch.empty((B, H), dtype=torch.int64, device=dev),
           'browsed_titles': torch.empty((B, H, T), dtype=torch.int64, device=dev),
           'browsed_mask': torch.empty((B, H), dtype=torch.uint8, device=dev),
           'candidate_ids': torch.empty((B, S), dtype=torch.int64, device=dev),
           'candidate_titles': torch.empty((B, S, T), dtype=torch.int64, device=dev),
           'candidate_mask': torch.empty((B, S), dtype=torch.uint8, device=dev)}
    check(_lib.load().nrms_assemble_batch(
        ptr(index), B, ptr(browsed_ids), ptr(browsed_lens), ptr(candidate_ids), ptr(candidate_lens), ptr(titles),
        titles.shape[0], H, S, T, ptr(out['browsed_ids']), ptr(out['browsed_lens']), ptr(out['browsed_titles']),
        ptr(out['browsed_mask']), ptr(out['candidate_ids']), ptr(out['candidate_titles']),
        ptr(out['candidate_mask']), _stream()), "nrms_assemble_batch")
    return out


def rank_positions(scores, lens):
    """Padded scores float32 [N, S], lens int64 [N] -> int32 [N, S]: 1-based rank of every real
    candidate inside its impression (train_eval.py:279-285), 0 in the padded slots."""
    _require_cuda(scores, lens)
    scores = _cf32(scores)
    n, S = scores.shape
    out = torch.empty((n, S), dtype=torch.int32, device=scores.device)
    check(_lib.load().nrms_rank_positions(ptr(scores), S, ptr(lens.contiguous()), n, ptr(out), _stream()),
          "nrms_rank_positions")
    return out


def gather_rows(src, idx, base: int = 0):
    """out[i] = src[idx[i] - base] (zeros when idx[i] < base).  src float32 or int64 [n, D]."""
    _require_cuda(src, idx)
    n_src, D = src.shape
    out = torch.empty((idx.numel(), D), dtype=src.dtype, device=src.device)
    lib = _lib.load()
    fn = lib.nrms_gather_rows_f32 if src.dtype == torch.float32 else lib.nrms_gather_rows_i64
    if src.dtype not in (torch.float32, torch.int64):
        raise NrmsError(f"gather_rows: unsupported dtype {src.dtype}")
    check(fn(ptr(src), n_src, D, ptr(idx), idx.numel(), int(base), ptr(out), _stream()),
          "nrms_gather_rows")
    return out


def dropout_mask(seed: int, stream_id: int, p: float, n_rows: int, n_cols: int, device) -> torch.Tensor:
    """The multipliers (0 or 1/(1-p)) the encoder kernels apply to a [n_rows, n_cols] activation."""
    out = torch.empty((n_rows, n_cols), dtype=torch.float32, device=device)
    check(_lib.load().nrms_dropout_mask(int(seed) & (2**64 - 1), stream_id, float(p), n_rows, n_cols,
                                        ptr(out), _stream()), "nrms_dropout_mask")
    return out


def gemm_selftest(variant: int, A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """tcgen05 GEMM layer self-test (see include/nrms_b200.h: nrms_gemm_selftest)."""
    _require_cuda(A, B)
    A, B = _cf32(A), _cf32(B)
    if variant in (0, 3):
        (M, K), N = A.shape, B.shape[0]
    elif variant == 1:
        (M, K), N = A.shape, B.shape[1]
    else:
        (K, M), N = A.shape, B.shape[1]
    lib = _lib.load()
    nbytes = int(lib.nrms_gemm_selftest_bytes(variant, M, N, K))
    if nbytes < 0:
        raise NrmsError("nrms_gemm_selftest_bytes: bad shape")
    work = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    check(lib.nrms_gemm_selftest(variant, ptr(A), ptr(B), ptr(out), M, N, K, ptr(work), nbytes, _stream()),
          "nrms_gemm_selftest")
    return out


def validate_ids(ids, vocab: int) -> bool:
    """True when every id is inside [0, vocab) (synchronises: debugging aid, not hot path)."""
    _require_cuda(ids)
    flag = torch.zeros(1, dtype=torch.int32, device=ids.device)
    check(_lib.load().nrms_validate_ids(ptr(ids), ids.numel(), vocab, ptr(flag), _stream()),
          "nrms_validate_ids")
    return int(flag.item()) == 0


# ---- the `nrms` sibling variant (model/nrms.py; include/nrms_b200.h, last section) ------------------------
DROP_CAND_VEC = 3
DROP_HIST_VEC = 4
DROP_ATTN_PROB = 5

_linear_work: Dict[Tuple[int, int, int, int], torch.Tensor] = {}


def _linear_blob(M: int, N: int, K: int, device) -> torch.Tensor:
    """Work blob of a Linear shape, kept per device: forward and backward of a layer run on one stream, one
    after the other, and every call re-packs what it reads."""
    key = (M, N, K, device.index if device.index is not None else torch.cuda.current_device())
    blob = _linear_work.get(key)
    if blob is None:
        n = int(_lib.load().nrms_linear_work_bytes(M, N, K))
        if n < 0:
            raise NrmsError(f"nrms_linear_work_bytes: bad shape M={M} N={N} K={K} (N, K multiples of 4)")
        blob = torch.empty(n, dtype=torch.uint8, device=device)
        _linear_work[key] = blob
    return blob


def linear_fwd(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """y[M, N] = x[M, K] W[N, K]^T + bias on the tcgen05 image GEMMs (fp32-grade)."""
    _require_cuda(x, W, bias)
    x, W = _cf32(x), _cf32(W)
    (M, K), N = x.shape, W.shape[0]
    work = _linear_blob(M, N, K, x.device)
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    check(_lib.load().nrms_linear_fwd(ptr(x), ptr(W), ptr(bias), ptr(y), M, N, K, ptr(work), work.numel(),
                                      _stream()), "nrms_linear_fwd")
    return y


def linear_bwd(x, W, dy, need_dx: bool = True, need_dbias: bool = True):
    _require_cuda(x, W, dy)
    x, W, dy = _cf32(x), _cf32(W), _cf32(dy)
    (M, K), N = x.shape, W.shape[0]
    work = _linear_blob(M, N, K, x.device)
    dx = torch.empty_like(x) if need_dx else None
    dW = torch.empty_like(W)
    db = torch.empty(N, dtype=torch.float32, device=x.device) if need_dbias else None
    check(_lib.load().nrms_linear_bwd(ptr(x), ptr(W), ptr(dy), ptr(dx), ptr(dW), ptr(db), M, N, K, ptr(work),
                                      work.numel(), _stream()), "nrms_linear_bwd")
    return dx, dW, db


def dropout_apply(x: torch.Tensor, seed: int, stream_id: int, p: float) -> torch.Tensor:
    """x * multiplier of (seed, stream_id, p) over x viewed as [rows, last dim]; also its own backward."""
    _require_cuda(x)
    x = _cf32(x)
    y = torch.empty_like(x)
    n_cols = x.shape[-1]
    check(_lib.load().nrms_dropout_apply(int(seed) & (2**64 - 1), stream_id, float(p), x.numel() // n_cols,
                                         n_cols, ptr(x), ptr(y), _stream()), "nrms_dropout_apply")
    return y


def masked_attention_fwd(qkv, mask, heads: int, p_drop: float, seed: int):
    """qkv [B, L, 3E], mask [B, L] uint8 or None -> (ctx [B, L, E], probs [B, heads, L, L])."""
    _require_cuda(qkv, mask)
    qkv = _cf32(qkv)
    B, L, E3 = qkv.shape
    E = E3 // 3
    probs = torch.empty((B, heads, L, L), dtype=torch.float32, device=qkv.device)
    ctx = torch.empty((B, L, E), dtype=torch.float32, device=qkv.device)
    check(_lib.load().nrms_masked_attention_fwd(ptr(qkv), ptr(mask), B, L, heads, E // heads, float(p_drop),
                                                int(seed) & (2**64 - 1), ptr(probs), ptr(ctx), _stream()),
          "nrms_masked_attention_fwd")
    return ctx, probs


def masked_attention_bwd(qkv, mask, probs, d_ctx, heads: int, p_drop: float, seed: int):
    _require_cuda(qkv, mask, probs, d_ctx)
    d_ctx = _cf32(d_ctx)
    B, L, E3 = qkv.shape
    d_qkv = torch.empty_like(qkv)
    check(_lib.load().nrms_masked_attention_bwd(ptr(qkv), ptr(mask), ptr(probs), ptr(d_ctx), B, L, heads,
                                                E3 // 3 // heads, float(p_drop), int(seed) & (2**64 - 1),
                                                ptr(d_qkv), _stream()), "nrms_masked_attention_bwd")
    return d_qkv


def masked_pool_fwd(t, qv, x, mask):
    """t [B, L, Q] pre-tanh, qv [Q], x [B, L, E], mask [B, L] uint8 or None -> (out [B, E], alpha [B, L])."""
    _require_cuda(t, qv, x, mask)
    t, qv, x = _cf32(t), _cf32(qv), _cf32(x)
    B, L, Q = t.shape
    E = x.shape[-1]
    alpha = torch.empty((B, L), dtype=torch.float32, device=x.device)
    out = torch.empty((B, E), dtype=torch.float32, device=x.device)
    check(_lib.load().nrms_masked_pool_fwd(ptr(t), ptr(qv), ptr(x), ptr(mask), B, L, Q, E, ptr(alpha), ptr(out),
                                           _stream()), "nrms_masked_pool_fwd")
    return out, alpha


def masked_pool_bwd(t, qv, x, mask, alpha, d_out):
    _require_cuda(t, qv, x, mask, alpha, d_out)
    d_out = _cf32(d_out)
    B, L, Q = t.shape
    E = x.shape[-1]
    d_t = torch.empty_like(t)
    d_x = torch.empty_like(x)
    d_qv = torch.empty_like(qv)
    work = torch.empty((B, Q), dtype=torch.float32, device=x.device)
    check(_lib.load().nrms_masked_pool_bwd(ptr(t), ptr(qv), ptr(x), ptr(mask), ptr(alpha), ptr(d_out), B, L, Q,
                                           E, ptr(d_t), ptr(d_x), ptr(d_qv), ptr(work), _stream()),
          "nrms_masked_pool_bwd")
    return d_t, d_x, d_qv
