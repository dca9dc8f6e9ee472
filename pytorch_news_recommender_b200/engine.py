"""Host-side engine: autograd bridges for the drop-in `nn.Module` path and the fused
training step (forward + cross-entropy + backward + Adam, all in our kernels).

Two ways to train, same arithmetic:
  * drop-in  — `outputs = model(datas); loss = criterion(outputs, y); loss.backward();
               optimizer.step()` exactly as train_eval.py:189-205: the autograd Functions
               below produce `.grad` for all 19 parameters (the table gradient is dense, as
               in the reference: SURVEY §0.6) and any torch optimizer can consume them.
  * fused    — `FusedTrainer.step(datas)`: no autograd graph, gradients land in one flat
               buffer (+ the table gradient), Adam runs in our kernel on flat views that
               alias the module's parameters, and with world_size > 1 the two gradient
               buffers are all-reduced over NCCL.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import ops
from ._lib import NrmsError
from .ops import BlobCache, EncoderShape
from .parallel import GradientExchange

_ENC_PARAM_ORDER = ("W_Q.weight", "W_K.weight", "W_V.weight", "W_Q.bias", "W_K.bias", "W_V.bias",
                    "linear.weight", "linear.bias", "attention_query_vector")


def encoder_param_list(enc) -> List[torch.nn.Parameter]:
    """The 9 tensors of one encoder in the flat-block order of include/nrms_b200.h."""
    m, a = enc.multihead_self_attention, enc.additive_attention
    return [m.W_Q.weight, m.W_K.weight, m.W_V.weight, m.W_Q.bias, m.W_K.bias, m.W_V.bias,
            a.linear.weight, a.linear.bias, a.attention_query_vector]


def pack_params(plist) -> torch.Tensor:
    first = plist[0]
    # already views of one flat buffer (FusedTrainer aliases them)? then no copy
    base = getattr(first, "_nrms_flat", None)
    if base is not None and base.data_ptr() == first.data_ptr():
        return base
    return torch.cat([p.detach().reshape(-1) for p in plist])


def split_flat(flat: torch.Tensor, plist) -> List[torch.Tensor]:
    out, off = [], 0
    for p in plist:
        n = p.numel()
        out.append(flat[off:off + n].view(p.shape))
        off += n
    return out


_blobs = BlobCache()


def _next_seed(module) -> int:
    cfg = module.config
    step = getattr(module, "_nrms_dropout_calls", 0)
    module._nrms_dropout_calls = step + 1
    return (int(getattr(cfg, "dropout_seed", 0)) + step) & (2**63 - 1)


def weights_version(model) -> int:
    """Changes whenever the model's weights do: torch's in-place version counters (torch optimizers,
    load_state_dict) plus the counter `FusedTrainer.step` bumps (its Adam kernel writes the
    parameters through raw pointers, which torch's counters cannot see)."""
    return sum(int(p._version) for p in model.parameters()) + 1_000_003 * int(getattr(model, "_nrms_weights_version", 0))


# Inference (torch.no_grad()): nothing is saved for a backward, so every encoder call reuses ONE
# activation blob and long title lists are encoded in chunks of at most this many token rows — the
# reference's dev shape (512 impressions x 350 titles x 20 words = 3.6 M rows, train_eval.py:238-251)
# would otherwise allocate ~40 GB of never-read saved state per batch.
INFER_ROWS_PER_CHUNK = 1 << 19


def news_encode_nograd(ids, table, gemm_mode, n_heads, params) -> torch.Tensor:
    """Eval-mode NewsEncoder.forward (nrms_v0.py:154-176, dropout off) without autograd state."""
    ids = ids.contiguous()
    n_seq, L = ids.shape
    D, Q = table.shape[1], params[8].numel()
    flat = pack_params(params)
    out = torch.empty((n_seq, D), dtype=torch.float32, device=table.device)
    per = max(1, INFER_ROWS_PER_CHUNK // L)
    for lo in range(0, n_seq, per):
        n = min(per, n_seq - lo)
        shape = EncoderShape(n, L, D, n_heads, Q, table.shape[0])
        blob = _blobs.get("infer_saved", ops.saved_bytes(shape, gemm_mode), table.device)
        ops.news_encoder_fwd(shape, ids[lo:lo + n], table, flat, blob, 0.0, 0, gemm_mode, out=out[lo:lo + n])
    return out


def user_encode_nograd(x, gemm_mode, n_heads, params) -> torch.Tensor:
    """UserEncoder.forward (nrms_v0.py:188-199) without autograd state."""
    x = x.contiguous()
    n_seq, L, D = x.shape
    Q = params[8].numel()
    flat = pack_params(params)
    out = torch.empty((n_seq, D), dtype=torch.float32, device=x.device)
    per = max(1, INFER_ROWS_PER_CHUNK // L)
    for lo in range(0, n_seq, per):
        n = min(per, n_seq - lo)
        shape = EncoderShape(n, L, D, n_heads, Q, 0)
        blob = _blobs.get("infer_saved", ops.saved_bytes(shape, gemm_mode), x.device)
        ops.user_encoder_fwd(shape, x[lo:lo + n], flat, blob, gemm_mode, out=out[lo:lo + n])
    return out


class NewsEncodeFn(torch.autograd.Function):
    """NewsEncoder.forward (nrms_v0.py:154-176) over a flat list of titles."""

    @staticmethod
    def forward(ctx, ids, table, dropout_p, seed, gemm_mode, n_heads, *params):
        ids = ids.contiguous()
        n_seq, L = ids.shape
        D = table.shape[1]
        Q = params[8].numel()
        shape = EncoderShape(n_seq, L, D, n_heads, Q, table.shape[0])
        flat = pack_params(params)
        # a private saved blob per call: several encoder calls may be alive in one graph
        saved = torch.empty(ops.saved_bytes(shape, gemm_mode), dtype=torch.uint8, device=table.device)
        out = ops.news_encoder_fwd(shape, ids, table.detach(), flat, saved, dropout_p, seed, gemm_mode)
        ctx.shape, ctx.dropout_p, ctx.seed, ctx.gemm_mode = shape, dropout_p, seed, gemm_mode
        ctx.save_for_backward(ids, table, flat, saved)
        ctx.param_shapes = [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, d_out):
        ids, table, flat, saved = ctx.saved_tensors
        shape = ctx.shape
        dev = table.device
        scratch = _blobs.get("news_scratch", ops.scratch_bytes(shape, ctx.gemm_mode), dev)
        d_flat = torch.empty_like(flat)
        M = shape.n_seq * shape.seq_len
        d_rows = torch.empty((M, shape.d_model), dtype=torch.float32, device=dev)
        ops.news_encoder_bwd(shape, ids, table.detach(), flat, d_out.contiguous(), saved, scratch,
                             d_flat, d_rows, ctx.dropout_p, ctx.seed, ctx.gemm_mode)
        plan = _blobs.get("plan", ops.embedding_plan_bytes(M, shape.vocab), dev)
        ops.embedding_plan(ids, shape.vocab, plan)
        d_table = torch.empty_like(table)
        ops.embedding_grad_dense(plan, d_rows, M, shape.vocab, shape.d_model, d_table)
        grads, off = [], 0
        for shp in ctx.param_shapes:
            n = int(torch.Size(shp).numel())
            grads.append(d_flat[off:off + n].view(shp))
            off += n
        return (None, d_table, None, None, None, None, *grads)


class UserEncodeFn(torch.autograd.Function):
    """UserEncoder.forward (nrms_v0.py:188-199)."""

    @staticmethod
    def forward(ctx, x, gemm_mode, n_heads, *params):
        x = x.contiguous()
        n_seq, L, D = x.shape
        Q = params[8].numel()
        shape = EncoderShape(n_seq, L, D, n_heads, Q, 0)
        flat = pack_params(params)
        saved = torch.empty(ops.saved_bytes(shape, gemm_mode), dtype=torch.uint8, device=x.device)
        out = ops.user_encoder_fwd(shape, x, flat, saved, gemm_mode)
        ctx.shape, ctx.gemm_mode = shape, gemm_mode
        ctx.save_for_backward(x, flat, saved)
        ctx.param_shapes = [p.shape for p in params]
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, flat, saved = ctx.saved_tensors
        shape = ctx.shape
        scratch = _blobs.get("user_scratch", ops.scratch_bytes(shape, ctx.gemm_mode), x.device)
        d_flat = torch.empty_like(flat)
        d_x = torch.empty_like(x)
        ops.user_encoder_bwd(shape, x, flat, d_out.contiguous(), saved, scratch, d_flat, d_x,
                             ctx.gemm_mode)
        grads, off = [], 0
        for shp in ctx.param_shapes:
            n = int(torch.Size(shp).numel())
            grads.append(d_flat[off:off + n].view(shp))
            off += n
        return (d_x, None, None, *grads)


class ScoreFn(torch.autograd.Function):
    """DotProductClickPredictor.forward + candidate masking (nrms_v0.py:205-216, 272-274)."""

    @staticmethod
    def forward(ctx, cand, user, mask):
        cand, user = cand.contiguous(), user.contiguous()
        logits = ops.score_fwd(cand, user, mask)
        ctx.save_for_backward(cand, user)
        ctx.mask = mask
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        cand, user = ctx.saved_tensors
        d_cand, d_user = ops.score_bwd(cand, user, ctx.mask, d_logits.contiguous())
        return d_cand, d_user, None


class DeviceAdam:
    """torch.optim.Adam with its defaults (train_eval.py:167) as ONE `nrms_adam_step` launch per parameter
    tensor — the optimizer of the autograd-driven plugins (the `nrms` sibling), whose step is not fused into a
    `FusedTrainer`.  Same surface as far as the reference's loop uses it: `param_groups[i]['lr']`,
    `zero_grad()`, `step()`."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params = [p for p in params if p.requires_grad]
        self.param_groups = [{"params": self.params, "lr": float(lr)}]
        self.betas, self.eps = betas, eps
        self.state: Dict[torch.nn.Parameter, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.step_count = 0

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        self.step_count += 1
        lr = float(self.param_groups[0]["lr"])
        for p in self.params:
            if p.grad is None:
                continue
            if not p.is_cuda:
                raise NrmsError("DeviceAdam runs on CUDA parameters only (no CPU fallback)")
            st = self.state.get(p)
            if st is None:
                st = self.state[p] = (torch.zeros_like(p), torch.zeros_like(p))
            ops.adam_step(p.data, p.grad.contiguous(), st[0], st[1], self.step_count, lr, self.betas[0],
                          self.betas[1], self.eps)
            torch.autograd.graph.increment_version(p)     # the kernel wrote through a raw pointer


# ------------------------------------------------------------------------------------------
# fused training step
# ------------------------------------------------------------------------------------------
class FusedTrainer:
    """forward + CrossEntropyLoss(label 0) + backward + Adam (train_eval.py:189-205) without an
    autograd graph.  The module's parameters are re-pointed at views of two flat buffers
    (dense block: news encoder then user encoder; table separate), so `state_dict()` and the
    drop-in forward keep seeing the live weights.

    Data parallel: one process per GPU; every rank passes its own shard of the
    global batch; gradients carry 1/B_global, are SUM-all-reduced, then every rank applies
    the same Adam update."""

    def __init__(self, model, lr: Optional[float] = None, betas=(0.9, 0.999), eps: float = 1e-8,
                 process_group=None, table_sync: str = "auto"):
        self.model = model
        cfg = model.config
        self.cfg = cfg
        self.lr = float(cfg.learning_rate if lr is None else lr)
        self.betas, self.eps = betas, eps
        self.pg = process_group
        self.exchange = GradientExchange(process_group)
        self.world = self.exchange.world
        if table_sync == "auto":
            # measured on 8 B200s (profiles/r02_scale_*): sharded 1.74 ms/step vs dense 1.84 at cfg2
            table_sync = "sharded" if self.world > 1 else "dense"
        if table_sync not in ("dense", "sharded"):
            raise ValueError("table_sync must be 'dense' (all-reduce + replicated Adam) or 'sharded' "
                             "(reduce-scatter + Adam on V/G rows + all-gather)")
        self.table_sync = table_sync
        self.step_count = 0
        # False: every kernel of the step on the caller's stream, in order (bench.py's per-kernel timing pass:
        # CUDA events around a launch on a side stream would also count the time it waits for SMs)
        self.overlap = True
        self.table = model.news_encoder.word_embedding[0].weight
        if not self.table.is_cuda:
            raise NrmsError("FusedTrainer needs the model on a CUDA device (no CPU fallback)")
        dev = self.table.device
        self.device = dev
        self.news_params = encoder_param_list(model.news_encoder)
        self.user_params = encoder_param_list(model.user_encoder)
        n_enc = sum(p.numel() for p in self.news_params)
        self.n_enc = n_enc
        flat = torch.empty(2 * n_enc, dtype=torch.float32, device=dev)
        off = 0
        for plist in (self.news_params, self.user_params):
            block = flat[off:off + n_enc]
            o = 0
            for p in plist:
                n = p.numel()
                view = block[o:o + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                o += n
            plist[0]._nrms_flat = block
            off += n_enc
        self.flat = flat
        self.flat_grad = torch.zeros_like(flat)
        self.flat_m = torch.zeros_like(flat)
        self.flat_v = torch.zeros_like(flat)
        V, D = self.table.shape
        if self.world > 1 and table_sync == "sharded":
            # ZeRO-1 style: the table gradient is reduce-SCATTERED, every rank runs Adam on its V/G rows
            # only (moments exist for those rows only) and the updated rows are all-gathered.  Both
            # collectives need world equal slices: table and gradient live in buffers padded to
            # V_pad = G * ceil(V / G) rows; the nn.Embedding weight is re-pointed at the first V rows.
            self.v_shard = (V + self.world - 1) // self.world
            v_pad = self.v_shard * self.world
            self.table_pad = torch.zeros((v_pad, D), dtype=torch.float32, device=dev)
            self.table_pad[:V].copy_(self.table.data)
            self.table.data = self.table_pad[:V]
            self.table_grad_pad = torch.zeros((v_pad, D), dtype=torch.float32, device=dev)
            self.table_grad = self.table_grad_pad[:V]
            lo = self.exchange.rank * self.v_shard
            self.table_shard = self.table_pad[lo:lo + self.v_shard]
            self.grad_shard = torch.empty((self.v_shard, D), dtype=torch.float32, device=dev)
            self.table_m = torch.zeros((self.v_shard, D), dtype=torch.float32, device=dev)
            self.table_v = torch.zeros((self.v_shard, D), dtype=torch.float32, device=dev)
            self._comm = torch.cuda.Stream(device=dev)
            # The small dense-block all-reduce gets its OWN communicator: collectives of one NCCL communicator
            # run in issue order on one stream, so behind the table's reduce-scatter / all-gather it would
            # wait for both although it shares no data with them.
            ranks = dist.get_process_group_ranks(self.pg) if self.pg is not None else list(range(dist.get_world_size()))
            self.exchange_flat = GradientExchange(dist.new_group(ranks=ranks))
        else:
            self.table_grad = torch.empty_like(self.table.data)
            self.table_m = torch.zeros_like(self.table.data)
            self.table_v = torch.zeros_like(self.table.data)
        self.blobs = BlobCache()
        self._bufs: Dict[Tuple, Dict[str, torch.Tensor]] = {}
        if self.world > 1:
            # replicas must START identical: rank 0's weights win (a rank that seeded differently would
            # otherwise diverge silently; the updates are identical by construction afterwards)
            dist.broadcast(self.flat, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
            dist.broadcast(self.table.data, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
        # last-CTA ticket of the scorer's loss mean (include/nrms_b200.h) and the loss itself
        self._ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self._loss_dev = torch.zeros(1, dtype=torch.float32, device=dev)

    # -- persistent device buffers per batch shape ----------------------------------------
    def _buffers(self, B, C, H, T):
        key = (B, C, H, T)
        b = self._bufs.get(key)
        if b is None:
            dev, D = self.device, self.table.shape[1]
            n_titles = B * (C + H)
            b = {
                # two input slots: the next batch's H2D copy (prefetch) lands in one while the step in
                # flight reads the other; "ids" / "mask" name the slot of the current step
                "ids_slots": [torch.empty((n_titles, T), dtype=torch.int64, device=dev) for _ in range(2)],
                "mask_slots": [torch.empty((B, C), dtype=torch.uint8, device=dev) for _ in range(2)],
                "slot_free": [None, None], "slot": 0,
                "news_vec": torch.empty((n_titles, D), dtype=torch.float32, device=dev),
                "d_news_vec": torch.empty((n_titles, D), dtype=torch.float32, device=dev),
                "user_vec": torch.empty((B, D), dtype=torch.float32, device=dev),
                "d_user_vec": torch.empty((B, D), dtype=torch.float32, device=dev),
                "logits": torch.empty((B, C), dtype=torch.float32, device=dev),
                "loss_rows": torch.empty((B,), dtype=torch.float32, device=dev),
                "d_rows": torch.empty((n_titles * T, D), dtype=torch.float32, device=dev),
            }
            b["ids"], b["mask"] = b["ids_slots"][0], b["mask_slots"][0]
            self._bufs[key] = b
        return b

    @staticmethod
    def _copy_in(b, slot, batch, B, C, H, T):
        ids, mask = b["ids_slots"][slot], b["mask_slots"][slot]
        ids[:B * C].view(B, C, T).copy_(batch["candidate_titles"], non_blocking=True)
        ids[B * C:].view(B, H, T).copy_(batch["browsed_titles"], non_blocking=True)
        mask.copy_(batch["candidate_mask"], non_blocking=True)

    def load_batch(self, batch) -> Dict[str, torch.Tensor]:
        """H2D copy of the three tensors the path reads (nrms_v0.py:248-250,272) into the
        persistent device buffers; candidate titles first, then clicked titles.  A batch announced
        with `prefetch` is already on its way: the step only waits for that copy."""
        ct, bt = batch["candidate_titles"], batch["browsed_titles"]
        B, C, T = ct.shape
        H = bt.shape[1]
        b = self._buffers(B, C, H, T)
        pf = getattr(self, "_prefetched", None)
        if pf is not None and pf[0] is ct and pf[1] is b:
            slot, ready = pf[2], pf[3]
            torch.cuda.current_stream(self.device).wait_event(ready)
            self._prefetched = None
        else:
            if pf is not None and pf[1] is b:     # an announced batch was dropped: let its copy finish first
                torch.cuda.current_stream(self.device).wait_event(pf[3])
                self._prefetched = None
            slot = b["slot"] ^ 1
            self._copy_in(b, slot, batch, B, C, H, T)
        b["slot"] = slot
        b["ids"], b["mask"] = b["ids_slots"][slot], b["mask_slots"][slot]
        b["dims"] = (B, C, H, T)
        return b

    def prefetch(self, batch) -> None:
        """Start the H2D copy of the NEXT step's host batch on a copy stream, underneath the step in
        flight (call it right after `step`).  The batch must stay alive and unchanged (pinned memory
        for a truly asynchronous copy) until the step that consumes it has been issued."""
        ct, bt = batch["candidate_titles"], batch["browsed_titles"]
        B, C, T = ct.shape
        H = bt.shape[1]
        b = self._buffers(B, C, H, T)
        slot = b["slot"] ^ 1                      # the slot the step in flight does not read
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        if b["slot_free"][slot] is not None:
            cs.wait_event(b["slot_free"][slot])   # its last reader (the step before the one in flight) is done
        with torch.cuda.stream(cs):
            self._copy_in(b, slot, batch, B, C, H, T)
            ready = torch.cuda.Event()
            ready.record(cs)
        self._prefetched = (ct, b, slot, ready)

    def step(self, batch, b_global: Optional[int] = None) -> torch.Tensor:
        """One optimisation step on this rank's shard; returns the mean loss of the shard as
        a 0-dim device tensor (no host sync)."""
        b = batch if "dims" in batch else self.load_batch(batch)
        B, C, H, T = b["dims"]
        cfg, dev = self.cfg, self.device
        D, V = self.table.shape[1], self.table.shape[0]
        h, Q = cfg.num_attention_heads, cfg.query_vector_dim
        from .model.nrms_v0 import _gemm_mode
        gm = _gemm_mode(cfg)
        training = self.model.training
        p = float(cfg.dropout) if training else 0.0
        self.step_count += 1
        # every rank draws its own masks: the kernels address the Philox stream by LOCAL row, so the rank is
        # mixed into the key (identical masks on every shard would correlate the noise across ranks)
        seed = (int(getattr(cfg, "dropout_seed", 0)) + self.step_count
                + 0x9E3779B97F4A7C15 * self.exchange.rank) & (2**63 - 1)
        if b_global is None:
            b_global = B * self.world              # even shards; see `global_batch` for ragged last batches
        n_titles = B * (C + H)
        # shapes and blob sizes depend only on (batch shape, GEMM mode): looked up once, not per step
        # (a step that follows a host read-back has this prologue on its critical path)
        meta = b.get("meta")
        if meta is None or meta[0] != (B, C, H, T, gm, V):
            ns, us = EncoderShape(n_titles, T, D, h, Q, V), EncoderShape(B, H, D, h, Q, 0)
            meta = ((B, C, H, T, gm, V), ns, us, ops.saved_bytes(ns, gm), ops.saved_bytes(us, gm),
                    max(ops.scratch_bytes(ns, gm), ops.scratch_bytes(us, gm)),
                    ops.embedding_plan_bytes(n_titles * T, V))
            b["meta"] = meta
        _, news_shape, user_shape, news_saved_b, user_saved_b, scratch_b, plan_b = meta
        news_flat, user_flat = self.flat[:self.n_enc], self.flat[self.n_enc:]
        news_saved = self.blobs.get("news_saved", news_saved_b, dev)
        user_saved = self.blobs.get("user_saved", user_saved_b, dev)
        news_scratch = self.blobs.get("scratch", scratch_b, dev)
        table = self.table.data
        # counting sort of the step's token ids (only needs the ids: off the backward's critical path)
        # (five small latency-bound launches: they run on a side stream underneath the forward)
        plan = self.blobs.get("plan", plan_b, dev)
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=dev)
            self._side2 = torch.cuda.Stream(device=dev)
        side = self._side if self.overlap else main
        side2 = self._side2 if self.overlap else main
        if side is not main:
            side.wait_stream(main)            # the ids are on the device; last step's use of the plan is over
        ops.embedding_plan(b["ids"], V, plan, stream=side)
        # ---- forward ----------------------------------------------------------------------
        ops.news_encoder_fwd(news_shape, b["ids"], table, news_flat, news_saved, p, seed, gm,
                             out=b["news_vec"])
        cand_vec = b["news_vec"][:B * C].view(B, C, D)
        hist_vec = b["news_vec"][B * C:].view(B, H, D)
        ops.user_encoder_fwd(user_shape, hist_vec, user_flat, user_saved, gm, out=b["user_vec"])
        # ---- scorer + CE + their backward ---------------------------------------------------
        d_cand = b["d_news_vec"][:B * C].view(B, C, D)
        d_hist = b["d_news_vec"][B * C:].view(B, H, D)
        if getattr(self, "_loss_event", None) is not None:
            main.wait_event(self._loss_event)     # the previous step's 4-byte read-back of _loss_dev is done
        ops.score_ce_fwd_bwd(cand_vec, b["user_vec"], b["mask"], b_global, b["logits"],
                             b["loss_rows"], d_cand, b["d_user_vec"], loss_mean=self._loss_dev,
                             ticket=self._ticket)
        # The loss is final here, half a step before the optimizer is: it starts its way to the host
        # now, so that `last_loss()` (the `loss.item()` of train_eval.py:198) does not wait for the
        # backward and Adam — and the host can already enqueue the next step while they run.
        loss = self._loss_dev[0]     # the mean over the shard, summed in index order by the scorer's last CTA
        if getattr(self, "_loss_host", None) is None:
            self._loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
            self._loss_event, self._loss_ready = torch.cuda.Event(), torch.cuda.Event()
            self._loss_stream = torch.cuda.Stream(device=dev)
        # (on its own stream: a D2H copy in the compute stream would hold back the backward's first kernel)
        self._loss_ready.record(torch.cuda.current_stream(dev))
        self._loss_stream.wait_event(self._loss_ready)
        with torch.cuda.stream(self._loss_stream):
            self._loss_host.copy_(self._loss_dev, non_blocking=True)
            self._loss_event.record(self._loss_stream)
        # ---- backward -----------------------------------------------------------------------
        # The step is a serial chain of kernels that each load ONE resource (tensor pipe, HBM, or just
        # latency); three pieces are off that chain and run on side streams next to it:
        #   * the user encoder's weight gradients (a dozen tiny launches) underneath the news encoder's
        #     data-gradient path;
        #   * the table path — deduplicating reduction of the embedding-row gradients, then (one GPU) the
        #     HBM-bound dense Adam over the 84 MB table, or (data parallel) reduce-scatter -> Adam on this
        #     rank's rows -> all-gather — underneath the news encoder's tensor-bound weight-gradient GEMMs.
        user_scratch = self.blobs.get("user_scratch", ops.scratch_bytes(user_shape, gm), dev)
        ub = (user_shape, hist_vec, user_flat, b["d_user_vec"], user_saved, user_scratch,
              self.flat_grad[self.n_enc:], d_hist, gm)
        ops.user_encoder_bwd(*ub, phase=ops.BWD_DATA)
        if side2 is not main:
            side2.wait_stream(main)
        ops.user_encoder_bwd(*ub, phase=ops.BWD_PARAMS, stream=side2)
        M = n_titles * T
        nb = (news_shape, b["ids"], table, news_flat, b["d_news_vec"], news_saved, news_scratch,
              self.flat_grad[:self.n_enc], b["d_rows"], p, seed, gm)
        ops.news_encoder_bwd(*nb, phase=ops.BWD_DATA)
        b1, b2 = self.betas
        sharded = self.world > 1 and self.table_sync == "sharded"
        pending = []
        if self.world > 1 and not sharded:
            # dense exchange: all-reduce of the whole table gradient underneath the weight-gradient GEMMs, then
            # the replicated Adam
            if side is not main:
                main.wait_stream(side)
            ops.embedding_grad_dense(plan, b["d_rows"], M, V, D, self.table_grad)
            pending = self.exchange.allreduce([self.table_grad], async_op=True)
        else:
            tstream = (self._comm if self.overlap else main) if sharded else side   # (the plan was computed on `side`)
            if tstream is not main:
                tstream.wait_stream(main)
            if sharded and side is not tstream:
                tstream.wait_stream(side)
            ops.embedding_grad_dense(plan, b["d_rows"], M, V, D, self.table_grad, stream=tstream)
            if sharded:
                with torch.cuda.stream(tstream):      # (torch.distributed takes the current stream)
                    self.exchange.reduce_scatter(self.table_grad_pad, self.grad_shard)
                    ops.adam_step(self.table_shard, self.grad_shard, self.table_m, self.table_v, self.step_count,
                                  self.lr, b1, b2, self.eps)
                    self.exchange.all_gather(self.table_pad, self.table_shard)
            else:
                ops.adam_step(table, self.table_grad, self.table_m, self.table_v, self.step_count, self.lr,
                              b1, b2, self.eps, stream=tstream)
        ops.news_encoder_bwd(*nb, phase=ops.BWD_PARAMS)
        if side2 is not main:
            main.wait_stream(side2)               # the user encoder's half of flat_grad
        pending += (self.exchange_flat if sharded else self.exchange).allreduce([self.flat_grad], async_op=True)
        for h_ in pending:
            h_.wait()
        # ---- Adam ---------------------------------------------------------------------------
        ops.adam_step(self.flat, self.flat_grad, self.flat_m, self.flat_v, self.step_count, self.lr,
                      b1, b2, self.eps)
        if self.world > 1 and not sharded:
            ops.adam_step(table, self.table_grad, self.table_m, self.table_v, self.step_count, self.lr,
                          b1, b2, self.eps)
        elif tstream is not main:
            main.wait_stream(tstream)             # the updated table is what the next forward reads
        self.model._nrms_weights_version = getattr(self.model, "_nrms_weights_version", 0) + 1
        if "slot_free" in b:                      # this step's input slot may be overwritten from here on
            ev = b["slot_free"][b["slot"]]
            if ev is None:
                ev = b["slot_free"][b["slot"]] = torch.cuda.Event()
            ev.record(main)
        return loss

    def global_batch(self, B: int) -> int:
        """Sum of the ranks' shard sizes (one small all-reduce + host read).  `step` assumes even
        shards (`B * world`); a loader with drop_last=False (the reference's, data_handler.py) leaves a
        ragged last batch, for which the caller passes `b_global=trainer.global_batch(B)`."""
        if self.world == 1:
            return int(B)
        t = torch.tensor([B], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
        return int(t.item())

    def last_loss(self) -> float:
        """The mean loss of the last `step` as a Python float: waits only for the forward + loss of
        that step (its 4-byte D2H copy), not for its backward and optimizer update."""
        self._loss_event.synchronize()
        return float(self._loss_host[0])

    def grads_as_state_dict(self) -> Dict[str, torch.Tensor]:
        """Last step's gradients keyed like model.state_dict() (for parity tests)."""
        tg = self.table_grad
        if self.world > 1 and self.table_sync == "sharded":
            # the reduced gradient exists as one shard per rank: gather it (collective: every rank calls)
            full = torch.empty_like(self.table_grad_pad)
            self.exchange.all_gather(full, self.grad_shard)
            tg = full[:self.table.shape[0]]
        out = {"news_encoder.word_embedding.0.weight": tg}
        for prefix, lo in (("news_encoder.", 0), ("user_encoder.", self.n_enc)):
            block = self.flat_grad[lo:lo + self.n_enc]
            plist = self.news_params if lo == 0 else self.user_params
            names = ["multihead_self_attention." + n for n in _ENC_PARAM_ORDER[:6]] + \
                    ["additive_attention." + n for n in _ENC_PARAM_ORDER[6:]]
            for name, g in zip(names, split_flat(block, plist)):
                out[prefix + name] = g
        return out
