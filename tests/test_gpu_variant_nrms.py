"""GPU parity of the `nrms` sibling variant (reference model/nrms.py): every new C-ABI op
against a float64 torch statement of the reference's arithmetic, and the whole plugin against the fixtures
generated from the reference module itself (tests/golden/make_golden_bert.py) and the oracle fed with the
kernels' own dropout masks.  Tolerances: scores 1e-3 relative (north_star), gradients 2e-3 of the norm."""
import math
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import nrms_bert_oracle as OB
from _golden import BertCase, check_summary

pytestmark = pytest.mark.gpu


def _close(got, want, tol, what):
    got, want = got.double().cpu(), want.double().cpu()
    err = (got - want).abs().max().item()
    scale = max(want.abs().max().item(), 1e-30)
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("M,N,K", [(100, 48, 52), (3520, 512, 512), (320, 1536, 512), (77, 400, 512), (1, 4, 4),
                                   (700, 244, 640)])
def test_linear_fwd_bwd(M, N, K, built_lib):
    from pytorch_news_recommender_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    y = ops.linear_fwd(x.cuda(), W.cuda(), b.cuda())
    _close(y, x.double() @ W.double().t() + b.double(), 2e-5, "y")
    _close(ops.linear_fwd(x.cuda(), W.cuda(), None), x.double() @ W.double().t(), 2e-5, "y (no bias)")
    dx, dW, db = ops.linear_bwd(x.cuda(), W.cuda(), dy.cuda())
    _close(dx, dy.double() @ W.double(), 2e-5, "dx")
    _close(dW, dy.double().t() @ x.double(), 2e-5, "dW")
    _close(db, dy.double().sum(0), 1e-5, "db")
    dx2, dW2, db2 = ops.linear_bwd(x.cuda(), W.cuda(), dy.cuda(), need_dx=False, need_dbias=False)
    assert dx2 is None and db2 is None and torch.equal(dW2, dW)       # same kernels, same order: bitwise


def test_masked_attention_rejects_what_does_not_fit(built_lib):
    from pytorch_news_recommender_b200 import ops
    from pytorch_news_recommender_b200._lib import NrmsError
    for B, L, E3, heads in ((1, 129, 3 * 64, 2), (1, 128, 3 * 128, 1), (1, 16, 3 * 160, 1)):   # L > 128; smem; head dim > 128
        with pytest.raises(NrmsError):
            ops.masked_attention_fwd(torch.zeros(B, L, E3, device="cuda"), None, heads, 0.0, 0)


def _attention_ref(qkv, mask, heads, mult):
    B, L, E3 = qkv.shape
    E = E3 // 3
    dk = E // heads
    q, k, v = [t.view(B, L, heads, dk).transpose(1, 2) for t in qkv.split(E, dim=-1)]
    s = q @ k.transpose(-1, -2) / math.sqrt(dk)
    if mask is not None:
        m = mask.to(s.dtype)
        s = s.masked_fill((m.unsqueeze(1) * m.unsqueeze(2)).unsqueeze(1).expand_as(s) == 0, -1e9)
    p = torch.softmax(s, -1)
    pd = p if mult is None else p * mult
    return (pd @ v).transpose(1, 2).reshape(B, L, E), p


@pytest.mark.parametrize("B,L,heads,dk,p", [(3, 7, 4, 12, 0.0), (5, 50, 8, 64, 0.2), (2, 96, 2, 32, 0.1),
                                            (4, 33, 5, 100, 0.0)])
def test_masked_attention(B, L, heads, dk, p, built_lib):
    from pytorch_news_recommender_b200 import ops
    g = torch.Generator().manual_seed(L)
    E = heads * dk
    qkv = torch.randn(B, L, 3 * E, generator=g, dtype=torch.float64)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    lens[0] = L
    if B > 1:
        lens[1] = 0                                         # an all-padding history: uniform rows, no gradient
    mask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.uint8)
    d_ctx = torch.randn(B, L, E, generator=g, dtype=torch.float64)
    seed = 77
    for use_mask in (True, False):
        mk = mask if use_mask else None
        mult = None
        if p > 0:
            mult = ops.dropout_mask(seed, ops.DROP_ATTN_PROB, p, B * heads * L, L, "cuda").view(B, heads, L, L).cpu().double()
            assert 0 < (mult == 0).double().mean().item() < 2 * p
        q64 = qkv.clone().requires_grad_(True)
        want, probs_want = _attention_ref(q64, mk, heads, mult)
        want.backward(d_ctx)
        q32 = qkv.float().cuda()
        mk_d = mk.cuda() if mk is not None else None
        ctx, probs = ops.masked_attention_fwd(q32, mk_d, heads, p, seed)
        _close(probs, probs_want.detach(), 2e-5, "probs")
        _close(ctx, want.detach(), 2e-5, "ctx")
        d_qkv = ops.masked_attention_bwd(q32, mk_d, probs, d_ctx.float().cuda(), heads, p, seed)
        _close(d_qkv, q64.grad, 5e-5, "d_qkv")
        if use_mask and B > 1:
            assert float(d_qkv[1, :, :2 * E].abs().max()) == 0.0        # masked scores pass no gradient to Q, K


@pytest.mark.parametrize("B,L,Q,E", [(3, 7, 20, 48), (6, 50, 400, 512), (2, 130, 36, 40)])
def test_masked_pool(B, L, Q, E, built_lib):
    from pytorch_news_recommender_b200 import ops
    g = torch.Generator().manual_seed(L + Q)
    t = torch.randn(B, L, Q, generator=g, dtype=torch.float64)
    qv = torch.rand(Q, generator=g, dtype=torch.float64) * 0.2 - 0.1
    x = torch.randn(B, L, E, generator=g, dtype=torch.float64)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    lens[0] = L
    lens[-1] = 0
    mask = (torch.arange(L)[None, :] < lens[:, None]).to(torch.uint8)
    d_out = torch.randn(B, E, generator=g, dtype=torch.float64)
    for use_mask in (True, False):
        mk = mask if use_mask else None
        t64, q64, x64 = [v.clone().requires_grad_(True) for v in (t, qv, x)]
        s = torch.tanh(t64) @ q64
        if mk is not None:
            s = s.masked_fill(mk == 0, -1e9)
        w = torch.softmax(s, 1)
        out = torch.bmm(w.unsqueeze(1), x64).squeeze(1)
        out.backward(d_out)
        mk_d = mk.cuda() if mk is not None else None
        got, alpha = ops.masked_pool_fwd(t.float().cuda(), qv.float().cuda(), x.float().cuda(), mk_d)
        _close(alpha, w.detach(), 2e-5, "alpha")
        _close(got, out.detach(), 2e-5, "out")
        d_t, d_x, d_qv = ops.masked_pool_bwd(t.float().cuda(), qv.float().cuda(), x.float().cuda(), mk_d, alpha,
                                             d_out.float().cuda())
        _close(d_t, t64.grad, 5e-5, "d_t")
        _close(d_x, x64.grad, 2e-5, "d_x")
        _close(d_qv, q64.grad, 5e-5, "d_qv")


def test_dropout_apply_is_the_exported_mask(built_lib):
    from pytorch_news_recommender_b200 import ops
    x = torch.randn(37, 52, device="cuda")
    for sid in (ops.DROP_CAND_VEC, ops.DROP_HIST_VEC):
        m = ops.dropout_mask(5, sid, 0.3, 37, 52, "cuda")
        assert torch.equal(ops.dropout_apply(x, 5, sid, 0.3), x * m)
    assert torch.equal(ops.dropout_apply(x, 5, 3, 0.0), x)
    assert not torch.equal(ops.dropout_mask(5, 3, 0.3, 37, 52, "cuda"), ops.dropout_mask(5, 4, 0.3, 37, 52, "cuda"))


# ---- the plugin against the reference's own outputs ----------------------------------------------------------
def _config_from_case(c, dropout=None):
    from pytorch_news_recommender_b200.config import Config
    from pytorch_news_recommender_b200 import synthetic as S
    cfg = Config("NRMS_BERT_TEST")
    cfg.__nrms__()
    cfg.history_len, cfg.sample_size = c.H, c.C - 1
    cfg.bert_embed_size = cfg.news_feature_size = c.E
    cfg.user_heads_num, cfg.query_vector_dim_large = c.heads, c.Q
    cfg.dropout = c.cfg.dropout if dropout is None else dropout
    tmp = tempfile.mkdtemp()
    S.save_embedding_npz(os.path.join(tmp, "bert.npz"), c.table)
    cfg.data_path, cfg.bert_embedding_pretrained = tmp + "/", "bert.npz"
    cfg.device = torch.device("cuda:0")
    return cfg


@pytest.mark.parametrize("name", ["bert_tiny", "bert_mind"])
def test_plugin_init_eval_and_grads_match_reference(name, built_lib):
    from pytorch_news_recommender_b200.model import nrms as plugin
    c = BertCase(name)
    cfg = _config_from_case(c)
    torch.manual_seed(42)
    model = plugin.Model(cfg).to(cfg.device)
    assert list(model.state_dict().keys()) == OB.state_dict_keys()
    check_summary(c, "sd0sum", dict(model.state_dict()), rtol=0.0, atol_frac=0.0, abs_floor=0.0)
    model.eval()
    with torch.no_grad():
        logits = model(c.batch)
    assert logits.is_cuda and logits.shape == (c.B, c.C)
    want = torch.from_numpy(c.z["eval/logits"])
    live = c.batch["candidate_mask"].bool()
    assert (logits.cpu()[~live] == -1e9).all()
    err = ((logits.cpu() - want).abs() / want.abs().clamp_min(1e-3))[live].max().item()
    assert err <= 1e-3, err
    # eval-mode loss + all gradients vs the reference's autograd
    model.zero_grad()
    out = model(c.batch)
    loss = torch.nn.functional.cross_entropy(out, torch.zeros(c.B, dtype=torch.long, device=out.device))
    loss.backward()
    assert abs(loss.item() - float(c.z["evalgrad/loss"])) < 1e-4
    grads = {k: p.grad for k, p in model.named_parameters()}
    check_summary(c, "evalgrad", grads, rtol=2e-3, atol_frac=2e-3)
    assert float(grads[OB.TABLE_KEY][0].abs().max()) == 0.0


@pytest.mark.parametrize("name", ["bert_tiny", "bert_mind"])
def test_plugin_train_mode_matches_oracle_with_the_kernels_masks(name, built_lib):
    """Train mode: ATen's Bernoulli stream cannot be reproduced, so the oracle (pinned to the reference with
    injected masks, tests/test_oracle_golden.py) is fed the multipliers the kernels use."""
    from pytorch_news_recommender_b200 import ops
    from pytorch_news_recommender_b200.model import nrms as plugin
    c = BertCase(name)
    cfg = _config_from_case(c)
    cfg.dropout_seed = 1234
    torch.manual_seed(42)
    model = plugin.Model(cfg).to(cfg.device)
    sd = c.state_dict()
    model.load_state_dict(sd)
    model.train()
    p = cfg.dropout
    for step in range(2):
        seed = cfg.dropout_seed + step
        mults = {
            "cand": ops.dropout_mask(seed, ops.DROP_CAND_VEC, p, c.B * c.C, c.E, "cuda").view(c.B, c.C, c.E).cpu(),
            "hist": ops.dropout_mask(seed, ops.DROP_HIST_VEC, p, c.B * c.H, c.E, "cuda").view(c.B, c.H, c.E).cpu(),
            "attn": ops.dropout_mask(seed, ops.DROP_ATTN_PROB, p, c.B * c.heads * c.H, c.H, "cuda")
                       .view(c.B, c.heads, c.H, c.H).cpu(),
        }
        want_loss, want_logits, want_grads = OB.loss_and_grads(sd, c.batch, c.cfg, mults)
        model.zero_grad()
        out = model(c.batch)
        loss = torch.nn.functional.cross_entropy(out, torch.zeros(c.B, dtype=torch.long, device=out.device))
        loss.backward()
        live = c.batch["candidate_mask"].bool()
        err = ((out.detach().cpu() - want_logits).abs() / want_logits.abs().clamp_min(1e-3))[live].max().item()
        assert err <= 1e-3, (step, err)
        assert abs(loss.item() - float(want_loss)) < 1e-4
        for k, prm in model.named_parameters():
            w = want_grads[k].double()
            e = (prm.grad.double().cpu() - w).norm().item()
            # the key projection's bias gradient is mathematically zero (softmax is invariant to a per-query
            # constant): only rounding noise is left on either side, hence the absolute floor
            assert e <= 2e-3 * w.norm().item() + 1e-7 * math.sqrt(w.numel()), (step, k, e, w.norm().item())


def test_plugin_trains_with_the_reference_loop_and_reloads(built_lib):
    """The literal statements of train_eval.py:189-205 over the plugin: the loss falls; state_dict round-trips
    through the reference's key names; the plugin wrapper resolves `args.model = 'NRMS'`."""
    from types import SimpleNamespace
    from pytorch_news_recommender_b200 import model as model_pkg
    c = BertCase("bert_mind")
    cfg = _config_from_case(c, dropout=0.2)
    torch.manual_seed(42)
    wrapped = model_pkg.Model(cfg, SimpleNamespace(model="NRMS", n_GPUs=1))
    model = wrapped.model
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3)
    criterion = torch.nn.CrossEntropyLoss()
    model.train()
    losses = []
    for _ in range(12):
        outputs = wrapped(c.batch)
        labels = torch.zeros(len(outputs)).long().to(outputs.device)
        model.zero_grad()
        loss = criterion(outputs, labels)
        loss.backward()
        optimizer.step()
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0], losses
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    torch.manual_seed(1)
    again = model_pkg.nrms.Model(cfg).to(cfg.device)
    again.load_state_dict(sd)
    model.eval(), again.eval()
    with torch.no_grad():
        assert torch.equal(model(c.batch), again(c.batch))


def test_plugin_rejects_the_unrunnable_default_dims_and_cpu(built_lib):
    from pytorch_news_recommender_b200._lib import NrmsError
    from pytorch_news_recommender_b200.model import nrms as plugin
    c = BertCase("bert_tiny")
    cfg = _config_from_case(c)
    cfg.news_feature_size = c.E + 16            # the shipped defaults (800 vs 512) do not compose either
    with pytest.raises(ValueError):
        plugin.Model(cfg)
    cfg = _config_from_case(c)
    cfg.device = torch.device("cpu")
    with pytest.raises(NrmsError):
        plugin.Model(cfg)(c.batch)
