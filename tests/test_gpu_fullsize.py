"""GPU parity AT THE BENCHMARKED SIZES (VERDICT r01 "What's weak" 1-2): the fused step of bench.py's
cfg2 workload against the oracle number for number, cfg4-shaped cached scoring + metrics against the
oracle, a Zipf-token step (heavy duplicates in the table gradient), and the bf16 mode (cfg3's per-GPU
shape) inside its stated bound.  The oracle runs each of these in seconds on the host."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import nrms_oracle as O

pytestmark = pytest.mark.gpu


def _build(V=70000, T=30, H=50, K=4, dropout=0.2, gemm_mode=1, max_cand=300):
    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200.config import Config
    from pytorch_news_recommender_b200.model import NRMS_V0
    cfg = Config("NRMS_V0_FULLSIZE").__nrms__()
    cfg.n_words_title, cfg.history_len, cfg.sample_size, cfg.dropout = T, H, K, dropout
    cfg.gemm_mode, cfg.max_candidate_size, cfg.dropout_seed = gemm_mode, max_cand, 77
    tmp = tempfile.mkdtemp()
    S.save_embedding_npz(os.path.join(tmp, "emb.npz"), S.make_embedding_table(V, 300, 0))
    cfg.data_path, cfg.word_embedding_pretrained, cfg.device = tmp + "/", "emb.npz", torch.device("cuda:0")
    torch.manual_seed(42)
    model = NRMS_V0(cfg).to(cfg.device)
    ocfg = O.OracleConfig(T, H, K, 300, 10, 200, dropout, cfg.learning_rate)
    return cfg, model, ocfg


def _masks(cfg, n_titles, T, seed):
    from pytorch_news_recommender_b200 import ops
    n = n_titles * T
    m1 = ops.dropout_mask(seed, ops.DROP_EMBEDDING, cfg.dropout, n, 300, "cuda:0").view(-1, T, 300).cpu()
    m2 = ops.dropout_mask(seed, ops.DROP_CONTEXT, cfg.dropout, n, 300, "cuda:0").view(-1, T, 300).cpu()
    return m1, m2


def _step_vs_oracle(cfg, model, ocfg, batch, logit_tol, grad_tol, loss_tol, floor=1e-2):
    from pytorch_news_recommender_b200.engine import FusedTrainer
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    model.train()
    trainer = FusedTrainer(model)
    loss = trainer.step(batch).item()
    B, C, T = batch["candidate_titles"].shape
    H = batch["browsed_titles"].shape[1]
    masks = _masks(cfg, B * (C + H), T, cfg.dropout_seed + 1) if cfg.dropout > 0 else None
    ref_loss, ref_logits, ref_grads = O.loss_and_grads(sd, batch, ocfg, training=cfg.dropout > 0, masks=masks,
                                                       per_slot=False)
    assert abs(loss - float(ref_loss)) < loss_tol * max(1.0, abs(float(ref_loss))), (loss, float(ref_loss))
    logits = trainer._bufs[(B, C, H, T)]["logits"].cpu()
    real = batch["candidate_mask"].bool()
    assert torch.equal(logits[~real], torch.full_like(logits[~real], -1e9))
    d = (logits - ref_logits.detach()).abs()[real].double()
    r = ref_logits.detach().abs()[real].double()
    assert float(d.max() / r.max()) < logit_tol / 10, float(d.max() / r.max())           # against the score scale
    # element by element, relative to max(|logit|, floor)
    assert float((d / r.clamp_min(floor)).max()) < logit_tol, float((d / r.clamp_min(floor)).max())
    got = trainer.grads_as_state_dict()
    assert len(ref_grads) == 19
    for k, g in ref_grads.items():
        gn, dn = float(g.norm()), float((got[k].cpu() - g).norm())
        assert dn <= grad_tol * gn + 1e-6 * np.sqrt(g.numel()), f"grad {k}: |d|={dn} |g|={gn}"
    return loss


def test_cfg2_fused_step_matches_oracle_at_the_benchmarked_size(built_lib):
    """bench.py's default workload, one whole fused step (train mode, dropout 0.2 through the exported
    Philox masks): loss, the 320 logits at 1e-3, all 19 gradients at 2e-3 of their norm — B=64, T=30,
    H=50, K=4, V=70k (105,600 title tokens, the 84 MB table gradient included)."""
    from pytorch_news_recommender_b200 import synthetic as S
    cfg, model, ocfg = _build()
    pool = S.make_news_pool(65000, 30, 70000, seed=0)
    batch = S.make_train_batch(pool, 64, 50, 4, seed=0)
    _step_vs_oracle(cfg, model, ocfg, batch, logit_tol=1e-3, grad_tol=2e-3, loss_tol=1e-4)


def test_zipf_tokens_step_matches_oracle(built_lib):
    """Zipf(1.0) tokens: the most frequent words occur thousands of times in one step, so the
    deduplicating table-gradient reduction runs over long multi-chunk segments."""
    from pytorch_news_recommender_b200 import synthetic as S
    cfg, model, ocfg = _build()
    pool = S.make_news_pool(65000, 30, 70000, seed=0, zipf=True)
    batch = S.make_train_batch(pool, 32, 50, 4, seed=1)
    counts = torch.bincount(torch.cat([batch["browsed_titles"].reshape(-1), batch["candidate_titles"].reshape(-1)]))
    assert int(counts[1:].max()) > 1000          # a genuinely heavy id
    _step_vs_oracle(cfg, model, ocfg, batch, logit_tol=1e-3, grad_tol=2e-3, loss_tol=1e-4)


def test_bf16_mode_at_cfg3_batch_within_stated_bound(built_lib):
    """gemm_mode 2 (plain bf16 products, BASELINE cfg3) at a cfg3-sized per-GPU batch: the looser
    bound north_star allows when bf16 is used — logits 5e-3 of the score scale and 5e-2 element by
    element relative to max(|logit|, 0.1) (bf16 rounds every product at 4e-3: a logit that is itself a
    small difference of large terms has no relative accuracy in this mode), gradients 5e-2 of the norm."""
    from pytorch_news_recommender_b200 import synthetic as S
    cfg, model, ocfg = _build(gemm_mode=2, dropout=0.0)
    ocfg.dropout = 0.0
    pool = S.make_news_pool(65000, 30, 70000, seed=0)
    batch = S.make_train_batch(pool, 256, 50, 4, seed=2)
    _step_vs_oracle(cfg, model, ocfg, batch, logit_tol=5e-2, grad_tol=5e-2, loss_tol=5e-3, floor=0.1)


def test_cfg4_shaped_cached_scoring_and_metrics_match_oracle(built_lib):
    """BASELINE cfg4's shape — T=30, H=50, 300 padded candidate slots, 1,024 impressions over a 3,000-news
    pool: `CachedScorer` (news vectors encoded once, gathers, user encoder, dot, on-device
    AUC/MRR/nDCG) against the oracle doing the same arithmetic on the host (every unique title through
    the oracle's news encoder, then user encoder + click predictor per impression, then the oracle's
    restatement of evaluation.py): scores 1e-3, the four metrics 1e-3 (north_star)."""
    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200.scoring import CachedScorer
    cfg, model, ocfg = _build(dropout=0.0)
    pool = S.make_news_pool(3000, 30, 70000, seed=5)
    imp = S.make_eval_impressions(pool, 1024, 50, 300, seed=1)
    scorer = CachedScorer(model, torch.from_numpy(pool.title_table()))
    res = scorer.evaluate(imp, batch=512)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        vec = O.news_encoder(torch.from_numpy(pool.title_table()), sd, ocfg, training=False)   # [n_news+1, D]
        user = O.user_encoder(vec[imp["browsed_ids"]], sd, ocfg)                                # [N, D]
        scores = O.click_predictor(vec[imp["candidate_ids"]], user)
        scores = scores.masked_fill(imp["candidate_mask"] == 0, -1e9)
    logits = torch.cat([scorer.score(imp["browsed_ids"][i:i + 512], imp["candidate_ids"][i:i + 512],
                                     imp["candidate_mask"][i:i + 512]) for i in range(0, 1024, 512)], 0).cpu()
    real = imp["candidate_mask"].bool()
    d = (logits - scores).abs()[real].double()
    assert float(d.max() / scores.abs()[real].max()) < 1e-4
    assert float((d / scores.abs()[real].double().clamp_min(1e-2)).max()) < 1e-3
    ref = np.nanmean(O.evaluate_scores(scores.numpy(), imp["y_true"])[1], 0)
    for k, col in (("auc", 0), ("mrr", 1), ("ndcg5", 2), ("ndcg10", 3)):
        assert abs(res[k] - float(ref[col])) < 1e-3, (k, res[k], float(ref[col]))
    assert res["n_defined"] == 1024


def test_zipf_step_is_bitwise_repeatable(built_lib):
    """Heavy-tailed tokens: the most frequent words occur thousands of times in one step, so their rows of
    the table gradient are sums over hundreds of 32-row pieces.  The (word, row) pairs are sorted by a stable
    radix sort and the pieces are combined in piece order (no floating-point atomics, no order handed out by
    an atomic cursor), so the WHOLE step — loss, every dense gradient, every row of the table gradient —
    is bit-identical run to run (round 1 was only repeatable for words with <= 32 occurrences)."""
    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200.engine import FusedTrainer
    pool = S.make_news_pool(65000, 30, 70000, seed=0, zipf=True)
    batch = S.make_train_batch(pool, 32, 50, 4, seed=1)
    counts = torch.bincount(torch.cat([batch["browsed_titles"].reshape(-1), batch["candidate_titles"].reshape(-1)]),
                            minlength=70000)
    assert int(counts[1:].max()) > 1000 and int((counts[1:] > 32).sum()) > 20
    runs = []
    for _ in range(4):
        cfg, model, _ = _build()
        model.train()
        tr = FusedTrainer(model)
        losses = [tr.step(batch).item() for _ in range(2)]
        runs.append((losses, tr.flat_grad.clone(), tr.table_grad.clone(), model.state_dict()[O.TABLE_KEY].clone()))
    for losses, fg, tg, tab in runs[1:]:
        assert losses == runs[0][0] and torch.equal(fg, runs[0][1])
        assert torch.equal(tg, runs[0][2]) and torch.equal(tab, runs[0][3])
