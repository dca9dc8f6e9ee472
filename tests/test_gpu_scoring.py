"""GPU tests of the scoring path: cached news vectors (BASELINE cfg4) vs the full forward, the
on-device metrics vs the oracle restatement of evaluation.py, and the train_demo / evaluate call
sequence of train_eval.py on synthetic loaders."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import nrms_oracle as O

pytestmark = pytest.mark.gpu


def _setup(T=12, H=9, K=3, D=300, vocab=500, n_news=300, gemm_mode=1, dropout=0.2):
    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200.config import Config
    from pytorch_news_recommender_b200.model import NRMS_V0
    cfg = Config("NRMS_V0_SCORING")
    cfg.__nrms__()
    cfg.n_words_title, cfg.history_len, cfg.sample_size, cfg.dropout = T, H, K, dropout
    cfg.word_embed_size, cfg.gemm_mode = D, gemm_mode
    cfg.max_candidate_size = 40
    tmp = tempfile.mkdtemp()
    S.save_embedding_npz(os.path.join(tmp, "emb.npz"), S.make_embedding_table(vocab, D, 0))
    cfg.data_path, cfg.word_embedding_pretrained, cfg.device = tmp + "/", "emb.npz", torch.device("cuda:0")
    pool = S.make_news_pool(n_news, T, vocab, seed=0)
    torch.manual_seed(42)
    model = NRMS_V0(cfg).to(cfg.device)
    return cfg, pool, model


@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_cached_scores_equal_full_forward(gemm_mode, built_lib):
    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200.scoring import CachedScorer
    cfg, pool, model = _setup(gemm_mode=gemm_mode)
    batch = S.make_train_batch(pool, 17, cfg.history_len, cfg.sample_size, seed=3)
    model.eval()
    with torch.no_grad():
        full = model(batch)
    scorer = CachedScorer(model, torch.from_numpy(pool.title_table()), chunk=128)
    vecs = scorer.build_cache()
    assert vecs.shape == (pool.n_news + 1, cfg.word_embed_size)
    cached = scorer.score(batch["browsed_ids"], batch["candidate_ids"], batch["candidate_mask"])
    real = batch["candidate_mask"].bool().cuda()
    assert torch.equal(cached[~real], full[~real])
    # same arithmetic; only the summation order of the final 300-term dot differs between the by-id
    # scorer (float4 per lane) and the [B,S,D] scorer (one column per lane)
    err = ((cached - full).abs()[real] / full.abs()[real].clamp_min(1e-3)).max().item()
    assert err < 1e-4, err


def test_eval_metrics_match_oracle(built_lib):
    from pytorch_news_recommender_b200 import evaluation, synthetic as S
    from pytorch_news_recommender_b200.scoring import CachedScorer
    cfg, pool, model = _setup()
    imp = S.make_eval_impressions(pool, 257, cfg.history_len, cfg.max_candidate_size, seed=1, mean_candidates=9.0)
    scorer = CachedScorer(model, torch.from_numpy(pool.title_table()))
    res = scorer.evaluate(imp, batch=64)
    # the same scores through the oracle's restatement of evaluation.py + train_eval.py:219-227
    logits = torch.cat([scorer.score(imp["browsed_ids"][i:i + 64], imp["candidate_ids"][i:i + 64],
                                     imp["candidate_mask"][i:i + 64]) for i in range(0, 257, 64)], 0)
    _, want = O.evaluate_scores(logits.cpu().numpy(), imp["y_true"])
    for k, col in (("auc", 0), ("mrr", 1), ("ndcg5", 2), ("ndcg10", 3)):
        assert abs(res[k] - float(np.nanmean(want[:, col]))) < 1e-9, k
    assert res["n_defined"] == 257
    got = evaluation.evaluate_scores(logits, imp["y_true"]).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12, equal_nan=True)
    # oracle scores (CPU restatement of the reference forward) -> metrics within 1e-3 (north_star)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ocfg = O.OracleConfig(cfg.n_words_title, cfg.history_len, cfg.sample_size, cfg.word_embed_size,
                          cfg.num_attention_heads, cfg.query_vector_dim, 0.0, 1e-3)
    tt = torch.from_numpy(pool.title_table())
    ob = {"browsed_titles": tt[imp["browsed_ids"]], "candidate_titles": tt[imp["candidate_ids"]],
          "candidate_mask": imp["candidate_mask"]}
    ref_scores = O.model_forward(sd, ob, ocfg, training=False, per_slot=False).numpy()
    ref = np.nanmean(O.evaluate_scores(ref_scores, imp["y_true"])[1], 0)
    for k, col in (("auc", 0), ("mrr", 1), ("ndcg5", 2), ("ndcg10", 3)):
        assert abs(res[k] - float(ref[col])) < 1e-3, (k, res[k], ref[col])


def test_metric_functions_known_answers(built_lib):
    """SURVEY.md §8c KATs generated from the reference's evaluation.py."""
    from pytorch_news_recommender_b200 import evaluation as E
    y, s = [0, 1, 0, 0, 1], [.1, .9, .3, .2, .25]
    assert abs(E.auc_score(y, s) - 0.8333333333333334) < 1e-12
    assert abs(E.mrr_score(y, s) - 0.6666666666666666) < 1e-12
    assert abs(E.ndcg_score(y, s, 5) - 0.9197207891481876) < 1e-12
    assert abs(E.ndcg_score(y, s, 10) - 0.9197207891481876) < 1e-12
    y, s = [1, 0, 0, 0, 0, 0], [.2, .5, .1, .3, .05, 0]
    assert abs(E.auc_score(y, s) - 0.6) < 1e-12 and abs(E.mrr_score(y, s) - 1 / 3) < 1e-12
    assert abs(E.ndcg_score(y, s, 5) - 0.5) < 1e-12
    assert abs(E.dcg_score(y, s, 10) - O.dcg_score(np.array(y), np.array(s), 10)) < 1e-12
    y, s = [1, 0, 1, 0], [.5, .5, .5, .1]          # ties: midrank AUC, highest-index-first order
    assert abs(E.auc_score(y, s) - 0.75) < 1e-12 and abs(E.mrr_score(y, s) - 2 / 3) < 1e-12
    assert np.isnan(E.auc_score([1, 1, 1], [.3, .2, .1]))


@pytest.mark.parametrize("fused", [True, False])
def test_train_demo_call_sequence(fused, built_lib):
    """run_demo.py:58-61: NRMS_V0(config).to(device); train_demo(config, model, train_iter, dev_iter)."""
    from pytorch_news_recommender_b200 import synthetic as S, train_eval
    cfg, pool, model = _setup(gemm_mode=1 if fused else 0, dropout=0.0)
    cfg.num_epochs, cfg.learning_rate = 2, 2e-3
    train_iter = [S.make_train_batch(pool, 16, cfg.history_len, cfg.sample_size, seed=i % 2) for i in range(6)]
    imp = S.make_eval_impressions(pool, 40, cfg.history_len, cfg.max_candidate_size, seed=2, mean_candidates=8.0)
    tt = torch.from_numpy(pool.title_table())
    dev_iter = [{"browsed_titles": tt[imp["browsed_ids"][i:i + 16]], "candidate_titles": tt[imp["candidate_ids"][i:i + 16]],
                 "candidate_mask": imp["candidate_mask"][i:i + 16]} for i in range(0, 40, 16)]
    lines = []
    auc = train_eval.train_demo(cfg, model, train_iter, dev_iter, y_true=imp["y_true"], fused=fused,
                                log=lambda *a: lines.append(" ".join(map(str, a))))
    assert 0.0 <= auc <= 1.0
    assert train_eval.last_metrics["n_impressions"] == 40
    losses = [float(l.split("Train Loss:")[1].split(",")[0]) for l in lines if "Train Loss" in l]
    assert len(losses) == 1 and losses[0] > 0          # one print per 100 batches (train_eval.py:208)
    assert sum("Epoch [" in l for l in lines) == 2 and sum(l.startswith("AUC:") for l in lines) == 2
    # memorising two batches: the training loss must have dropped below the untrained log(C)
    model.eval()
    with torch.no_grad():
        out = model(train_iter[0])
    final = torch.nn.functional.cross_entropy(out, torch.zeros(16, dtype=torch.long, device=out.device)).item()
    assert final < float(np.log(cfg.sample_size + 1))


def test_train_loop_evaluates_logs_and_checkpoints(tmp_path, built_lib):
    """train_eval.train (reference train_eval.py:34-153): warm-up iterations with the rising learning rate,
    an evaluation every `eval_step` batches and after each epoch, one `log_res` line per evaluation, a
    checkpoint whenever the dev AUC beats AUC_best — and the reference's quirk that the model stays in eval
    mode after the first evaluation."""
    from pytorch_news_recommender_b200 import synthetic as S, train_eval
    cfg, pool, model = _setup(gemm_mode=1, dropout=0.2)
    cfg.num_epochs, cfg.learning_rate, cfg.eval_step = 1, 2e-3, 2
    cfg.warm_up, cfg.warm_up_steps = True, 4
    cfg.save_flag, cfg.save_path, cfg.log_path = True, str(tmp_path / "save") + "/", str(tmp_path / "logs")
    train_iter = [S.make_train_batch(pool, 16, cfg.history_len, cfg.sample_size, seed=i) for i in range(5)]
    imp = S.make_eval_impressions(pool, 40, cfg.history_len, cfg.max_candidate_size, seed=2, mean_candidates=8.0)
    tt = torch.from_numpy(pool.title_table())
    dev_iter = [{"browsed_titles": tt[imp["browsed_ids"][i:i + 16]], "candidate_titles": tt[imp["candidate_ids"][i:i + 16]],
                 "candidate_mask": imp["candidate_mask"][i:i + 16]} for i in range(0, 40, 16)]
    lines = []
    w0 = model.state_dict()["user_encoder.additive_attention.linear.weight"].clone()
    res = train_eval.train(cfg, model, train_iter, dev_iter, y_true=imp["y_true"], log=lambda *a: lines.append(" ".join(map(str, a))))
    assert res["total_batch"] == 5 and len(res["loss_records"]) == 5
    assert sum(l.startswith("AUC:") for l in lines) == 3            # batches 2 and 4, then the end of the epoch
    assert any("warm-up training" in l for l in lines) and any("Warm-up Steps" in l for l in lines)
    log = open(cfg.log_path + "/res.txt").read().splitlines()
    assert len(log) == 3 and "_2_:auc_" in log[0] and "_4_:auc_" in log[1]
    assert log[-1].split("_:auc_")[0].endswith("epoch_0")
    assert not model.training                                       # (sic) left in eval mode, as the reference
    assert not torch.equal(w0, model.state_dict()["user_encoder.additive_attention.linear.weight"])
    ckpts = os.listdir(cfg.save_path) if os.path.isdir(cfg.save_path) else []
    if res["auc_best"] > 0.56:                                      # a checkpoint exists exactly when the AUC beat 0.56
        assert ckpts and all(c.endswith(".ckpt") and cfg.model_name in c for c in ckpts)
        assert train_eval.best_checkpoint(cfg) in ckpts
    else:
        assert not ckpts


def test_cached_scorer_rebuilds_after_a_weight_update(built_lib):
    """the news-vector cache must not outlive the weights it was built from."""
    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200.engine import FusedTrainer
    from pytorch_news_recommender_b200.scoring import CachedScorer
    cfg, pool, model = _setup(gemm_mode=1, dropout=0.0)
    batch = S.make_train_batch(pool, 9, cfg.history_len, cfg.sample_size, seed=3)
    scorer = CachedScorer(model, torch.from_numpy(pool.title_table()))
    before = scorer.score(batch["browsed_ids"], batch["candidate_ids"], batch["candidate_mask"]).clone()
    model.train()
    FusedTrainer(model, lr=1e-2).step(batch)
    model.eval()
    after = scorer.score(batch["browsed_ids"], batch["candidate_ids"], batch["candidate_mask"])
    with torch.no_grad():
        full = model(batch)
    real = batch["candidate_mask"].bool().cuda()
    assert not torch.allclose(before[real], after[real])
    assert ((after - full).abs()[real] / full.abs()[real].clamp_min(1e-3)).max().item() < 1e-4
