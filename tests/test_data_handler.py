"""CPU tests of the batch-assembly host logic: `MyDataset` against batches
produced by the REFERENCE's MyDataset + DataLoader (tests/golden/make_golden_data.py), the packed
id matrices the device batcher uploads, and the on-disk formats `run_demo.py` resolves."""
import json
import os
import pickle
import tempfile

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Cfg:
    mode = 'demo'


def load_case():
    with open(os.path.join(GOLDEN, "batches.json")) as f:
        j = json.load(f)
    cfg = Cfg()
    for k, v in j["config"].items():
        setattr(cfg, k, v)
    titles = {int(k): v for k, v in j["title_dict"].items()}
    absts = {int(k): v for k, v in j["abst_dict"].items()}
    return cfg, j["train"], j["dev"], titles, absts, np.load(os.path.join(GOLDEN, "batches.npz"))


@pytest.mark.parametrize("name,typ", [("train", 0), ("dev", 1)])
def test_mydataset_matches_reference_batches(name, typ):
    from pytorch_news_recommender_b200.data_handler import MyDataset
    cfg, train, dev, titles, absts, z = load_case()
    ds = MyDataset(cfg, train if name == "train" else dev, type=typ, words_infos=(titles, absts))
    dl = DataLoader(dataset=ds, batch_size=cfg.batch_size, num_workers=0, drop_last=False, shuffle=False)
    batches = list(dl)
    keys = [k[len(name) + 1:] for k in z.files if k.startswith(name + ".")]
    assert sorted(keys) == sorted(batches[0].keys())
    for k in keys:
        got = torch.cat([torch.as_tensor(b[k]) for b in batches], 0)
        ref = torch.from_numpy(z[f"{name}.{k}"])
        assert got.dtype == ref.dtype, (k, got.dtype, ref.dtype)
        assert got.shape == ref.shape and torch.equal(got, ref), k


def test_pack_samples_equals_mydataset_ids():
    from pytorch_news_recommender_b200.data_handler import pack_samples, title_table_from_dict
    cfg, train, dev, titles, absts, z = load_case()
    for name, lst, S in (("train", train, cfg.sample_size + 1), ("dev", dev, cfg.max_candidate_size)):
        b, bl, c, cl = pack_samples(lst, cfg.history_len, S)
        assert np.array_equal(b, z[f"{name}.browsed_ids"]) and np.array_equal(c, z[f"{name}.candidate_ids"])
        assert np.array_equal(bl, z[f"{name}.browsed_lens"])
        assert np.array_equal((np.arange(S)[None, :] < cl[:, None]).astype(np.uint8), z[f"{name}.candidate_mask"])
        # title rows by news id (id 0 -> zero title) reproduce the reference's title tensors
        tt = np.concatenate([np.zeros((1, cfg.n_words_title), np.int64), title_table_from_dict(titles, cfg.n_words_title)])
        assert np.array_equal(tt[b], z[f"{name}.browsed_titles"]) and np.array_equal(tt[c], z[f"{name}.candidate_titles"])


def test_too_many_items_is_an_error_like_the_reference():
    from pytorch_news_recommender_b200.data_handler import MyDataset, pack_samples
    cfg, train, dev, titles, absts, z = load_case()
    bad = [[[1], [1], [1], [1, 2, 3, 4, 5], [1] * 5, [1] * 5]]          # 5 candidates, 4 slots (data_handler.py:230)
    with pytest.raises(ValueError):
        MyDataset(cfg, bad, type=0, words_infos=(titles, absts))[0]
    with pytest.raises(ValueError):
        pack_samples(bad, cfg.history_len, cfg.sample_size + 1)
    with pytest.raises(ValueError):
        pack_samples([[list(range(1, 9)), [1] * 8, [1] * 8, [1], [1], [1]]], cfg.history_len, 4)   # 8 clicks, H = 6
    assert pack_samples([], 6, 4)[0].shape == (0, 6)


def test_demo_files_round_trip():
    """write_demo_files -> the literal names run_demo.py opens -> load_dataset / get_Demo_Words_Infos /
    load_y_true read them back; the csv fallback of the title dictionary parses like the reference."""
    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200.config import Config
    from pytorch_news_recommender_b200.data_handler import MyDataset, get_Demo_Words_Infos, load_dataset
    from pytorch_news_recommender_b200.train_eval import load_y_true
    cfg = Config("NRMS_V0_DEMO").__nrms__()
    cfg.mode, cfg.n_words_title, cfg.history_len, cfg.sample_size, cfg.max_candidate_size = 'demo', 7, 6, 3, 12
    cfg.word_embed_size = 16
    tmp = tempfile.mkdtemp() + "/"
    cfg.data_path = tmp
    files = S.write_demo_files(tmp, cfg, n_news=30, vocab=40, n_train=9, n_dev=5, seed=3, n_words_abst=4)
    for f in ("demo_word_embedding.npz", "demo_news_title.pkl", "demo_news_abst.pkl", "idx_small_train.pkl",
              "idx_small_dev.pkl", "small_dev_behaviors.csv"):
        assert os.path.exists(tmp + f), f
    assert np.load(tmp + "demo_word_embedding.npz")["embeddings"].shape == (40, 16)
    train = load_dataset(cfg, 'small_train.pkl', tmp, _type=0)
    assert train == files["train"]
    titles, absts = get_Demo_Words_Infos(cfg)
    assert titles == files["title_dict"] and len(absts) == 30
    y = load_y_true(tmp + 'small_dev_behaviors.csv')
    assert y == files["y_true"] and all(0 < sum(v) < len(v) for v in y)
    cfg.n_words_abst = 4
    item = MyDataset(cfg, train, type=0)[0]
    assert item['candidate_titles'].shape == (4, 7) and item['browsed_absts'].shape == (6, 4)
    with pytest.raises(FileNotFoundError):
        load_dataset(cfg, 'missing.pkl', tmp)
    # csv fallback (data_handler.py:119-135)
    os.remove(tmp + "demo_news_title.pkl"); os.remove(tmp + "demo_news_abst.pkl")
    with open(tmp + "demo_news_words.csv", "w") as f:
        for i in range(30):
            f.write('N%d,"%s","%s"\n' % (i, files["title_dict"][i], files["abst_dict"][i]))
    t2, a2 = get_Demo_Words_Infos(cfg)
    assert t2 == files["title_dict"] and a2 == files["abst_dict"] and os.path.exists(tmp + "demo_news_title.pkl")


def test_rank_known_answers_numpy_restatement():
    """The fixture's ranks come from the reference's `_cal_test`; the stable-argsort restatement the
    GPU kernel is tested against reproduces them (tie-free scores)."""
    z = np.load(os.path.join(GOLDEN, "batches.npz"))
    s, lens, ranks = z["rank.scores"], z["rank.lens"], z["rank.ranks"]
    for i in range(s.shape[0]):
        n = int(lens[i])
        order = np.argsort(-s[i, :n], kind="stable")
        r = np.zeros(n, np.int64)
        r[order] = np.arange(1, n + 1)
        assert np.array_equal(r, ranks[i, :n]) and not ranks[i, n:].any()


def test_checkpoint_name_scheme_and_best_checkpoint(tmp_path):
    """train_eval.py:142 file names and the selection rule of train_eval.py:296-303 (highest AUC in
    the name, only above 0.5, only this model's files)."""
    from pytorch_news_recommender_b200 import train_eval as TE
    from pytorch_news_recommender_b200.config import Config
    cfg = Config("NRMS_V0_DEMO")
    cfg.save_path = str(tmp_path) + "/"
    name = TE.checkpoint_name(cfg, 1200, 0.6789)
    assert name.startswith("T") and name.endswith("_NRMS_V0_DEMO_epoch%d_iter_1200_auc_0.679.ckpt" % cfg.num_epochs)
    assert TE.best_checkpoint(cfg) is None
    for fn in ("T01-01_00.00_NRMS_V0_DEMO_epoch5_iter_10_auc_0.612.ckpt", "T01-01_00.01_NRMS_V0_DEMO_epoch5_iter_20_auc_0.655.ckpt",
               "T01-01_00.02_OTHER_epoch5_iter_30_auc_0.900.ckpt", "T01-01_00.03_NRMS_V0_DEMO_epoch5_iter_40_auc_0.480.ckpt",
               "notes_NRMS_V0_DEMO.txt"):
        (tmp_path / fn).write_bytes(b"")
    assert TE.best_checkpoint(cfg) == "T01-01_00.01_NRMS_V0_DEMO_epoch5_iter_20_auc_0.655.ckpt"


from pytorch_news_recommender_b200 import train_eval as TE  # noqa: E402
from pytorch_news_recommender_b200.config import Config  # noqa: E402


def test_warmup_lr_follows_the_reference_scheduler():
    """Values printed by the reference's own GradualWarmupScheduler(optimizer, multiplier=1,
    total_epoch=500) stepped as in train_eval.py:64-98 (optimizer lr 1e-3; run in the build container
    with /root/reference/MIND_2020/lr_scheduler.py): the rate iteration i trains with."""
    ref = {0: 0.0, 1: 0.0, 2: 2e-06, 3: 4e-06, 100: 0.000198, 250: 0.000498, 500: 0.000998, 501: 0.001, 502: 0.001}
    for i, lr in ref.items():
        assert abs(TE.warmup_lr(1e-3, i, 500) - lr) < 1e-12, i


def test_log_res_appends_the_reference_line_format(tmp_path):
    cfg = Config("NRMS_V0_DEMO").__nrms__()
    cfg.log_path = str(tmp_path / "logs" / "NRMS_V0_DEMO")
    TE.log_res(cfg, 0.6123, 5000)      # called (auc, total_batch) like train_eval.py:139
    TE.log_res(cfg, 0.65, "epoch_0")
    lines = open(cfg.log_path + "/res.txt").read().splitlines()
    assert len(lines) == 2 and lines[0].endswith("_5000_:auc_0.6123") and lines[1].endswith("_epoch_0_:auc_0.65")


def test_test_without_a_checkpoint_raises(tmp_path):
    cfg = Config("NRMS_V0_DEMO").__nrms__()
    cfg.save_path = str(tmp_path) + "/"
    with pytest.raises(FileNotFoundError):
        TE.test(cfg, None, [], None, test_list_nums=[1])
