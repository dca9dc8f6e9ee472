"""CPU: host logic of the `nrms` sibling plugin (no kernel runs without a GPU): construction replays the
reference's RNG stream and key names, the product path refuses to compute off the device, the training-loop
selection picks the right step."""
import os
import tempfile

import pytest
import torch

from oracle import nrms_bert_oracle as OB
from _golden import BertCase, check_summary


def _config(c, device="cpu"):
    from pytorch_news_recommender_b200.config import Config
    from pytorch_news_recommender_b200 import synthetic as S
    cfg = Config("NRMS_BERT_CPU")
    cfg.__nrms__()
    cfg.history_len, cfg.sample_size = c.H, c.C - 1
    cfg.bert_embed_size = cfg.news_feature_size = c.E
    cfg.user_heads_num, cfg.query_vector_dim_large = c.heads, c.Q
    tmp = tempfile.mkdtemp()
    S.save_embedding_npz(os.path.join(tmp, "bert.npz"), c.table)
    cfg.data_path, cfg.bert_embedding_pretrained = tmp + "/", "bert.npz"
    cfg.device = torch.device(device)
    return cfg


@pytest.mark.parametrize("name", ["bert_tiny", "bert_mind"])
def test_construction_replays_the_reference_init(name):
    from pytorch_news_recommender_b200.model import nrms as plugin
    c = BertCase(name)
    torch.manual_seed(42)
    model = plugin.Model(_config(c))
    assert list(model.state_dict().keys()) == OB.state_dict_keys()
    check_summary(c, "sd0sum", dict(model.state_dict()), rtol=0.0, atol_frac=0.0, abs_floor=0.0)


def test_no_cpu_path_and_default_dims_are_rejected():
    from pytorch_news_recommender_b200._lib import NrmsError
    from pytorch_news_recommender_b200.config import Config
    from pytorch_news_recommender_b200.engine import DeviceAdam
    from pytorch_news_recommender_b200.model import nrms as plugin
    c = BertCase("bert_tiny")
    model = plugin.Model(_config(c))
    with pytest.raises(NrmsError):
        model(c.batch)
    with pytest.raises(NrmsError):
        model.user_encoder(torch.zeros(2, c.H, c.E), torch.ones(2, c.H, dtype=torch.uint8))
    opt = DeviceAdam(model.parameters(), lr=1e-3)
    for p in model.parameters():
        p.grad = torch.zeros_like(p)
    with pytest.raises(NrmsError):
        opt.step()
    opt.zero_grad()
    assert all(p.grad is None for p in model.parameters())
    cfg = Config("x").__nrms__()          # shipped defaults: news_feature_size 800 != bert_embed_size 512
    with pytest.raises(ValueError):
        plugin.Model(cfg)
    with pytest.raises(AttributeError):
        plugin.Model(Config("y"))         # __nrms__() not called


def test_training_loop_picks_the_step_by_plugin():
    from pytorch_news_recommender_b200 import train_eval
    from pytorch_news_recommender_b200.engine import DeviceAdam
    from pytorch_news_recommender_b200.model import nrms as plugin
    c = BertCase("bert_tiny")
    cfg = _config(c)
    model = plugin.Model(cfg)
    fused, opt = train_eval._pick_step(cfg, model, True)
    assert fused is False and isinstance(opt, DeviceAdam)
    fused, opt = train_eval._pick_step(cfg, model, False)
    assert fused is False and isinstance(opt, torch.optim.Adam)
