"""CPU: the oracle restatement vs the fixtures generated from the reference itself."""
import numpy as np
import pytest
import torch

from oracle import nrms_oracle as O
from _golden import Case, check_summary


@pytest.mark.parametrize("name", ["tiny", "mind", "long"])
def test_init_matches_reference(name):
    c = Case(name)
    sd = O.init_state_dict(c.cfg, c.table, seed=42)
    check_summary(c, "sd0sum", sd, rtol=0.0, atol_frac=0.0, abs_floor=0.0)


@pytest.mark.parametrize("name,per_slot", [("tiny", True), ("tiny", False), ("mind", False), ("long", False)])
def test_eval_forward(name, per_slot):
    c = Case(name)
    sd = c.state_dict()
    with torch.no_grad():
        logits, cand, hist, user = O.model_forward(sd, c.batch, c.cfg, per_slot=per_slot, return_parts=True)
    np.testing.assert_allclose(logits.numpy(), c.z["eval/logits"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(cand.numpy(), c.z["eval/cand_vec"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(user.numpy(), c.z["eval/user_vec"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(hist.numpy()[:, ::7], c.z["eval/hist_vec_sample"], rtol=2e-5, atol=2e-6)


def test_per_slot_equals_flat():
    """looping over slots == one flattened encoder call (eval mode)."""
    c = Case("tiny")
    sd = c.state_dict()
    with torch.no_grad():
        a = O.model_forward(sd, c.batch, c.cfg, per_slot=True)
        b = O.model_forward(sd, c.batch, c.cfg, per_slot=False)
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["tiny", "mind"])
def test_eval_grads(name):
    c = Case(name)
    sd = c.state_dict()
    loss, _, grads = O.loss_and_grads(sd, c.batch, c.cfg, training=False, per_slot=False)
    assert abs(float(loss) - float(c.z["evalgrad/loss"])) < 1e-5
    check_summary(c, "evalgrad", grads, rtol=2e-4)
    assert float(grads[O.TABLE_KEY][0].abs().max()) == 0.0      # padding_idx row


def _adam_stable(sd):
    """W_K.bias has a mathematically zero gradient (softmax is invariant to a per-query
    constant), so its computed gradient is rounding noise and Adam turns the noise SIGN into
    +-lr steps: those two tensors are not comparable after an optimizer step, in any
    implementation (the reference included)."""
    return {k: v for k, v in sd.items() if not k.endswith("W_K.bias")}


@pytest.mark.parametrize("name", ["tiny", "mind"])
def test_train_steps_with_reference_masks(name):
    """Two Adam steps in train mode with the dropout masks the reference run used."""
    c = Case(name)
    sd = c.state_dict()
    st = O.adam_init(sd)
    for step in range(2):
        masks = c.masks(step)
        loss, logits, grads = O.loss_and_grads(sd, c.batch, c.cfg, training=True, masks=masks, per_slot=False)
        assert abs(float(loss) - float(c.z[f"train/step{step}/loss"])) < 2e-5
        np.testing.assert_allclose(logits.numpy(), c.z[f"train/step{step}/logits"], rtol=1e-4, atol=2e-5)
        check_summary(c, f"train/step{step}/grad", grads, rtol=5e-4)
        O.adam_step(sd, grads, st, c.cfg.learning_rate)
        if step == 0:
            # after step 1 Adam moves every touched weight by ~lr*sign(g): robust to rounding
            # (entries whose gradient is ~eps-sized move by a noise-dependent fraction of lr)
            check_summary(c, "train/step0/param", _adam_stable(sd), rtol=1e-5, atol_frac=0.0,
                          abs_floor=0.05 * c.cfg.learning_rate)
    if name == "tiny":
        for k in _adam_stable(sd):
            np.testing.assert_allclose(sd[k].numpy(), c.z[f"train/final/{k}"], rtol=1e-4,
                                       atol=0.1 * c.cfg.learning_rate)


def test_metrics_vs_reference():
    z = np.load(__import__("os").path.join(__import__("_golden").GOLDEN_DIR, "metrics.npz"))
    off = z["offsets"]
    for i in range(len(off) - 1):
        y = z["labels"][off[i]:off[i + 1]].astype(np.int64)
        s = z["scores"][off[i]:off[i + 1]]
        got = O.impression_metrics(y, s)
        # With tied scores the reference's order is whatever np.argsort's (unstable, SIMD)
        # quicksort yields on the host CPU, so MRR/nDCG are only defined tie-free; AUC
        # (midrank) is order independent and is always compared.
        cols = slice(0, 4) if len(np.unique(s)) == len(s) else slice(0, 1)
        np.testing.assert_allclose(got[cols], z["expected"][i][cols], rtol=1e-12, atol=1e-12, equal_nan=True)


def test_metric_kats():
    """SURVEY §8c known answers."""
    kat = [([0, 1, 0, 0, 1], [.1, .9, .3, .2, .25], 0.8333333333333334, 0.6666666666666666, 0.9197207891481876, 0.9197207891481876),
           ([1, 0, 0, 0, 0, 0], [.2, .5, .1, .3, .05, 0], 0.6, 0.3333333333333333, 0.5, 0.5),
           ([1, 0, 1, 0], [.5, .5, .5, .1], 0.75, 0.6666666666666666, 0.9197207891481876, None)]
    for y, s, auc, mrr, n5, n10 in kat:
        got = O.impression_metrics(y, s)
        assert abs(got[0] - auc) < 1e-12 and abs(got[1] - mrr) < 1e-12 and abs(got[2] - n5) < 1e-12
        if n10 is not None:
            assert abs(got[3] - n10) < 1e-12
    assert np.isnan(O.impression_metrics([0, 0, 0], [.1, .2, .3])).all()


# ---- the `nrms` sibling variant: oracle/nrms_bert_oracle.py vs reference model/nrms.py ----
from oracle import nrms_bert_oracle as OB  # noqa: E402
from _golden import BertCase  # noqa: E402


@pytest.mark.parametrize("name", ["bert_tiny", "bert_mind"])
def test_bert_init_and_eval_forward(name):
    c = BertCase(name)
    check_summary(c, "sd0sum", OB.init_state_dict(c.cfg, c.table, seed=42), rtol=0.0, atol_frac=0.0, abs_floor=0.0)
    sd = c.state_dict()
    with torch.no_grad():
        logits, parts = OB.model_forward(sd, c.batch, c.cfg, return_parts=True)
    np.testing.assert_allclose(logits.numpy(), c.z["eval/logits"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(parts["user_vec"].numpy(), c.z["eval/user_vec"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(parts["cand_vec"].numpy(), c.z["eval/cand_vec"], rtol=2e-5, atol=2e-6)
    assert (logits.numpy()[c.batch["candidate_mask"].numpy() == 0] == -1e9).all()


@pytest.mark.parametrize("name", ["bert_tiny", "bert_mind"])
def test_bert_grads_eval_and_train_masks(name):
    c = BertCase(name)
    sd = c.state_dict()
    loss, _, grads = OB.loss_and_grads(sd, c.batch, c.cfg)
    assert abs(float(loss) - float(c.z["evalgrad/loss"])) < 1e-5
    check_summary(c, "evalgrad", grads, rtol=2e-4)
    # no padding_idx in this variant (nrms.py:222-224), but every use of the pad id is masked: padded
    # history slots get probability exp(-1e9 - max) == 0.0f as keys and as pooled slots, padded candidates
    # are overwritten by -1e9 — so row 0's gradient is exactly zero anyway
    assert float(grads[OB.TABLE_KEY][0].abs().max()) == 0.0
    loss, logits, grads = OB.loss_and_grads(sd, c.batch, c.cfg, c.mults(0))
    assert abs(float(loss) - float(c.z["train/step0/loss"])) < 2e-5
    np.testing.assert_allclose(logits.numpy(), c.z["train/step0/logits"], rtol=1e-4, atol=2e-5)
    check_summary(c, "train/step0/grad", grads, rtol=5e-4)
