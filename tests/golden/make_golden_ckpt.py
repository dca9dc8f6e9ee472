#!/usr/bin/env python
"""A checkpoint written by the REFERENCE model (`torch.save(model.state_dict(), ...)`,
train_eval.py:142) plus the reference's eval logits for a fixed batch, so that the test can
prove this package loads reference checkpoints and scores like the reference.

Tiny dims (D=24, 4 heads, Q=8, vocab 50) keep the fixture at a few KB; the weights are the
reference's seed-42 initialisation after two of its own Adam steps (eval-mode batches, no dropout).

    python tests/golden/make_golden_ckpt.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import make_golden as G  # noqa: E402
from oracle import nrms_oracle as O  # noqa: E402
from pytorch_news_recommender_b200 import synthetic as S  # noqa: E402


def main():
    torch.set_num_threads(2)
    ref, _ = G.load_reference()
    cfg = O.OracleConfig(n_words_title=6, history_len=5, sample_size=2, word_embed_size=24,
                         num_attention_heads=4, query_vector_dim=8, dropout=0.0, learning_rate=1e-2)
    vocab, n_news, B, seed = 50, 40, 6, 21
    table = S.make_embedding_table(vocab, cfg.word_embed_size, seed=seed)
    pool = S.make_news_pool(n_news, cfg.n_words_title, vocab, seed=seed)
    tmp = tempfile.mkdtemp()
    S.save_embedding_npz(os.path.join(tmp, "emb.npz"), table)
    torch.manual_seed(42)
    model = ref.Model(G.RefConfig(cfg, tmp + "/", "emb.npz"))
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate)       # train_eval.py:167
    crit = torch.nn.CrossEntropyLoss()
    model.train()
    for step in range(2):
        batch = S.make_train_batch(pool, B, cfg.history_len, cfg.sample_size, seed=seed + step)
        out = model(batch)
        model.zero_grad()
        loss = crit(out, torch.zeros(len(out)).long())
        loss.backward()
        opt.step()
    torch.save(model.state_dict(), os.path.join(HERE, "ref_tiny.ckpt"))
    model.eval()
    batch = S.make_train_batch(pool, B, cfg.history_len, cfg.sample_size, seed=seed + 9, short_tail=0.4)
    with torch.no_grad():
        logits = model(batch)
    np.savez_compressed(os.path.join(HERE, "ref_tiny_ckpt.npz"),
                        dims=np.array([cfg.n_words_title, cfg.history_len, cfg.sample_size, cfg.word_embed_size,
                                       cfg.num_attention_heads, cfg.query_vector_dim, vocab, n_news, B, seed]),
                        browsed_titles=batch["browsed_titles"].numpy(), candidate_titles=batch["candidate_titles"].numpy(),
                        candidate_mask=batch["candidate_mask"].numpy(), logits=logits.numpy())
    print("ckpt bytes", os.path.getsize(os.path.join(HERE, "ref_tiny.ckpt")), "logits", logits.shape)


if __name__ == "__main__":
    main()
