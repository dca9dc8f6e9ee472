#!/usr/bin/env python
"""Data-parallel equivalence on real GPUs (run under torchrun, one rank per GPU):
every rank takes its shard of a global batch through `FusedTrainer.step` (NCCL sum-allreduce of
the gradient buffers; TABLE_SYNC=sharded: reduce-scatter + Adam on V/G rows + all-gather) for a few steps; the resulting
parameters must equal a single-process run over the whole batch (dropout off: the masks are
addressed by local row numbers).  Prints one JSON line on rank 0 and exits non-zero on mismatch.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_check.py
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pytorch_news_recommender_b200 import parallel, synthetic as S  # noqa: E402
from pytorch_news_recommender_b200.config import Config  # noqa: E402
from pytorch_news_recommender_b200.engine import FusedTrainer  # noqa: E402
from pytorch_news_recommender_b200.model import NRMS_V0  # noqa: E402


def build(dev, tmp, vocab):
    cfg = Config("NRMS_V0_DP").__nrms__()
    cfg.n_words_title, cfg.history_len, cfg.sample_size, cfg.dropout = 30, 50, 4, 0.0
    cfg.data_path, cfg.word_embedding_pretrained, cfg.device = tmp + "/", "emb.npz", dev
    torch.manual_seed(42)
    return cfg, NRMS_V0(cfg).to(dev)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    rank, world, _ = parallel.init_from_env(device=dev)
    # a one-rank group per rank (collective: every rank creates all of them): rank 0's single-process
    # reference run must not join the job's collectives (FusedTrainer broadcasts its weights at construction)
    solo = None
    if world > 1:
        for r in range(world):
            g = dist.new_group(ranks=[r])
            if r == rank:
                solo = g
    vocab, B, steps = 5000, 16 * world, 3
    tmp = tempfile.mkdtemp()
    S.save_embedding_npz(os.path.join(tmp, "emb.npz"), S.make_embedding_table(vocab, 300, seed=0))
    pool = S.make_news_pool(2000, 30, vocab, seed=0)
    batches = [S.make_train_batch(pool, B, 50, 4, seed=10 + i) for i in range(steps)]
    cfg, model = build(dev, tmp, vocab)
    model.train()
    tr = FusedTrainer(model, table_sync=os.environ.get("TABLE_SYNC", "auto"))
    # gradients of the FIRST step (before Adam's sign-like update amplifies rounding differences of
    # mathematically-zero gradients such as the W_K bias), then the remaining steps for the weights
    tr.step(parallel.shard_batch(batches[0], rank, world), b_global=B)
    grads = {k: v.clone() for k, v in tr.grads_as_state_dict().items()}
    for b in batches[1:]:
        tr.step(parallel.shard_batch(b, rank, world), b_global=B)
    torch.cuda.synchronize()
    ok, worst_g, worst_p = True, 0.0, 0.0
    if rank == 0:
        # single-process reference on the same device: a fresh model, the whole batch, no exchange
        cfg1, ref = build(dev, tmp, vocab)
        ref.train()
        tr1 = FusedTrainer(ref, process_group=solo)   # world 1: no exchange, this rank alone sees the whole batch
        assert tr1.world == 1
        tr1.step(batches[0], b_global=B)
        g1 = {k: v.clone() for k, v in tr1.grads_as_state_dict().items()}
        for b in batches[1:]:
            tr1.step(b, b_global=B)
        torch.cuda.synchronize()
        for k in g1:
            # 2e-3 of the tensor norm (the parity tests' bound) + the rounding floor of zero gradients
            err = float((grads[k] - g1[k]).norm())
            tol = 2e-3 * float(g1[k].norm()) + 1e-7 * g1[k].numel() ** 0.5
            worst_g = max(worst_g, err / max(tol, 1e-30))
            if err > tol:
                ok = False
                print("GRAD MISMATCH", k, err, tol, file=sys.stderr)
        for (k, a), (_, r) in zip(model.state_dict().items(), ref.state_dict().items()):
            # after 3 Adam steps of lr 1e-3 an element whose gradient is rounding noise may differ by ~3e-3
            d = float((a - r).abs().max())
            worst_p = max(worst_p, d)
            if d > 1e-2:
                ok = False
                print("PARAM MISMATCH", k, d, file=sys.stderr)
        print(json.dumps({"check": "data_parallel_equals_single_process", "world": world, "steps": steps,
                          "table_sync": tr.table_sync,
                          "grad_err_over_tol_max": worst_g, "max_abs_param_diff": worst_p, "ok": ok}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
