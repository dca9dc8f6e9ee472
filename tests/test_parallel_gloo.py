"""Data-parallel host logic on CPU: world_size 2 over gloo.

What is checked: the impression sharding, and that per-rank gradients of the shard's loss
scaled by B_shard/B_global (i.e. the mean over the GLOBAL batch, as the fused kernels emit
them) SUM-allreduce to exactly the single-process gradient — the invariant that makes every
rank's Adam step identical.  The per-rank gradients come from the oracle (the CUDA kernels
need a GPU); the exchange code under test is the product's `parallel.GradientExchange`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nrms_oracle as O
from pytorch_news_recommender_b200 import parallel
from pytorch_news_recommender_b200 import synthetic as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    cfg = O.OracleConfig(8, 6, 2, 20, 2, 8, 0.0, 1e-3)
    vocab = 50
    table = S.make_embedding_table(vocab, cfg.word_embed_size, seed=3)
    sd = O.init_state_dict(cfg, table, seed=7)
    pool = S.make_news_pool(40, cfg.n_words_title, vocab, seed=1)
    batch = S.make_train_batch(pool, 6, cfg.history_len, cfg.sample_size, seed=5)
    return cfg, sd, batch


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    cfg, sd, batch = _problem()
    B = batch["candidate_titles"].shape[0]
    shard = parallel.shard_batch(batch, rank, world)
    lo, hi = parallel.shard_range(B, rank, world)
    assert shard["candidate_titles"].shape[0] == hi - lo
    loss, _, grads = O.loss_and_grads(sd, shard, cfg, training=False, per_slot=False)
    scale = (hi - lo) / B                       # mean over the shard -> share of the global mean
    keys = sorted(grads)
    flat = torch.cat([grads[k].reshape(-1) for k in keys if k != O.TABLE_KEY]) * scale
    table = grads[O.TABLE_KEY].clone() * scale
    ex = parallel.GradientExchange()
    assert ex.world == world
    ex.allreduce([flat, table])
    t = ex.sum_over_ranks(torch.tensor([float(loss) * (hi - lo), float(hi - lo)], dtype=torch.float64))
    slowest = ex.max_over_ranks(float(rank + 1), "cpu")
    # sparse form of the table exchange: gather the touched rows of every rank, rebuild the dense sum
    local = grads[O.TABLE_KEY].clone() * scale
    ids = torch.nonzero(local.abs().amax(dim=1) > 0).squeeze(1)
    ids_all, rows_all = ex.allgather_rows(ids, local[ids])
    assert ids_all.numel() % world == 0 and rows_all.shape == (ids_all.numel(), local.shape[1])
    sparse_sum = torch.zeros_like(local).index_add_(0, ids_all, rows_all)
    sparse_sum[0] = 0                            # id 0 doubles as the padding of the shorter ranks
    if rank == 0:
        np.savez(os.path.join(out_dir, "dp.npz"), flat=flat.numpy(), table=table.numpy(), loss=(t[0] / t[1]).item(),
                 slowest=slowest, sparse_table=sparse_sum.numpy(), n_gathered=ids_all.numel())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_the_batch():
    for n in (1, 5, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_world2_sum_allreduce_equals_single_process_gradient(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    z = np.load(tmp_path / "dp.npz")
    cfg, sd, batch = _problem()
    loss, _, grads = O.loss_and_grads(sd, batch, cfg, training=False, per_slot=False)
    keys = sorted(grads)
    flat = torch.cat([grads[k].reshape(-1) for k in keys if k != O.TABLE_KEY]).numpy()
    np.testing.assert_allclose(z["flat"], flat, rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(z["table"], grads[O.TABLE_KEY].numpy(), rtol=2e-5, atol=1e-7)
    assert abs(float(z["loss"]) - float(loss)) < 1e-6
    assert float(z["slowest"]) == 2.0
    # the sparse exchange rebuilds the same dense table gradient from fewer rows than the vocabulary
    np.testing.assert_allclose(z["sparse_table"], grads[O.TABLE_KEY].numpy(), rtol=2e-5, atol=1e-7)
    assert 0 < int(z["n_gathered"]) < 2 * grads[O.TABLE_KEY].shape[0]
