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
    # sharded form of the table exchange: reduce-scatter, then all-gather of the (here: unchanged) shards
    V, D = local_shape = grads[O.TABLE_KEY].shape
    vs = (V + world - 1) // world
    full = torch.zeros((vs * world, D))
    full[:V] = grads[O.TABLE_KEY] * scale
    mine = torch.empty((vs, D))
    ex.reduce_scatter(full, mine)
    rebuilt = torch.zeros_like(full)
    rebuilt[rank * vs:(rank + 1) * vs] = mine
    ex.all_gather(rebuilt, rebuilt[rank * vs:(rank + 1) * vs])
    if rank == 0:
        np.savez(os.path.join(out_dir, "dp.npz"), flat=flat.numpy(), table=table.numpy(), loss=(t[0] / t[1]).item(),
                 slowest=slowest, sharded_table=rebuilt[:V].numpy())
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
    # reduce-scatter + all-gather of the padded table gradient rebuilds the same dense sum on every rank
    np.testing.assert_allclose(z["sharded_table"], grads[O.TABLE_KEY].numpy(), rtol=2e-5, atol=1e-7)
