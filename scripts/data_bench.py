#!/usr/bin/env python
"""Batch-assembly throughput at the bench workload's shapes (cfg2: batch 64,
H=50, 5 candidate slots, T=30, 65k-news pool): `DeviceBatcher` (three row gathers on the GPU) next
to the host path it replaces (`MyDataset.__getitem__` + default_collate, data_handler.py:185-250,
one process).  HBM-bound integer work: algorithmic bytes per impression =
(H+S)*8 (ids read) * 2 (written) + (H+S)*T*8 * 2 (title rows read + written).

    python scripts/data_bench.py [n_samples=20000] [batch=64]
Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

import bench  # noqa: E402
from pytorch_news_recommender_b200 import synthetic as S  # noqa: E402
from pytorch_news_recommender_b200.config import Config  # noqa: E402
from pytorch_news_recommender_b200.data_handler import DeviceBatcher, MyDataset  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    w = bench.WORKLOAD
    cfg = Config("NRMS_V0_DATA").__nrms__()
    cfg.n_words_title, cfg.history_len, cfg.sample_size = w["n_words_title"], w["history_len"], w["n_neg"]
    cfg.device = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    pool = S.make_news_pool(w["n_news"], cfg.n_words_title, w["vocab"], seed=0)
    titles = {i: pool.titles[i].tolist() for i in range(pool.n_news)}
    S_ = cfg.sample_size + 1
    datas = S.make_sample_lists(pool, n, cfg.history_len, S_, S_, seed=0)
    t0 = time.perf_counter()
    db = DeviceBatcher(cfg, datas, type=0, batch_size=batch, shuffle=True, seed=1, words_infos=(titles, {}))
    torch.cuda.synchronize()
    t_pack = time.perf_counter() - t0
    for _ in db:                                   # warm-up epoch
        pass
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nb = 0
    for b in db:
        nb += 1
    e1.record()
    torch.cuda.synchronize()
    dev_s = e0.elapsed_time(e1) / 1e3
    # host path on a bounded sample (it is ~4 orders of magnitude slower)
    m = min(n, 2048)
    ds = MyDataset(cfg, datas[:m], type=0, words_infos=(titles, {}))
    t0 = time.perf_counter()
    for _ in DataLoader(dataset=ds, batch_size=batch, num_workers=0, shuffle=False):
        pass
    host_s = time.perf_counter() - t0
    H, T = cfg.history_len, cfg.n_words_title
    bytes_per_imp = 2 * (H + S_) * 8 + 2 * (H + S_) * T * 8
    peaks = bench.load_peaks()
    out = {"metric": "assembled_impressions_per_sec", "value": n / dev_s, "n_samples": n, "batch": batch,
           "batches": nb, "ms_per_batch": 1e3 * dev_s / nb, "one_off_pack_s": t_pack,
           "algorithmic_bytes_per_impression": bytes_per_imp, "achieved_GBps": n * bytes_per_imp / dev_s / 1e9,
           "hbm_peak_GBps": peaks["hbm_gbs"], "frac_of_hbm_peak": n * bytes_per_imp / dev_s / 1e9 / peaks["hbm_gbs"],
           "host_mydataset_impressions_per_sec": m / host_s, "host_sample": m}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
