#!/usr/bin/env python
"""Eval-only scoring throughput (BASELINE cfg4): synthetic impressions with ~37 candidates
(clipped lognormal in [2, 300]) and 50-slot histories over a 65k-news pool, scored from cached
news vectors with AUC / MRR / nDCG@5/10 computed on the device.

    python scripts/eval_bench.py [n_impressions=200000] [batch=8192]
Prints one JSON line (impressions/s for cache build + scoring + metrics, device-resident ids)."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pytorch_news_recommender_b200 import synthetic as S  # noqa: E402
from pytorch_news_recommender_b200.model import NRMS_V0  # noqa: E402
from pytorch_news_recommender_b200.scoring import CachedScorer  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    tmp = os.path.join(tempfile.gettempdir(), "nrms_bench")
    os.makedirs(tmp, exist_ok=True)
    cfg = bench.make_config(tmp, dev, 1)
    w = bench.WORKLOAD
    torch.manual_seed(42)
    model = NRMS_V0(cfg).to(dev).eval()
    pool = S.make_news_pool(w["n_news"], w["n_words_title"], w["vocab"], seed=0)
    imp = S.make_eval_impressions(pool, n, w["history_len"], 300, seed=1)
    # trim the padded candidate axis to the longest impression of this sample (the reference pads
    # to max_candidate_size=300; slots beyond the longest impression are all padding)
    smax = int(imp["n_candidates"].max())
    dimp = {k: (v[:, :smax].contiguous().to(dev) if k in ("candidate_ids", "candidate_mask", "labels") else v.to(dev))
            for k, v in imp.items() if torch.is_tensor(v)}
    scorer = CachedScorer(model, torch.from_numpy(pool.title_table()))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scorer.build_cache()
    torch.cuda.synchronize()
    t_cache = time.perf_counter() - t0
    scorer.evaluate({k: v[:batch] for k, v in dimp.items()}, batch=batch)          # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = scorer.evaluate(dimp, batch=batch)
    torch.cuda.synchronize()
    t_eval = time.perf_counter() - t0
    print(json.dumps({"metric": "eval_impressions_per_sec", "value": n / t_eval, "n_impressions": n,
                      "candidate_slots": smax, "mean_candidates": float(imp["n_candidates"].float().mean()),
                      "cache_build_s": t_cache, "news_encodes_per_sec": (pool.n_news + 1) / t_cache,
                      "eval_s": t_eval, "metrics": res}))


if __name__ == "__main__":
    main()
