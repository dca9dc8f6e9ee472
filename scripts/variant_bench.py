#!/usr/bin/env python
"""Training-step throughput of the `nrms` sibling plugin (reference model/nrms.py).

    python scripts/variant_bench.py [batch=64] [steps=50] [warmup=5]

Workload: cfg2's impression shape on the variant's model — batch 64, 50-slot histories (1..50 real), 1 + 4
candidates, 65,000 news with 512-d vectors, 8 heads, 400-wide additive head, dropout 0.2, dense Adam over
every parameter including the 133 MB vector table.  Synthetic, seeded; there is no published number for this
model, so `vs_baseline` is null and the arms below are the comparison:

  value         this plugin, the reference's loop statements (train_eval.py:189-205) with `DeviceAdam`;
                ids start in pinned host memory, H2D and the loss read-back inside the timed region
  torch_adam    the same with `torch.optim.Adam` (what a user gets by only swapping the model import)
  eager_gpu     the UNMODIFIED reference module (oracle/_ref/nrms.py) in torch eager on the same GPU
  cpu_baseline  the same module on this box's host cores

Prints one JSON line.  The step is launch-bound at this size (a few dozen small kernels), so alongside
impressions/s the line carries the per-kernel breakdown of one step from the library's event profiler.
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pytorch_news_recommender_b200 import _lib, synthetic as S  # noqa: E402
from pytorch_news_recommender_b200.config import Config  # noqa: E402
from pytorch_news_recommender_b200.engine import DeviceAdam  # noqa: E402
from pytorch_news_recommender_b200.model import nrms as plugin  # noqa: E402

W = dict(history_len=50, n_neg=4, n_news=65000, embed=512, heads=8, query=400, dropout=0.2, lr=1e-3)
KEYS = ("browsed_ids", "candidate_ids", "browsed_mask", "candidate_mask", "browsed_titles", "candidate_titles")


def make_config(tmp, device):
    cfg = Config("NRMS_BERT_BENCH")
    cfg.__nrms__()
    cfg.history_len, cfg.sample_size = W["history_len"], W["n_neg"]
    cfg.bert_embed_size = cfg.news_feature_size = W["embed"]
    cfg.user_heads_num, cfg.query_vector_dim_large = W["heads"], W["query"]
    cfg.dropout, cfg.learning_rate = W["dropout"], W["lr"]
    path = os.path.join(tmp, "bert.npz")
    if not os.path.exists(path):
        S.save_embedding_npz(path, S.make_news_vector_table(W["n_news"], W["embed"], seed=0))
    cfg.data_path, cfg.bert_embedding_pretrained = tmp + "/", "bert.npz"
    cfg.device = device
    return cfg


def make_batches(n, batch):
    pool = S.make_news_pool(W["n_news"], 4, 10, seed=0)
    out = []
    for i in range(n):
        b = S.make_train_batch(pool, batch, W["history_len"], W["n_neg"], seed=100 + i, min_hist=1)
        out.append({k: b[k] for k in KEYS})
    return out


def run_loop(model, optimizer, batches, steps, warmup, device):
    """The reference's statements, timed with CUDA events around exactly `steps` iterations."""
    criterion = torch.nn.CrossEntropyLoss()
    model.train()

    def one(datas):
        outputs = model(datas)
        model.zero_grad()
        y = torch.zeros(len(outputs)).long().to(outputs.device)
        loss = criterion(outputs, y)
        value = loss.item()
        loss.backward()
        optimizer.step()
        return value

    for i in range(warmup):
        one(batches[i % len(batches)])
    if device.type == "cuda":
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
    t0 = time.perf_counter()
    losses = [one(batches[i % len(batches)]) for i in range(steps)]
    if device.type == "cuda":
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 1e3 / steps, losses
    return (time.perf_counter() - t0) / steps, losses


def reference_arm(tmp, device, batches, steps, warmup):
    from oracle import ref_runner as R
    ref = R.load_variant()
    cfg = make_config(tmp, device)
    torch.manual_seed(42)
    model = ref.Model(cfg).to(device)
    opt = torch.optim.Adam(model.parameters(), lr=W["lr"])
    return run_loop(model, opt, batches, steps, warmup, device)


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    tmp = os.path.join(tempfile.gettempdir(), "nrms_variant_bench")
    os.makedirs(tmp, exist_ok=True)
    lib = _lib.load()
    host = [{k: v.pin_memory() for k, v in b.items()} for b in make_batches(8, batch)]
    h2d = sum(host[0][k].numel() * host[0][k].element_size() for k in KEYS[:4])
    out = {"metric": "train_impressions_per_sec", "unit": "impressions/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
           "higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "f4: nrms sibling plugin (BERT-vector news encoder + masked 8-head user encoder), "
                                  "batch %d, H=50, 1+4 candidates, 65k news x 512, dense Adam" % batch}}
    results = {}
    for name, make_opt in (("device_adam", lambda m: DeviceAdam(m.parameters(), lr=W["lr"])),
                           ("torch_adam", lambda m: torch.optim.Adam(m.parameters(), lr=W["lr"]))):
        torch.manual_seed(42)
        model = plugin.Model(make_config(tmp, dev)).to(dev)
        n0 = lib.nrms_launch_count()
        sec, losses = run_loop(model, make_opt(model), host, steps, warmup, dev)
        launches = (lib.nrms_launch_count() - n0) / (steps + warmup)
        results[name] = {"value": batch / sec, "ms_per_step": sec * 1e3, "gpu_launches_per_step": launches,
                         "loss_first": losses[0], "loss_last": losses[-1]}
        if name == "device_adam":
            lib.nrms_profile_enable(1)
            run_loop(model, make_opt(model), host, 10, 0, dev)
            import ctypes
            cbuf = ctypes.create_string_buffer(1 << 16)
            lib.nrms_profile_collect(cbuf, len(cbuf))
            lib.nrms_profile_enable(0)
            br = {}
            for line in cbuf.value.decode().splitlines():
                nm, cnt, ms = line.split()
                br[nm] = {"launches_per_step": int(cnt) / 10, "ms_per_step": float(ms) / 10}
            results[name]["kernels"] = br
            results[name]["kernel_ms_per_step"] = sum(v["ms_per_step"] for v in br.values())
        del model
    out.update(value=results["device_adam"]["value"], ms_per_step=results["device_adam"]["ms_per_step"],
               gpu_launches=results["device_adam"]["gpu_launches_per_step"],
               e2e={"value": results["device_adam"]["value"], "unit": "impressions/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": 4, "note": "ids and masks from pinned host memory, loss.item() every step"},
               arms=results)
    try:
        sec, _ = reference_arm(tmp, dev, host, max(10, steps // 2), 3)
        out["eager_gpu"] = {"value": batch / sec, "ms_per_step": sec * 1e3,
                            "note": "unmodified reference model/nrms.py, torch eager + torch.optim.Adam on this GPU"}
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_b = [{k: v.clone() for k, v in b.items()} for b in host[:2]]
        sec, _ = reference_arm(tmp, torch.device("cpu"), cpu_b, 6, 1)
        out["cpu_baseline"] = {"value": batch / sec, "unit": "impressions/s", "cores": cores, "kind": "reference",
                               "sample": "6 steps of the unmodified reference model/nrms.py + torch.optim.Adam, torch CPU"}
    except FileNotFoundError as e:
        out["eager_gpu"] = {"unavailable": str(e)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
