#!/usr/bin/env python
"""Runs a few fused NRMS train steps at the bench workload (cfg2) so that `ncu` can capture
every kernel of one step.  Usage (on the GPU box):

    python scripts/prof_step.py                       # plain run, must exit 0 first
    ncu --set full --clock-control none --import-source on --launch-skip <W*L> --launch-count <L> \
        -o gpurun_out/prof_step python scripts/prof_step.py

Env: STEPS (default 3), CONFIG (cfg2 | cfg3 | cfg5: bench.CONFIGS), GEMM_MODE / B override the config's.
Prints the number of library launches per step so --launch-skip can be computed."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from pytorch_news_recommender_b200 import _lib  # noqa: E402
from pytorch_news_recommender_b200.engine import FusedTrainer  # noqa: E402
from pytorch_news_recommender_b200.model import NRMS_V0  # noqa: E402


def main():
    steps = int(os.environ.get("STEPS", "3"))
    conf = dict(bench.CONFIGS[os.environ.get("CONFIG", "cfg2")])
    conf.pop("label")
    gemm_mode = int(os.environ.get("GEMM_MODE", conf.pop("gemm_mode")))
    bench.WORKLOAD.update(conf)
    if "B" in os.environ:
        bench.WORKLOAD["batch_per_gpu"] = int(os.environ["B"])
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    tmp = os.path.join(tempfile.gettempdir(), "nrms_bench")
    os.makedirs(tmp, exist_ok=True)
    cfg = bench.make_config(tmp, dev, gemm_mode)
    torch.manual_seed(42)
    model = NRMS_V0(cfg).to(dev)
    model.train()
    trainer = FusedTrainer(model)
    batch = trainer.load_batch(bench.make_batches(1, 0)[0])
    lib = _lib.load()
    torch.cuda.synchronize()
    for i in range(steps):
        n0 = lib.nrms_launch_count()
        loss = trainer.step(batch)
        torch.cuda.synchronize()
        print(f"step {i}: loss {loss.item():.6f} launches {lib.nrms_launch_count() - n0}", flush=True)
    print("ok")


if __name__ == "__main__":
    main()
