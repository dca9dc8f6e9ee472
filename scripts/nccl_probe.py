#!/usr/bin/env python
"""The collectives of the data-parallel step, timed alone and inside the step.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/nccl_probe.py [steps=40] [caps=16,8,4]

One process per GPU.  Prints one JSON line on rank 0:
  alone     CUDA-event time (max over ranks, mean of 20 calls after 5 warm-ups, nothing else on the GPU) of
            the step's four collectives at their real sizes: reduce-scatter and all-gather of the padded
            84 MB table gradient / table, the 2.65 MB all-reduce of the dense block, and round 1's 84 MB
            all-reduce for comparison; bus bandwidth by NCCL's convention
  in_step   the cfg2 step (device-resident batches, as bench.py's `value`) with the table exchange on the
            default communicator and on communicators whose CTA count is capped (ncclConfig max_ctas),
            plus the per-kernel event profile of the serialised step for the default one — the collectives
            are torch launches and do not appear there; what appears is what they slow down
The cap answers whether NCCL's SMs are what slows the weight-gradient GEMMs that run beside the exchange.
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from pytorch_news_recommender_b200 import _lib  # noqa: E402
from pytorch_news_recommender_b200.engine import FusedTrainer  # noqa: E402
from pytorch_news_recommender_b200.model import NRMS_V0  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    caps = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "16,8,4").split(",") if c]
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    w = bench.WORKLOAD
    V, D = w["vocab"], w["d_model"]
    rows = (V + world - 1) // world * world

    def max_ms(ms):
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_op(fn, n=20, warm=5):
        for _ in range(warm):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return max_ms(a.elapsed_time(b) / n)

    full = torch.randn(rows * D, device=dev)
    shard = torch.empty(rows * D // world, device=dev)
    flat = torch.randn(662600, device=dev)
    nbytes = full.numel() * 4
    alone = {}
    for name, fn, factor, size in (
            ("reduce_scatter_table_grad", lambda: dist.reduce_scatter_tensor(shard, full), (world - 1) / world, nbytes),
            ("all_gather_table", lambda: dist.all_gather_into_tensor(full, shard), (world - 1) / world, nbytes),
            ("all_reduce_table_grad(round 1)", lambda: dist.all_reduce(full), 2 * (world - 1) / world, nbytes),
            ("all_reduce_dense_block", lambda: dist.all_reduce(flat), 2 * (world - 1) / world, flat.numel() * 4)):
        ms = time_op(fn)
        alone[name] = {"ms": ms, "bytes": size, "busbw_GBs": size * factor / ms / 1e6}
    del full, shard

    tmp = os.path.join(tempfile.gettempdir(), "nrms_bench")
    os.makedirs(tmp, exist_ok=True)
    if rank == 0:
        bench.make_config(tmp, dev, 1)
    dist.barrier()
    lib = _lib.load()
    host_batches = bench.make_batches(4, rank)
    in_step = {}
    for label, cap in [("default", None)] + [("max_ctas_%d" % c, c) for c in caps]:
        group = None
        if cap is not None:
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = cap
                group = dist.new_group(ranks=list(range(world)), pg_options=opts)
            except Exception as e:                                      # noqa: BLE001
                in_step[label] = {"unavailable": repr(e)}
                continue
        cfg = bench.make_config(tmp, dev, 1)
        torch.manual_seed(42)
        model = NRMS_V0(cfg).to(dev)
        model.train()
        trainer = FusedTrainer(model, process_group=group)
        resident = []
        for b in host_batches:
            bufs = trainer.load_batch({k: v.pin_memory() for k, v in b.items()})
            resident.append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in bufs.items()
                             if k not in ("ids_slots", "mask_slots", "slot_free", "slot")})
        for i in range(5):
            trainer.step(resident[i % 4])
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            trainer.step(resident[i % 4])
        b.record()
        torch.cuda.synchronize()
        ms = max_ms(a.elapsed_time(b) / steps)
        rec = {"ms_per_step": ms, "impressions_per_s": world * w["batch_per_gpu"] / ms * 1e3}
        if cap is None:
            # which of OUR kernels run longer while the exchange shares the GPU: per-launch events, overlap ON
            lib.nrms_profile_enable(1)
            for i in range(10):
                trainer.step(resident[i % 4])
            torch.cuda.synchronize()
            import ctypes
            buf = ctypes.create_string_buffer(1 << 16)
            lib.nrms_profile_collect(buf, len(buf))
            lib.nrms_profile_enable(0)
            prof = {}
            for ln in buf.value.decode().splitlines():
                nm, cnt, tms = ln.split()
                prof[nm] = round(float(tms) / 10, 4)
            rec["kernel_ms_per_step_with_overlap"] = {k: v for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:14]}
        in_step[label] = rec
        del trainer, model, resident
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"what": "collectives of the cfg2 data-parallel step", "n_gpus": world, "alone": alone,
                          "in_step": in_step, "nccl": ".".join(str(v) for v in torch.cuda.nccl.version())}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
