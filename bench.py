#!/usr/bin/env python
"""bench.py — NRMS train-step throughput (train impressions/s) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps K --warmup W      # CPU arm

Workload (BASELINE.json configs[1], "cfg2"): NRMS training, fp32, batch 64 impressions per
GPU, synthetic MIND-small-shaped data — title 30 tokens, history 50, 1 positive + 4 negatives,
300-d random embeddings over a 70k-word vocabulary, dropout 0.2, Adam lr 1e-3; one step =
forward + cross-entropy + backward + dense Adam on all 21.66 M parameters
(reference train_eval.py:189-205).  Weak scaling: every rank runs its own 64-impression shard
and gradients are all-reduced over NCCL.

`value`  : device-resident batches (4 distinct batches rotated), CUDA events around exactly K
           steps, barrier + synchronize on both sides, max over ranks.
`e2e`    : the same step through the public training API with HOST batches: every step
           `FusedTrainer.step(host_batch)` (pinned host -> device copy of that step's inputs inside
           the timed region; the NEXT batch's copy is announced with `prefetch` and overlaps the
           step), and the step's loss read back to the host (`last_loss()`: a 4-byte D2H copy).
`roofline`: the dominant kernel of the step, timed live with CUDA events on its own stream
           by the library's opt-in profiler in a separate profiled pass of the same steps.
`cpu_baseline`: the oracle port of the reference step (oracle/nrms_oracle.py, per-slot loops
           as reference nrms_v0.py:255-260, torch CPU, all host cores) timed on this box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOAD = dict(batch_per_gpu=64, n_words_title=30, history_len=50, n_neg=4, vocab=70000,
                d_model=300, n_heads=10, d_query=200, dropout=0.2, n_news=65000)
# BASELINE.json configs (cfg4, cached-vector scoring, is scripts/eval_bench.py).  cfg2 is the default
# line the driver records; the others are selected with --config and name themselves in
# config.workload.  cfg5's per-GPU batch is not fixed by BASELINE.json: 128 keeps its activations
# (5.1 k token rows per impression) at ~20 GB per GPU.
CONFIGS = {
    "cfg2": dict(batch_per_gpu=64, n_words_title=30, history_len=50, n_neg=4, gemm_mode=1,
                 label="cfg2: NRMS training fp32 on 1xB200, batch 64"),
    "cfg3": dict(batch_per_gpu=512, n_words_title=30, history_len=50, n_neg=4, gemm_mode=2,
                 label="cfg3: NRMS training bf16 data-parallel, batch 512/GPU, 70k-word vocab"),
    "cfg5": dict(batch_per_gpu=128, n_words_title=48, history_len=200, n_neg=8, gemm_mode=2,
                 label="cfg5: long-history stress, history 200, title 48, 8 negatives (bf16 products, batch 128/GPU)"),
}
METRIC = "train_impressions_per_sec"
UNIT = "impressions/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), tflops=float(d["bf16_tflops"]),
                    tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured")
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source="fallback")


def make_config(tmp, device, gemm_mode):
    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200.config import Config
    w = WORKLOAD
    cfg = Config("NRMS_V0_BENCH")
    cfg.__nrms__()
    cfg.n_words_title, cfg.history_len, cfg.sample_size = w["n_words_title"], w["history_len"], w["n_neg"]
    cfg.word_embed_size, cfg.num_attention_heads, cfg.query_vector_dim = w["d_model"], w["n_heads"], w["d_query"]
    # (NRMS_BENCH_DROPOUT: experiment knob, e.g. 0 to see what the keep-bit generation costs; the workload's
    # value is what every reported line uses)
    w["dropout"] = float(os.environ.get("NRMS_BENCH_DROPOUT", w["dropout"]))
    cfg.dropout, cfg.learning_rate, cfg.batch_size = w["dropout"], 1e-3, w["batch_per_gpu"]
    cfg.gemm_mode = gemm_mode
    path = os.path.join(tmp, "emb.npz")
    if not os.path.exists(path):
        S.save_embedding_npz(path, S.make_embedding_table(w["vocab"], w["d_model"], seed=0))
    cfg.data_path, cfg.word_embedding_pretrained, cfg.device = tmp + "/", "emb.npz", device
    return cfg


def make_batches(n, rank, zipf=False):
    from pytorch_news_recommender_b200 import synthetic as S
    w = WORKLOAD
    pool = S.make_news_pool(w["n_news"], w["n_words_title"], w["vocab"], seed=0, zipf=zipf)
    return [S.make_train_batch(pool, w["batch_per_gpu"], w["history_len"], w["n_neg"], seed=1000 * rank + i)
            for i in range(n)]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# reference / CPU arm
# ------------------------------------------------------------------------------------------
def cpu_step_time(batch_size, steps, warmup, budget_s, device="cpu"):
    """Times the oracle port of the reference train step (per-slot encoder loops, dropout from
    torch's RNG, autograd backward, Adam) on the host cores.  Returns (impressions/s, info).
    device="cuda:0" runs the same torch-eager port on the GPU instead — what the reference itself
    does when it is given a CUDA device (`--eager-gpu-baseline`; SURVEY.md §8d "the real bar")."""
    from oracle import nrms_oracle as O
    from pytorch_news_recommender_b200 import synthetic as S
    w = WORKLOAD
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    ocfg = O.OracleConfig(w["n_words_title"], w["history_len"], w["n_neg"], w["d_model"], w["n_heads"],
                          w["d_query"], w["dropout"], 1e-3)
    sd = O.init_state_dict(ocfg, S.make_embedding_table(w["vocab"], w["d_model"], seed=0), seed=42)
    on_gpu = str(device) != "cpu"
    if on_gpu:
        sd = {k: v.to(device) for k, v in sd.items()}
    st = O.adam_init(sd)
    pool = S.make_news_pool(w["n_news"], w["n_words_title"], w["vocab"], seed=0)

    def run(bs, i):
        batch = S.make_train_batch(pool, bs, w["history_len"], w["n_neg"], seed=i)
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if on_gpu:   # the model owns the H2D copy in the reference too (nrms_v0.py:248-250,272)
            batch = {k: v.to(device) for k, v in batch.items()}
        O.train_step(sd, st, batch, ocfg, training=True, per_slot=True)   # float(loss): D2H sync
        if on_gpu:
            torch.cuda.synchronize()
        return time.perf_counter() - t0

    bs = batch_size
    t_first = run(bs, 0)
    done_warm = 1
    # bound the whole run: shrink the per-step sample if K+W full batches would not fit
    est = t_first * (steps + max(warmup - 1, 0))
    if est > budget_s:
        bs = max(8, int(batch_size * budget_s / est))
    for i in range(done_warm, warmup):
        run(bs, i)
    ts = [run(bs, 100 + i) for i in range(steps)]
    t = float(np.mean(ts))
    info = {"cores": cores, "threads": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} timed train steps of {bs} impressions (per-slot oracle port of "
                      f"train_eval.py:189-205, torch {'eager on ' + str(device) if on_gpu else 'CPU'} fp32, "
                      f"dropout {w['dropout']}, dense Adam over V={w['vocab']})",
            "batch": bs, "s_per_step": t}
    return bs / t, info


def reference_step_time(batch_size, steps, warmup, budget_s, device="cpu"):
    """Times the UNMODIFIED reference model file (oracle/_ref/nrms_v0.py, staged by
    oracle/stage_ref.py) stepped with the literal statements of train_eval.py:189-205
    (oracle/ref_runner.py) — on the host cores (device="cpu": `torch.device('cuda')` answered with
    the CPU device, the only injection) or as torch eager on the GPU (the file exactly as it is).
    Returns (impressions/s, info), or None when the reference is not staged."""
    from oracle import ref_runner as R
    if not R.available():
        return None
    from pytorch_news_recommender_b200 import synthetic as S
    w = WORKLOAD
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tmp = os.path.join(tempfile.gettempdir(), "nrms_bench")
    os.makedirs(tmp, exist_ok=True)
    path = os.path.join(tmp, "emb.npz")
    if not os.path.exists(path):
        S.save_embedding_npz(path, S.make_embedding_table(w["vocab"], w["d_model"], seed=0))
    dev = torch.device(device)
    on_gpu = dev.type == "cuda"
    rcfg = R.RefConfig(tmp + "/", "emb.npz", w["d_model"], w["n_heads"], w["d_query"], w["dropout"], dev)
    tr = R.ReferenceTrainer(rcfg, 1e-3, seed=42)
    pool = S.make_news_pool(w["n_news"], w["n_words_title"], w["vocab"], seed=0)

    def run(bs, i):
        batch = S.make_train_batch(pool, bs, w["history_len"], w["n_neg"], seed=i)
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        tr.train_step(batch)          # the model owns the H2D copy (nrms_v0.py:248-250,272); loss.item() syncs
        if on_gpu:
            torch.cuda.synchronize()
        return time.perf_counter() - t0

    bs = batch_size
    t_first = run(bs, 0)
    est = t_first * (steps + max(warmup - 1, 0))
    if est > budget_s:
        bs = max(8, int(batch_size * budget_s / est))
    for i in range(1, warmup):
        run(bs, i)
    ts = [run(bs, 100 + i) for i in range(steps)]
    t = float(np.mean(ts))
    info = {"cores": cores, "threads": torch.get_num_threads(), "kind": "reference",
            "sample": f"{steps} timed train steps of {bs} impressions: the unmodified reference model/nrms_v0.py "
                      f"(oracle/_ref, sha256-checked) + Adam + CrossEntropyLoss stepped as train_eval.py:189-205, "
                      f"torch {'eager on ' + str(device) if on_gpu else 'CPU, ' + str(cores) + ' threads'}, fp32, "
                      f"dropout {w['dropout']}, dense Adam over V={w['vocab']}",
            "batch": bs, "s_per_step": t}
    return bs / t, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    got = None
    try:
        got = reference_step_time(WORKLOAD["batch_per_gpu"], args.steps, max(args.warmup, 1), budget_s=200.0)
    except Exception as e:      # a broken staging must not lose the arm: fall back to the port, and say so
        print(f"[bench] staged reference failed ({e!r}); timing the oracle port instead", file=sys.stderr)
    if got is None:
        got = cpu_step_time(WORKLOAD["batch_per_gpu"], args.steps, max(args.warmup, 1), budget_s=200.0)
    val, info = got
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["s_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        # the SAME config object as our arm prints (the driver compares the two); what this arm sampled of it
        # is stated in cpu_baseline
        "config": line_config(args, args.gpus, resolve_table_sync(args, args.gpus)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                         "sample": info["sample"], "sample_batch": info["batch"]},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def algorithmic_counts():
    """SURVEY §8(d) per-unit figures for the workload (per impression / per step per GPU)."""
    w = WORKLOAD
    T, H, D, Qd, C = w["n_words_title"], w["history_len"], w["d_model"], w["d_query"], w["n_neg"] + 1
    N = H + C

    def enc(L):
        return 6 * L * D * D + 4 * L * L * D + 2 * L * D * Qd + 2 * L * Qd + 2 * L * D
    fwd = N * enc(T) + enc(H) + 2 * C * D
    bytes_train = N * T * 8 + 2 * N * T * D * 4 + 2 * N * T * D * 4
    adam_bytes = 7 * 4 * (w["vocab"] * D + 662600)
    return dict(flop_fwd_per_impr=fwd, flop_train_per_impr=3 * fwd, bytes_train_per_impr=bytes_train,
                adam_bytes_per_step=adam_bytes, titles_per_impr=N)


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` per step, from the committed ncu
    capture of this workload (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py); None
    when the capture has no entry for it."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    rec = json.load(open(p)).get("kernels", {}).get(kernel)
    return rec["dram_bytes_per_step"] if rec else None


KERNEL_WORK = {
    # kernel name -> (bound, algorithmic work per LAUNCH for n_seq sequences of length L)
    # flops for GEMMs (2*M*N*K), bytes for HBM-bound kernels
}


def workload_label(args):
    """The same string for both arms (the driver compares them)."""
    w = WORKLOAD
    base = CONFIGS[args.config]["label"] if not args.custom else "custom"
    return (f"{base}: NRMS train step (fwd + CE + bwd + dense Adam), batch {w['batch_per_gpu']} per GPU, "
            f"T={w['n_words_title']} H={w['history_len']} K={w['n_neg']} D={w['d_model']} heads={w['n_heads']} "
            f"Q={w['d_query']} V={w['vocab'] // 1000}k, dropout {w['dropout']}")


def resolve_table_sync(args, world):
    """engine.FusedTrainer's resolution of table_sync='auto' (engine.py: sharded for world > 1)."""
    if args.table_sync != "auto":
        return args.table_sync
    return "sharded" if world > 1 else "dense"


def line_config(args, world, table_sync):
    """`config` of the JSON line — one function for both arms, so the two objects are equal key for key."""
    return {"workload": workload_label(args),
            "global_batch": world * WORKLOAD["batch_per_gpu"], "parallelism": f"dp{world}",
            "gemm_mode": args.gemm_mode, "tokens": "zipf" if args.zipf else "uniform",
            "table_sync": table_sync,
            "l2": "per-step working set (>= 1.5 GB activations + 607 MB Adam state) exceeds the 126 MB L2; "
                  "4 distinct batches rotated"}


def kernel_work(name, B, gemm_mode=1):
    w = WORKLOAD
    e = 2.0 if gemm_mode == 2 else 4.0      # bytes per activation element (bf16 plane / fp32 or hi+lo pair)
    T, H, D, Qd, C, V = w["n_words_title"], w["history_len"], w["d_model"], w["d_query"], w["n_neg"] + 1, w["vocab"]
    M_news, M_user = B * (H + C) * T, B * H
    # the profiler aggregates news + user launches under one name: work is their sum
    M = M_news + M_user
    table = {
        "gemm_fwd_qkv": ("tensor", 2.0 * M * 3 * D * D),
        "gemm_fwd_additive": ("tensor", 2.0 * M * Qd * D),
        "gemm_dgrad_qkv": ("tensor", 2.0 * M * 3 * D * D),
        "gemm_dgrad_additive": ("tensor", 2.0 * M * Qd * D),
        "gemm_wgrad_qkv": ("tensor", 2.0 * M * 3 * D * D),
        "gemm_wgrad_additive": ("tensor", 2.0 * M * Qd * D),
        # bytes per token row (4 B per element: fp32, or a split-bf16 hi + lo pair): fwd reads Q|K|V
        # (3D), writes the context image (D); bwd reads Q|K|V + d_ctx (4D), writes the dQ|dK|dV image (3D)
        "attn_fwd": ("hbm", e * M * (3 * D + D)),
        "attn_bwd": ("hbm", e * M * 3 * D + 4.0 * M * D + e * M * 3 * D),
        "adam": ("hbm", 7.0 * 4 * (V * D + 662600)),
        "pool_fwd": ("hbm", M * (e * D + 4.0)),
        "pool_bwd": ("hbm", M * (e * D + 4.0 * Qd + e * Qd)),
        "gather": ("hbm", M_news * (8.0 + 4.0 * D + e * D)),
        "embgrad_reduce": ("hbm", 4.0 * M_news * D + 4.0 * V * D),
    }
    return table.get(name)


GEMM_LABELS = ("gemm_fwd_qkv", "gemm_fwd_additive", "gemm_dgrad_qkv", "gemm_dgrad_additive",
               "gemm_wgrad_qkv", "gemm_wgrad_additive")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS),
                    help="BASELINE.json config: cfg2 (default, the driver's line), cfg3 (bf16, 512/GPU), cfg5 (long history)")
    ap.add_argument("--gemm-mode", type=int, default=None,
                    help="override the config's GEMM mode: 1 = tcgen05 split-bf16 (fp32-grade), 2 = tcgen05 plain bf16, "
                         "0 = exact-fp32 CUDA-core GEMMs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the eager-on-GPU reference, the loader arm and the drop-in arm (N=1 only anyway)")
    ap.add_argument("--eager-gpu-baseline", action="store_true", help="(kept for compatibility: on by default at N=1)")
    ap.add_argument("--zipf", action="store_true", help="Zipf(1.0) token distribution instead of uniform")
    ap.add_argument("--batch-per-gpu", type=int, default=None, help="override the config's batch per GPU")
    ap.add_argument("--table-sync", default="auto", choices=["auto", "dense", "sharded"],
                    help="exchange of the embedding-table gradient under data parallelism (engine.FusedTrainer)")
    ap.add_argument("--title-len", type=int, default=None)
    ap.add_argument("--history-len", type=int, default=None)
    ap.add_argument("--negatives", type=int, default=None)
    args = ap.parse_args()
    c = CONFIGS[args.config]
    for key in ("batch_per_gpu", "n_words_title", "history_len", "n_neg"):
        WORKLOAD[key] = c[key]
    if args.gemm_mode is None:
        args.gemm_mode = int(os.environ.get("NRMS_GEMM_MODE", c["gemm_mode"]))
    args.custom = False
    for key, val in (("batch_per_gpu", args.batch_per_gpu), ("n_words_title", args.title_len),
                     ("history_len", args.history_len), ("n_neg", args.negatives)):
        if val and val != WORKLOAD[key]:
            WORKLOAD[key] = val
            args.custom = True
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    import __graft_entry__ as entry
    from pytorch_news_recommender_b200 import _lib
    from pytorch_news_recommender_b200.engine import FusedTrainer
    from pytorch_news_recommender_b200.model import NRMS_V0

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    lib = _lib.load()
    warmup = max(args.warmup, 3)
    K = args.steps

    tmp = os.path.join(tempfile.gettempdir(), "nrms_bench")
    os.makedirs(tmp, exist_ok=True)
    if rank == 0:
        cfg = make_config(tmp, device, args.gemm_mode)
    if world > 1:
        dist.barrier()
    cfg = make_config(tmp, device, args.gemm_mode)
    torch.manual_seed(42)
    model = NRMS_V0(cfg).to(device)
    model.train()
    trainer = FusedTrainer(model, table_sync=args.table_sync)
    host_batches = make_batches(4, rank, zipf=args.zipf)
    pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host_batches]
    B = WORKLOAD["batch_per_gpu"]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident arm: each distinct batch lives in its own trainer buffers ----------
    resident = []
    for b in pinned:
        bufs = trainer.load_batch(b)
        resident.append({k: (v.clone() if torch.is_tensor(v) else v) for k, v in bufs.items()
                         if k not in ("ids_slots", "mask_slots", "slot_free", "slot")})
    torch.cuda.synchronize()

    def step_resident(i):
        trainer.step(resident[i % len(resident)])

    # clocks are sampled from the warm-up to the end of the end-to-end region (the device is under
    # load throughout; the timed regions alone can be shorter than one nvidia-smi sample)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(warmup):
        step_resident(i)
    n0 = lib.nrms_launch_count()
    ms = timed(step_resident, K)
    launches = int(lib.nrms_launch_count() - n0)
    ms_per_step = ms / K
    value = world * B * K / (ms / 1e3)

    # ---- end-to-end arm: pinned host batch -> H2D, loss -> host, every step ------------------
    losses = []

    def step_e2e(i):
        trainer.step(pinned[i % len(pinned)])
        trainer.prefetch(pinned[(i + 1) % len(pinned)])   # next batch's H2D runs underneath this step
        losses.append(trainer.last_loss())  # D2H read of the step's result (train_eval.py:198), every step

    for i in range(3):
        step_e2e(i)
    ms_e2e = timed(step_e2e, K)
    e2e_value = world * B * K / (ms_e2e / 1e3)
    h2d = sum(pinned[0][k].numel() * pinned[0][k].element_size()
              for k in ("candidate_titles", "browsed_titles", "candidate_mask"))
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel breakdown (separate profiled pass, CUDA events per launch) ---------------
    roofline, roofline_gemm, breakdown = None, None, {}
    nprof = min(K, 5)
    # per-kernel durations are taken with the step's side-stream work serialised onto one stream: an event
    # pair around a launch that shares the GPU with another stream's kernel would time the sharing, not the
    # kernel.  (`value` above is the overlapped step; the sum of these durations is larger than it.)
    trainer.overlap = False
    step_resident(0)
    sync_all()
    if rank == 0:
        lib.nrms_profile_enable(1)
    for i in range(nprof):          # every rank steps: the step contains the gradient exchange
        step_resident(i)
    sync_all()
    trainer.overlap = True
    if rank == 0:
        peaks = load_peaks()
        import ctypes
        buf = ctypes.create_string_buffer(1 << 16)
        lib.nrms_profile_collect(buf, len(buf))
        lib.nrms_profile_enable(0)
        total = 0.0
        for ln in buf.value.decode().splitlines():
            nm, cnt, tms = ln.split()
            # "<kernel>@user" = the user encoder's launch of a shared kernel: folded into the kernel's line
            # (kernel_work counts both encoders' rows) and listed on its own as well
            base = nm.split("@")[0]
            rec = breakdown.setdefault(base, {"launches_per_step": 0.0, "ms_per_step": 0.0})
            rec["launches_per_step"] += int(cnt) / nprof
            rec["ms_per_step"] += float(tms) / nprof
            if nm != base:
                rec["user_encoder_ms_per_step"] = float(tms) / nprof
            total += float(tms) / nprof

        def roof(name, rec):
            kw = kernel_work(name, B, args.gemm_mode)
            if not kw:
                return None
            bound, work = kw
            sec = rec["ms_per_step"] / 1e3
            if bound == "tensor":
                achieved, peak, unit = work / sec / 1e12, peaks["tflops_sustained"], "TFLOP/s"
            else:
                achieved, peak, unit = work / sec / 1e9, peaks["hbm_gbs"], "GB/s"
            return {"kernel": name, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                    "frac": achieved / peak, "traffic": ncu_traffic(name) if args.config == "cfg2" and not args.custom else None,
                    "peak_source": peaks["source"] + (" sustained bf16 (kernel timed inside the step)" if bound == "tensor"
                                                      else " copy bandwidth"),
                    "share_of_step": rec["ms_per_step"] / total if total else None,
                    "ms_per_launch_group": rec["ms_per_step"]}
        # (1) the largest single kernel of the step ...
        name, rec = max(((k, v) for k, v in breakdown.items() if kernel_work(k, B, args.gemm_mode)),
                        key=lambda kv: kv[1]["ms_per_step"])
        roofline = roof(name, rec)
        # ... (2) and the tcgen05 GEMM family as ONE entry (the same template under six labels would
        # otherwise never be the largest label although it owns half of the step): USEFUL flops of the six
        # contractions (2*M*N*K with the model's own dims, no padding, one product per multiply) over their
        # summed time, against the sustained bf16 peak.  gemm_mode 1 issues three bf16 MMAs per useful
        # product (fp32-grade split), so its ceiling on this scale is 1/3.
        fam = [(nm, breakdown[nm]) for nm in GEMM_LABELS if nm in breakdown]
        if fam:
            flops = sum(kernel_work(nm, B, args.gemm_mode)[1] for nm, _ in fam)
            msf = sum(r["ms_per_step"] for _, r in fam)
            ach = flops / (msf / 1e3) / 1e12
            roofline_gemm = {"kernel": "ig_gemm_kernel family (6 labels)", "bound": "tensor", "achieved": ach,
                             "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tflops_sustained"],
                             "mma_terms_per_product": 1 if args.gemm_mode == 2 else 3,
                             "frac_of_issued_mma_peak": ach * (1 if args.gemm_mode == 2 else 3) / peaks["tflops_sustained"],
                             "share_of_step": msf / total if total else None, "ms_per_step": msf,
                             "peak_source": peaks["source"] + " sustained bf16"}
        for nm in breakdown:
            breakdown[nm]["share"] = breakdown[nm]["ms_per_step"] / total if total else None

    # ---- CPU baseline on this box's host cores, rank 0 at N=1 only: the unmodified reference file when it
    # is staged (oracle/_ref), else the oracle port ---------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            got = None
            try:
                got = reference_step_time(B, steps=2, warmup=1, budget_s=40.0)
            except Exception as e:
                print(f"[bench] staged reference failed ({e!r}); timing the oracle port", file=sys.stderr)
            if got is None:
                got = cpu_step_time(B, steps=2, warmup=1, budget_s=40.0)
            v, info = got
            cpu_baseline = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                            "sample": info["sample"]}
        except Exception as e:  # the GPU numbers stand on their own
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                            "sample": f"failed: {e!r}"}

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        # ---- the reference in torch eager on THIS GPU (SURVEY.md §8d "the real bar"): the unmodified file
        try:
            got = reference_step_time(B, steps=5, warmup=2, budget_s=60.0, device=f"cuda:{local_rank}")
            if got is None:
                got = cpu_step_time(B, steps=5, warmup=2, budget_s=60.0, device=f"cuda:{local_rank}")
            v, info = got
            extras["eager_gpu_baseline"] = {"value": v, "unit": UNIT, "kind": info["kind"],
                                            "ms_per_step": info["s_per_step"] * 1e3, "sample": info["sample"]}
        except Exception as e:
            extras["eager_gpu_baseline"] = {"value": None, "unit": UNIT, "sample": f"failed: {e!r}"}
        # ---- loader arm: DeviceBatcher (batch assembly on the GPU from resident id tables) -> step
        try:
            extras["e2e_loader"] = loader_arm(cfg, trainer, B, K)
        except Exception as e:
            extras["e2e_loader"] = {"value": None, "unit": UNIT, "note": f"failed: {e!r}"}
        # ---- drop-in arm: the literal reference lines over the nn.Module (autograd + torch.optim.Adam)
        try:
            extras["dropin"] = dropin_arm(cfg, pinned, B, min(K, 20))
        except Exception as e:
            extras["dropin"] = {"value": None, "unit": UNIT, "note": f"failed: {e!r}"}

    if rank == 0:
        alg = algorithmic_counts()
        peaks = load_peaks()
        step_flops = alg["flop_train_per_impr"] * B
        step_bytes = alg["bytes_train_per_impr"] * B + alg["adam_bytes_per_step"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.gemm_mode == 2 else "f32", "data": "synthetic",
            "config": line_config(args, world, trainer.table_sync),
            "news_encodes_per_sec": value * alg["titles_per_impr"],
            "step_model": {"algorithmic_flop_per_step_per_gpu": step_flops,
                           "algorithmic_bytes_per_step_per_gpu": step_bytes,
                           "achieved_tflops": step_flops / (ms_per_step / 1e3) / 1e12,
                           "hbm_floor_ms": step_bytes / peaks["hbm_gbs"] / 1e6,
                           "tensor_floor_ms": step_flops / peaks["tflops"] / 1e9},
            "roofline": roofline,
            "roofline_gemm_family": roofline_gemm,
            "kernel_breakdown": breakdown,
            "kernel_breakdown_note": "durations of a SERIALISED step (side streams off); the timed step overlaps the "
                                     "user encoder's weight gradients and the table path (embedding-gradient reduction + "
                                     "Adam / exchange) with the main chain",
            "cpu_baseline": cpu_baseline,
            **extras,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "clocks": clocks,
            "final_loss": losses[-1] if losses else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def loader_arm(cfg, trainer, B, K):
    """`for datas in DeviceBatcher(...): trainer.step(datas)` — the loop train_eval.train_demo runs when
    it is given the GPU loader: per step one batch-assembly launch (news ids -> title tokens, masks;
    data_handler.py:185-250 semantics) and the fused step; the loss is read back every step."""
    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200.data_handler import DeviceBatcher
    w = WORKLOAD
    pool = S.make_news_pool(w["n_news"], w["n_words_title"], w["vocab"], seed=0)
    titles = {i: pool.titles[i].tolist() for i in range(pool.n_news)}
    S_ = w["n_neg"] + 1
    n = B * min(K, 50)
    datas = S.make_sample_lists(pool, n, w["history_len"], S_, S_, seed=0)
    db = DeviceBatcher(cfg, datas, type=0, batch_size=B, shuffle=True, drop_last=True, seed=1, words_infos=(titles, {}))
    it = iter(db)
    for _ in range(3):
        trainer.step(next(it))
        trainer.last_loss()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 0
    e0.record()
    for datas_b in db:
        trainer.step(datas_b)
        trainer.last_loss()
        steps += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"value": B * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
            "note": "DeviceBatcher (one assembly launch per batch from resident id / title tables) -> FusedTrainer.step -> last_loss()"}


def dropin_arm(cfg, pinned, B, K):
    """The literal statements of train_eval.py:189-205 over OUR nn.Module with HOST batches:
    outputs = model(datas); model.zero_grad(); loss = criterion(outputs, zeros); loss.item();
    loss.backward(); torch.optim.Adam.step() — what a user gets by only swapping the model import."""
    import torch.nn as nn
    from pytorch_news_recommender_b200.model import NRMS_V0
    torch.manual_seed(42)
    model = NRMS_V0(cfg).to(cfg.device)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=cfg.learning_rate)
    crit = nn.CrossEntropyLoss()

    def step(i):
        datas = pinned[i % len(pinned)]
        outputs = model(datas)
        model.zero_grad()
        y = torch.zeros(len(outputs)).long().to(outputs.device)
        loss = crit(outputs, y)
        v = loss.item()
        loss.backward()
        opt.step()
        return v

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    del opt, model
    return {"value": B * K / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / K, "steps": K,
            "note": "model(datas) / criterion / loss.backward() / torch.optim.Adam over the drop-in nn.Module "
                    "(autograd Functions over the same kernels; dense table gradient + torch Adam as in the reference)"}


if __name__ == "__main__":
    main()
