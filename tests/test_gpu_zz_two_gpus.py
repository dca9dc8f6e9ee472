"""(Named to sort LAST: under `pytest -x` nothing may hide behind a multi-process launch.)

Data-parallel equivalence on real GPUs as a test: `scripts/dp_check.py` under torchrun,
one rank per GPU over NCCL — three `FusedTrainer.step`s on the ranks' shards of a global batch must give the
gradients (2e-3 of each tensor's norm) and the weights of a single-process run over the whole batch, with
both table exchanges (all-reduce + replicated Adam; reduce-scatter -> Adam on V/G rows -> all-gather).

Needs >= 2 visible GPUs: skipped on a one-GPU box (the builder's runs of the same script at 2 and 8 ranks are
profiles/r02_dp_check_*.json).  The host-side exchange logic is covered on CPU by test_parallel_gloo.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("table_sync", ["dense", "sharded"])
def test_data_parallel_equals_single_process_on_two_gpus(table_sync, built_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (one rank per GPU)")
    env = dict(os.environ, TABLE_SYNC=table_sync)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                          os.path.join(ROOT, "scripts", "dp_check.py")],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-4000:])
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["ok"] and rec["world"] == 2 and rec["table_sync"] == table_sync
    assert rec["grad_err_over_tol_max"] < 1.0
