"""CPU test of the bench contract's reference arm: `bench.py --impl reference` prints ONE JSON line
with the keys the driver reads, times the reference step (the staged unmodified model file, else the oracle port) on the host cores
and launches nothing on a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_impressions_per_sec" and d["unit"] == "impressions/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    # the unmodified reference file when oracle/_ref is staged (build() does it where /root/reference exists),
    # else the oracle port
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "nrms_v0.py"))
    assert cb["kind"] == ("reference" if staged else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "train_eval.py" in cb["sample"]
    assert d["config"]["workload"].startswith("cfg2")
    # both arms build `config` with bench.line_config: the reference arm's object is our arm's, key for key
    assert d["config"] == {**d["config"], "global_batch": 64, "parallelism": "dp1", "gemm_mode": 1,
                           "tokens": "uniform", "table_sync": "dense"}
    assert set(d["config"]) == {"workload", "global_batch", "parallelism", "gemm_mode", "tokens", "table_sync", "l2"}
    assert cb["sample_batch"] >= 8
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_on_other_ranks_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
