"""Stages the UNMODIFIED reference model file for the reference arm — TEST / BENCH INFRASTRUCTURE.

    python oracle/stage_ref.py        # needs /root/reference (build container only)

The reference is pure Python with no setup.py / pyproject.toml, so `pip install --target
baseline/_ref /root/reference` is impossible.  What the hot path needs from it is
ONE file, `MIND_2020/model/nrms_v0.py` (imports: torch, numpy).  This recipe copies that file
byte for byte from where it lies under /root/reference into `oracle/_ref/` — a git-ignored,
NOT gpurun-ignored directory, so the copy never enters history but travels to the GPU box like
a built .so — and records its sha256 in `oracle/_ref/MANIFEST.json`.  `oracle/ref_runner.py`
imports the staged file and verifies the hash.  Nothing under `pytorch_news_recommender_b200/`
reads `oracle/_ref/`; only `bench.py`'s reference / eager-baseline legs and `tests/` do.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference/MIND_2020"
DST = os.path.join(HERE, "_ref")
FILES = {"nrms_v0.py": "model/nrms_v0.py", "evaluation.py": "evaluation.py",
         "nrms.py": "model/nrms.py"}          # the sibling variant's module (scripts/variant_bench.py)


def sha256(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(verbose: bool = True) -> bool:
    """Copies the files; returns False (and leaves any earlier staging alone) when the reference
    tree is absent — e.g. on the GPU box, which only uses what was staged here."""
    if not os.path.isdir(REF_ROOT):
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name, rel in FILES.items():
        src = os.path.join(REF_ROOT, rel)
        shutil.copyfile(src, os.path.join(DST, name))
        manifest[name] = {"source": "MIND_2020/" + rel, "sha256": sha256(src), "bytes": os.path.getsize(src)}
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print("[stage_ref] staged", ", ".join(sorted(manifest)), "->", DST, file=sys.stderr)
    return True


if __name__ == "__main__":
    ok = stage()
    sys.exit(0 if ok else 1)
