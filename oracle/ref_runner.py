"""Runs the staged, UNMODIFIED reference `model/nrms_v0.py` (oracle/_ref/, see stage_ref.py) —
TEST / BENCH INFRASTRUCTURE, never imported by the product.

`load()` imports the staged file by path after checking its sha256 against the manifest.  On a
CPU-only host the module-global `torch` of the reference module is wrapped so that
`torch.device('cuda')` answers with the CPU device (the reference hard-codes CUDA at
nrms_v0.py:248,250,272); on a GPU the file runs exactly as it is.  `train_step` is the literal
statement sequence of the reference's loop (train_eval.py:189-205); train_eval.py itself cannot be
imported (matplotlib / nltk / tqdm side imports).
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "nrms_v0.py")) and os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


class _TorchProxy(types.ModuleType):
    """`torch` for a host without a GPU: device('cuda*') -> device('cpu'); everything else untouched."""

    def __init__(self, real):
        super().__init__("torch_proxy")
        self.__dict__["_real"] = real

    def __getattr__(self, k):
        return getattr(self.__dict__["_real"], k)

    def device(self, name, *a):
        real = self.__dict__["_real"]
        return real.device("cpu") if str(name).startswith("cuda") else real.device(name, *a)


def load(cpu_proxy: bool):
    if not available():
        raise FileNotFoundError("oracle/_ref is not staged (python oracle/stage_ref.py in the build container)")
    manifest = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))
    path = os.path.join(REF_DIR, "nrms_v0.py")
    got = hashlib.sha256(open(path, "rb").read()).hexdigest()
    if got != manifest["nrms_v0.py"]["sha256"]:
        raise RuntimeError("oracle/_ref/nrms_v0.py does not match its manifest: not the unmodified reference file")
    spec = importlib.util.spec_from_file_location("ref_nrms_v0_staged", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    if cpu_proxy:
        ref.torch = _TorchProxy(torch)
    return ref


def load_variant():
    """The staged, unmodified `model/nrms.py` (the BERT-vector sibling).  It imports `torchsnooper` (absent
    here; every `@snoop()` in the file is commented out) and `tools.log_exec_time` (tools.py pulls in
    matplotlib): both are answered with empty stand-ins for the duration of the import."""
    import sys
    manifest = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))
    path = os.path.join(REF_DIR, "nrms.py")
    if "nrms.py" not in manifest or not os.path.exists(path):
        raise FileNotFoundError("oracle/_ref/nrms.py is not staged (python oracle/stage_ref.py in the build container)")
    if hashlib.sha256(open(path, "rb").read()).hexdigest() != manifest["nrms.py"]["sha256"]:
        raise RuntimeError("oracle/_ref/nrms.py does not match its manifest: not the unmodified reference file")
    snoop, tools = types.ModuleType("torchsnooper"), types.ModuleType("tools")
    snoop.snoop = lambda *a, **k: (lambda f: f)
    tools.log_exec_time = lambda f: f
    saved = {k: sys.modules.get(k) for k in ("torchsnooper", "tools")}
    sys.modules["torchsnooper"], sys.modules["tools"] = snoop, tools
    try:
        spec = importlib.util.spec_from_file_location("ref_nrms_variant_staged", path)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return ref


class RefConfig:
    """The attributes nrms_v0.Model reads (config.py:30-57 + __nrms__ :65-88)."""

    def __init__(self, data_path, npz, word_embed_size, num_attention_heads, query_vector_dim, dropout, device):
        self.data_path, self.word_embedding_pretrained = data_path, npz
        self.word_embed_size, self.num_attention_heads = word_embed_size, num_attention_heads
        self.query_vector_dim, self.dropout, self.device = query_vector_dim, dropout, device


class ReferenceTrainer:
    """model + Adam + CrossEntropyLoss exactly as train_eval.py:166-181, stepped as :189-205."""

    def __init__(self, cfg: RefConfig, lr: float, seed: int = 42):
        on_gpu = torch.device(cfg.device).type == "cuda"
        self.ref = load(cpu_proxy=not on_gpu)
        torch.manual_seed(seed)                                   # run_demo.py:22
        self.model = self.ref.Model(cfg).to(cfg.device)           # run_demo.py:58
        self.model.train()                                        # train_eval.py:166
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr)   # :167
        self.criterion = nn.CrossEntropyLoss()                    # :181
        self.device = cfg.device

    def train_step(self, datas) -> float:
        model = self.model
        outputs = model(datas)                                    # :189
        model.zero_grad()                                         # :193
        y = torch.zeros(len(outputs)).long().to(self.device)      # :194
        loss = self.criterion(outputs, y)                         # :195
        value = loss.item()                                       # :198
        loss.backward()                                           # :204
        self.optimizer.step()                                     # :205
        return value
