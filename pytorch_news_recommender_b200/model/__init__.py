"""Model plugin discovery, as the reference's scripts expect it:

    from model import NRMS_V0            # run_demo.py:5  (intended: model.nrms_v0.Model)
    from model import Model; Model(config, args)   # run_v0.py:13,51 -> import_module('model.'+args.model.lower()).Model(config)
"""
from importlib import import_module

import torch
import torch.nn as nn

from .nrms_v0 import Model as NRMS_V0  # noqa: F401
from .nrms import Model as NRMS  # noqa: F401  (the BERT-vector sibling: reference model/__init__.py:1)


class Model(nn.Module):
    """Wrapper of reference model/__init__.py:13-38.  The reference's `data_parallel` branch is
    dead code; multi-GPU training here is one process per GPU through
    `pytorch_news_recommender_b200.engine.FusedTrainer` + torch.distributed (NCCL)."""

    def __init__(self, config, args):
        super().__init__()
        self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self.n_GPUs = getattr(args, 'n_GPUs', 1)
        module = import_module(__name__ + '.' + args.model.lower())
        self.model = module.Model(config).to(self.device)

    def forward(self, batch):
        return self.model(batch)
