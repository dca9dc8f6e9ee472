"""The train / scoring loops of the reference's `train_eval.py`, same names and call sequence
(`train_demo(config, model, train_iter, dev_iter)`, `evaluate(config, model, data_iter,
AUC_best)`: run_demo.py:61, train_eval.py:156-273), driving the B200 kernels.

What is kept: Adam(lr=config.learning_rate) + CrossEntropyLoss against label 0, a loss print
every 100 batches, one evaluation per epoch, positional pairing of score row i with label list
i, the mean of per-impression AUCs as the return value of `evaluate` (NaN-poisoned like
`np.mean` when an impression is single-class).
What changes underneath: the step runs through `FusedTrainer` (forward + CE + backward + Adam in
our kernels, no autograd graph) unless `fused=False`, which runs the literal reference lines
(`model(datas)`, `criterion`, `loss.backward()`, `optimizer.step()`) on the drop-in autograd
path; scores stay on the GPU and the metrics kernel replaces the fork-pool over sklearn.
"""
from __future__ import annotations

import time
from datetime import timedelta
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import evaluation
from .engine import FusedTrainer

_y_true: Optional[List[List[int]]] = None     # the reference keeps the dev labels in a module global
last_metrics = {}                             # AUC / MRR / nDCG@5 / nDCG@10 of the last evaluate()


def get_time_dif(start_time):
    """tools.py: elapsed wall time as a timedelta (seconds resolution)."""
    return timedelta(seconds=int(round(time.time() - start_time)))


def load_y_true(csv_path: str) -> List[List[int]]:
    """The `y_true` column of *_behaviors.csv: space separated 0/1 per impression
    (train_eval.py:157-159)."""
    import pandas as pd
    behaviors = pd.read_csv(csv_path)
    return [[int(_) for _ in x.split(' ')] for x in behaviors['y_true'].tolist()]


def set_y_true(y_true: Sequence[Sequence[int]]) -> None:
    global _y_true
    _y_true = [list(map(int, y)) for y in y_true]


def train_demo(config, model, train_iter, dev_iter, y_true: Optional[Sequence[Sequence[int]]] = None,
               fused: bool = True, log=print):
    """train_eval.py:156-215.  `y_true` replaces the hard-coded read of
    './data_processed/small_dev_behaviors.csv' when given."""
    if y_true is not None:
        set_y_true(y_true)
    elif _y_true is None:
        set_y_true(load_y_true(config.data_path + 'small_dev_behaviors.csv'))
    log('result_length::::', len(_y_true))
    start_time = time.time()
    model.train()
    trainer = FusedTrainer(model, lr=config.learning_rate) if fused else None
    optimizer = None if fused else torch.optim.Adam(model.parameters(), lr=config.learning_rate)
    criterion = nn.CrossEntropyLoss()
    total_batch, AUC_best, loss_list, STEP_SIZE, improve = 0, 0, [], 100, '*'
    auc = float('nan')
    for epoch in range(config.num_epochs):
        log('Epoch [{}/{}]'.format(epoch + 1, config.num_epochs))
        for i, datas in enumerate(train_iter):
            if fused:
                loss = trainer.step(datas)
            else:
                outputs = model(datas)
                model.zero_grad()
                y = torch.zeros(len(outputs)).long().to(outputs.device)
                loss = criterion(outputs, y)
                loss.backward()
                optimizer.step()
            loss_list.append(trainer.last_loss() if fused else loss.item())
            if total_batch % STEP_SIZE == 0:
                msg = 'Iter: {0:>6},  Train Loss: {1:>5.6},  Time: {2} {3}'
                log(msg.format(total_batch, np.mean(loss_list), get_time_dif(start_time), improve))
                loss_list = []
            total_batch += 1
        auc = evaluate(config, model, dev_iter, AUC_best, log=log)
        model.train()
    return auc


def evaluate(config, model, data_iter, AUC_best, y_true: Optional[Sequence[Sequence[int]]] = None, log=print):
    """train_eval.py:229-273: eval-mode forward over the loader, concatenate the [B, S] score
    rows, per-impression AUC of rank_score[i][:len(y_true[i])], mean."""
    if y_true is not None:
        set_y_true(y_true)
    if _y_true is None:
        raise RuntimeError("evaluate: no dev labels; pass y_true or call set_y_true/train_demo first")
    model.eval()
    scores = []
    with torch.no_grad():
        for datas in data_iter:
            scores.append(model(datas))
    rank_score = torch.cat(scores, 0)
    m = evaluation.evaluate_scores(rank_score, _y_true).cpu().numpy()
    AUC = float(np.mean(m[:, 0]))
    last_metrics.update(auc=AUC, mrr=float(np.mean(m[:, 1])), ndcg5=float(np.mean(m[:, 2])),
                        ndcg10=float(np.mean(m[:, 3])), auc_nanmean=float(np.nanmean(m[:, 0])),
                        n_impressions=int(m.shape[0]))
    log('AUC:', AUC)
    return AUC


# ---- checkpoints and the test-set ranking writer (train_eval.py:139-142, 279-341) -------------
def checkpoint_name(config, total_batch, auc) -> str:
    """The reference's checkpoint file name (train_eval.py:142)."""
    return 'T{}_{}_epoch{}_iter_{}_auc_{:.3f}.ckpt'.format(time.strftime('%m-%d_%H.%M'), config.model_name,
                                                           config.num_epochs, total_batch, auc)


def save_checkpoint(config, model, total_batch, auc) -> str:
    """`torch.save(model.state_dict(), config.save_path + name)`: the state_dict keys and shapes are
    the reference's, so either implementation loads the other's files."""
    import os
    os.makedirs(config.save_path, exist_ok=True)
    path = config.save_path + checkpoint_name(config, total_batch, auc)
    torch.save(model.state_dict(), path)
    return path


def best_checkpoint(config) -> Optional[str]:
    """train_eval.py:296-303: the checkpoint of `config.model_name` with the highest AUC in its
    file name (only files above 0.5 qualify, as there)."""
    import os
    auc_best, ckpt_file = 0.5, None
    for ckpt in sorted(os.listdir(config.save_path)):
        if config.model_name in ckpt and ckpt.endswith('.ckpt'):
            try:
                tmp_auc = float(ckpt[:-len('.ckpt')].split('_')[-1])
            except ValueError:
                continue
            if tmp_auc > auc_best:
                auc_best, ckpt_file = tmp_auc, ckpt
    return ckpt_file


def test(config, model, data_iter, ckpt_file=None, test_list_nums: Optional[Sequence[int]] = None,
         out_dir: str = '.', log=print) -> str:
    """train_eval.py:294-341: load a checkpoint, score the test impressions, write
    `sumbit_<model>_<time>.txt` with one line per impression: `<1-based index> [r1,r2,...]` where
    r_j is the rank of candidate j among the impression's `test_list_nums[i]` real candidates.
    The rank lists are computed on the GPU (`nrms_rank_positions`); only they cross to the host."""
    import os
    from . import ops
    if ckpt_file is None:
        ckpt_file = best_checkpoint(config)
    if ckpt_file is not None:
        path = ckpt_file if os.path.isabs(ckpt_file) or os.path.exists(ckpt_file) else config.save_path + ckpt_file
        model.load_state_dict(torch.load(path, map_location='cpu'))
        log('load the ckpt_file:{}'.format(ckpt_file))
    if test_list_nums is None:
        import pickle
        with open(config.data_path + 'test_imps_list.pkl', 'rb') as f:
            test_list_nums = pickle.load(f)
    model.eval()
    scores = []
    with torch.no_grad():
        for datas in data_iter:
            scores.append(model(datas))
    rank_score = torch.cat(scores, 0)
    lens = torch.as_tensor(list(test_list_nums), dtype=torch.int64, device=rank_score.device)
    if lens.numel() != rank_score.shape[0]:
        raise ValueError(f"{lens.numel()} impression lengths for {rank_score.shape[0]} score rows")
    ranks = ops.rank_positions(rank_score, lens).cpu().numpy()
    file_name = os.path.join(out_dir, 'sumbit_{}_{}.txt'.format(config.model_name,
                                                                time.strftime('%m-%d_%H.%M', time.localtime())))
    with open(file_name, 'w') as f:
        for i, n in enumerate(test_list_nums):
            f.write(str(i + 1) + ' ')
            f.write(str(ranks[i, :n].tolist()).replace(' ', '') + '\n')
    log('saved to {}'.format(file_name))
    return file_name
