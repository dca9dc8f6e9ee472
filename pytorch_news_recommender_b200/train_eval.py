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
from .engine import DeviceAdam, FusedTrainer

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


def log_res(config, step, auc):
    """train_eval.py:274-278, same signature and file format (`<time>_<auc>_:auc_<step>` appended to
    config.log_path/res.txt).  The reference calls it as `log_res(config, auc, total_batch)`, i.e.
    with the two values swapped; `train` below makes the same call, so the files read the same."""
    import os
    os.makedirs(config.log_path, exist_ok=True)
    with open(config.log_path + '/res.txt', 'a+') as f:
        f.write('{}_{}_:auc_{}\n'.format(time.strftime('%m-%d_%H.%M'), auc, step))


def warmup_lr(base_lr: float, i: int, warm_up_steps: int) -> float:
    """Learning rate of warm-up iteration i in the reference's loop (train_eval.py:64-98):
    GradualWarmupScheduler(multiplier=1, total_epoch=warm_up_steps) is stepped with `i` AFTER
    iteration i (lr_scheduler.py:39-40: base_lr * last_epoch / total_epoch), so iteration i runs
    with the rate set after iteration i-1: base_lr * max(i-1, 0) / warm_up_steps."""
    e = max(i - 1, 0)
    return base_lr if e > warm_up_steps else base_lr * float(e) / warm_up_steps


def _pick_step(config, model, fused: bool):
    """fused=True asks for the device-side training step: `FusedTrainer` (forward + loss + backward + Adam
    without an autograd tape) exists for the nrms_v0 plugin; every other plugin (the `nrms` sibling) runs the
    reference's statements over its autograd nodes with `DeviceAdam`.  fused=False is the literal reference
    loop with `torch.optim.Adam`.  Returns (use FusedTrainer, optimizer or None)."""
    from .model import nrms_v0
    inner = getattr(model, 'model', model)
    if fused and isinstance(inner, nrms_v0.Model):
        return True, None
    if fused:
        return False, DeviceAdam(model.parameters(), lr=config.learning_rate)
    return False, torch.optim.Adam(model.parameters(), lr=config.learning_rate)


def train(config, model, train_iter, dev_iter, y_true: Optional[Sequence[Sequence[int]]] = None,
          fused: bool = True, restore_train_mode: bool = False, log=print):
    """The reference's main loop (train_eval.py:34-153, entered from run_v0.py): optional warm-up
    (`config.warm_up`: 502 iterations with a linearly rising learning rate), then
    `config.num_epochs` epochs with an evaluation every `config.eval_step` batches and after every
    epoch, `log_res` after each evaluation, and a checkpoint whenever the dev AUC improves on
    `AUC_best` (initially 0.56) and `config.save_flag` is set.  `y_true` replaces the hard-coded
    read of './data_processed/dev_behaviors.csv'.

    Deviation kept from the reference on purpose: `evaluate` puts the model in eval mode and the
    reference never switches back (train_eval.py:230), so everything after the first evaluation
    trains WITHOUT dropout.  `restore_train_mode=True` opts out of that.  `plot_loss` (matplotlib,
    tools.py) is out of scope; the loss records are returned instead."""
    if y_true is not None:
        set_y_true(y_true)
    elif _y_true is None:
        set_y_true(load_y_true(config.data_path + 'dev_behaviors.csv'))
    start_time = time.time()
    model.train()
    fused, optimizer = _pick_step(config, model, fused)
    trainer = FusedTrainer(model, lr=config.learning_rate) if fused else None
    criterion = nn.CrossEntropyLoss()
    total_batch, AUC_best, loss_list, STEP_SIZE, improve = 0, 0.56, [], 100, '*'
    loss_records: List[float] = []

    def one_step(datas, lr):
        if fused:
            trainer.lr = lr
            trainer.step(datas)
            return trainer.last_loss()
        for g in optimizer.param_groups:
            g['lr'] = lr
        outputs = model(datas)
        model.zero_grad()
        y = torch.zeros(len(outputs)).long().to(outputs.device)
        loss = criterion(outputs, y)
        loss.backward()
        optimizer.step()
        return loss.item()

    def after_eval(auc, tag):
        nonlocal AUC_best
        log_res(config, auc, tag)                 # (sic: the reference passes them in this order)
        if auc > AUC_best:
            AUC_best = auc
            if config.save_flag:
                save_checkpoint(config, model, total_batch, AUC_best)
        if restore_train_mode:
            model.train()

    if getattr(config, 'warm_up', False):
        log('warm-up training...')
        for i, datas in enumerate(train_iter):
            loss_list.append(one_step(datas, warmup_lr(config.learning_rate, i, config.warm_up_steps)))
            if i % 100 == 0:
                msg = 'Warm-up Steps: {0:>6},  Train Loss: {1:>5.6},  Time: {2} {3}'
                log(msg.format(i, np.mean(loss_list), get_time_dif(start_time), improve))
                loss_list = []
            if i > 500:
                break
    auc = float('nan')
    for epoch in range(config.num_epochs):
        log('Epoch [{}/{}]'.format(epoch + 1, config.num_epochs))
        loss_records = []
        for i, datas in enumerate(train_iter):
            loss = one_step(datas, config.learning_rate)
            loss_list.append(loss)
            loss_records.append(loss)
            if total_batch % STEP_SIZE == 0:
                msg = 'Iter: {0:>6},  Train Loss: {1:>5.6},  Time: {2} {3}'
                log(msg.format(total_batch, np.mean(loss_list), get_time_dif(start_time), improve))
                loss_list = []
            total_batch += 1
            if total_batch % config.eval_step == 0 and total_batch > 0:
                auc = evaluate(config, model, dev_iter, AUC_best, log=log)
                after_eval(auc, total_batch)
        auc = evaluate(config, model, dev_iter, AUC_best, log=log)
        after_eval(auc, 'epoch_{}'.format(epoch))
    return {"auc": auc, "auc_best": AUC_best, "total_batch": total_batch, "loss_records": loss_records}


def train_demo(config, model, train_iter, dev_iter, y_true: Optional[Sequence[Sequence[int]]] = None,
               fused: bool = True, restore_train_mode: bool = False, log=print):
    """train_eval.py:156-215.  `y_true` replaces the hard-coded read of
    './data_processed/small_dev_behaviors.csv' when given.  As in the reference, the model is left
    in eval mode by `evaluate` (train_eval.py:230), so epochs after the first train without dropout;
    `restore_train_mode=True` switches back to train mode after each evaluation instead."""
    if y_true is not None:
        set_y_true(y_true)
    elif _y_true is None:
        set_y_true(load_y_true(config.data_path + 'small_dev_behaviors.csv'))
    log('result_length::::', len(_y_true))
    start_time = time.time()
    model.train()
    fused, optimizer = _pick_step(config, model, fused)
    trainer = FusedTrainer(model, lr=config.learning_rate) if fused else None
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
        if restore_train_mode:
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
    if ckpt_file is None:
        # the reference fails here too (train_eval.py:309: torch.load(save_path + None)); scoring the
        # current weights silently would write a submission nobody trained
        raise FileNotFoundError("test(): no checkpoint of '{}' with auc > 0.5 under {} and no ckpt_file given"
                                .format(config.model_name, config.save_path))
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
